// Host build of the device arithmetic (namaste_b200/csrc/transit_math.cuh) for the CPU test suite:
// g++ compiles the very same header with NMB_HOST_EMULATION, so formula or fma-placement mistakes in
// the strict, relaxed and FP32 hot paths show up without a GPU.  Differences from the GPU build: the
// MUFU seeds are emulated (24-bit), libm is glibc's.  Test infrastructure, never loaded by the product.
#define NMB_HOST_EMULATION 1
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __constant__ const
#define __launch_bounds__(...)
#include <cmath>
#include <cstdint>
using std::fma;
static inline float __frcp_rn(float x) { return 1.0f / x; }
#define __log2f(x) log2f(x)
#include "../../namaste_b200/csrc/transit_math.cuh"

extern "C" {
// path: 0 general IEEE (occultquad_point), 1 strict hot path, 3 relaxed hot path, 2 FP32 hot path (as 1 - depth).
// Values the hot paths decline (return false / non-finite) come back as NaN with handled[i] = 0.
void nmb_host_occultquad(const double* z, int64_t n, double rprs, double ld1, double ld2, int path, double* F,
                         int* handled) {
    nmb::QuadConsts c = nmb::make_quad_consts(rprs, ld1, ld2);
    nmb::finish_quad_consts(c);
    const nmb::QuadConstsF cf = nmb::to_f32(c);
    for (int64_t i = 0; i < n; ++i) {
        double f = NAN;
        bool ok = true;
        if (path == 0) {
            f = nmb::occultquad_point<false, nmb::IeeeOps>(z[i], c, nullptr, nullptr, nullptr);
        } else if (path == 1) {
            ok = nmb::occultquad_hot(z[i], c, f) && std::fabs(f) <= 1e300;
        } else if (path == 3) {
            ok = nmb::occultquad_relaxed(z[i], c, f) && std::fabs(f) <= 1e300;
        } else {
            float d = 0.f;
            ok = nmb::occultquad_hot_f32((float)z[i], cf, d) && std::fabs(d) <= 1e30f;
            f = 1.0 - (double)d;
        }
        F[i] = ok ? f : NAN;
        handled[i] = ok ? 1 : 0;
    }
}
}
