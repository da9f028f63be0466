// The inner loop of sweep_walkers_kernel in isolation: per timestamp S = 15 independent DFMA y_k = fma(V, t, c_k), the
// AND of their high words (LOP3), one DADD, eight timestamps per trip with one bit test and branch.  Variants tell what
// an instruction costs inside this mix on the part (cycles per timestamp and sub-partition):
//   MODE 0  the loop as in the kernel (t_i from shared memory, broadcast LDS.128 per two timestamps)
//   MODE 1  t_i from the constant bank (kernel parameter array): one register operand less per DFMA
//   MODE 2  DFMA only: 15 dependent-free chains y_k = fma(V, y_k, c_k), no logic, no loads
//   MODE 3  LOP3 over one high and one low word (operands of both register parities: bank-conflict probe)
//   MODE 5  an FP64 consumer (DADD) instead of the logic
//   MODE 6  LOP3 with one operand passed through a funnel shift
//   MODE 4  logic only: the 8 LOP3 per timestamp on integer registers, no DFMA
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o sweep_mix sweep_mix.cu
#include <cstdio>
constexpr int S = 15, NT = 8;
struct Times { double t[64]; };

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) mix(double* out, int iters, double vs, double c0, Times tc) {
    __shared__ __align__(16) double sm[64], tm[64];
    if (threadIdx.x < 64) { sm[threadIdx.x] = 100.0 + threadIdx.x * 0.02; tm[threadIdx.x] = 1e-3 * threadIdx.x; }
    __syncthreads();
    double ck[S];
#pragma unroll
    for (int k = 0; k < S; ++k) ck[k] = c0 + 1e-3 * k + 1e-6 * threadIdx.x;
    double acc = 0.0;
    int hits = 0;
    if (MODE == 2) {
        double y[S];
#pragma unroll
        for (int k = 0; k < S; ++k) y[k] = ck[k];
        for (int it = 0; it < iters * NT; ++it) {
#pragma unroll
            for (int k = 0; k < S; ++k) y[k] = fma(vs, y[k], ck[k]);
        }
#pragma unroll
        for (int k = 0; k < S; ++k) acc += y[k];
    } else if (MODE == 4) {
        unsigned h[S];
#pragma unroll
        for (int k = 0; k < S; ++k) h[k] = __double2hiint(ck[k]) | 0x40000000u;
        unsigned all = 0xffffffffu;
        for (int it = 0; it < iters * NT; ++it) {
            unsigned a = h[S - 1] ^ (unsigned)it;
#pragma unroll
            for (int k = 0; k + 1 < S; k += 2) a &= (h[k] + it) & h[k + 1];
            all &= a;
        }
        hits = all;
    } else {
        for (int it = 0; it < iters; ++it) {
            const int i0 = (it * NT) & 63;
            double tt[NT], mm[NT];
            unsigned q[NT];
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
                if (MODE == 1) {
                    tt[j] = tc.t[j]; tt[j + 1] = tc.t[j + 1];
                } else {
                    const double2 tv = *reinterpret_cast<const double2*>(sm + i0 + j);
                    tt[j] = tv.x; tt[j + 1] = tv.y;
                }
                const double2 mv = *reinterpret_cast<const double2*>(tm + i0 + j);
                mm[j] = mv.x; mm[j + 1] = mv.y;
            }
            unsigned qall = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                unsigned h[S];
#pragma unroll
                for (int k = 0; k < S; ++k) h[k] = (unsigned)__double2hiint(fma(vs, tt[j], ck[k]));
                unsigned a = h[S - 1];
                if (MODE == 3) {   // one high and one low word per LOP3: operands of both register parities
                    unsigned l[S];
#pragma unroll
                    for (int k = 0; k < S; ++k) l[k] = (unsigned)__double2loint(fma(vs, tt[j], ck[k]));
#pragma unroll
                    for (int k = 0; k + 1 < S; k += 2) a &= h[k] & l[k + 1];
                } else if (MODE == 5) {   // FP64 consumer instead of logic: the sum of the samples
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < S; ++k) sacc += fma(vs, tt[j], ck[k]);
                    a = (unsigned)__double2hiint(sacc);
                } else if (MODE == 6) {   // shifts the high words into separate registers first (SHF), then LOP3
#pragma unroll
                    for (int k = 0; k + 1 < S; k += 2) a &= __funnelshift_r(h[k], h[k], 0) & h[k + 1];
                } else {
#pragma unroll
                    for (int k = 0; k + 1 < S; k += 2) a &= h[k] & h[k + 1];
                }
                q[j] = a;
                qall &= a;
            }
            if (qall & 0x40000000u) {
#pragma unroll
                for (int j = 0; j < NT; ++j) acc += mm[j];
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (q[j] & 0x40000000u) acc += mm[j]; else hits++;
                }
            }
        }
    }
    if (acc == 12345.678 || hits == 77) out[0] = acc;
}

template <int MODE, int MINB>
void run(const char* what) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    Times tc; for (int i = 0; i < 64; ++i) tc.t[i] = 100.0 + 0.02 * i;
    const int iters = 2048, blocks = 148 * MINB, threads = 128;
    // V t + c_k ~ 5e3: every sample far outside the window, like the common trip of the sweep
    mix<MODE, MINB><<<blocks, threads>>>(d, iters, 50.0, 7.0, tc);
    cudaEventRecord(e0);
    mix<MODE, MINB><<<blocks, threads>>>(d, iters, 50.0, 7.0, tc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ts_per_smsp = (double)MINB * iters * NT;   // warp-timestamps per sub-partition
    printf("%-44s %d warps/SMSP: %.3f ms, %.1f cycles per timestamp per SMSP (at 1.965 GHz)\n", what, MINB, ms,
           ms * 1e-3 * 1.965e9 / ts_per_smsp);
    cudaFree(d);
}
int main() {
    run<0, 6>("kernel loop (LDS t, 15 DFMA + 8 LOP3)");
    run<0, 4>("kernel loop (LDS t, 15 DFMA + 8 LOP3)");
    run<0, 8>("kernel loop (LDS t, 15 DFMA + 8 LOP3)");
    run<2, 6>("15 DFMA only");
    run<2, 4>("15 DFMA only");
    run<3, 6>("LOP3 over high and low words");
    run<5, 6>("15 DFMA + 15 DADD, no logic");
    run<6, 6>("LOP3 with one shifted operand");
    run<4, 6>("8 LOP3 only");
    return 0;
}
