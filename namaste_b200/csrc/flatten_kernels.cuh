// Light-curve detrending of namaste/k2flatten.py (ReduceNoise / dopolyfit, :5-17, :51-84) on the GPU:
// one thread block per detrending step.  The block holds the step's fitting window (the points within
// winsize/2 of the step centre, the step's own box excluded) in shared memory and runs the
// reference's clipped weighted polynomial fit: an initial fit, then `niter` rounds of
//   offset = |y - p(x)| / err over the whole window;
//   keep offset < sigclip if more than int(0.8 * npoints) points pass, else keep offset < mean(offset);
//   refit,
// and finally writes y - p(x) + 1 for the box points this step owns (a later step overwrites an
// earlier one where boxes overlap, as in the reference's loop).
// np.polyfit(x, y, w = 1/err^2) minimises sum (w_i (y_i - p(x_i)))^2 through an SVD of the scaled
// Vandermonde matrix in the raw times; here the same least-squares problem is solved in the centred,
// scaled variable u = (x - xc) / h by Cholesky on the normal equations, which is well conditioned
// (|u| <= 1) and yields the same polynomial *values* to rounding.
#pragma once
#include <cuda_runtime.h>

namespace nmb {

constexpr int kFlatMaxDeg = 5;
constexpr int kFlatThreads = 128;

struct FlatStep {
    int wlo, blo, bhi, whi;  // window = [wlo, blo) U [bhi, whi), box = [blo, bhi)
    int olo, ohi;            // box indices this step owns (writes)
    double xc, h;            // centre and half-width used to scale the abscissa
};

__device__ __forceinline__ double block_sum(double v, double* s_red /*[THREADS/32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kFlatThreads / 32; ++i) t += s_red[i];
    return t;
}

// weighted least-squares polynomial of degree deg through the masked window points; coefficients
// c[0..deg] in powers of u (every thread computes the same solve from the block-reduced moments)
__device__ void fit_poly(const double* su, const double* sy, const double* sw, const unsigned char* keep, int n, int deg,
                         double* c, double* s_red) {
    double mom[2 * kFlatMaxDeg + 1], rhs[kFlatMaxDeg + 1];
    for (int k = 0; k <= 2 * deg; ++k) mom[k] = 0.0;
    for (int k = 0; k <= deg; ++k) rhs[k] = 0.0;
    for (int i = threadIdx.x; i < n; i += kFlatThreads) {
        if (!keep[i]) continue;
        const double w2 = sw[i] * sw[i];
        double p = w2;
        for (int k = 0; k <= 2 * deg; ++k) {
            mom[k] += p;
            if (k <= deg) rhs[k] += p * sy[i];
            p *= su[i];
        }
    }
    for (int k = 0; k <= 2 * deg; ++k) mom[k] = block_sum(mom[k], s_red);
    for (int k = 0; k <= deg; ++k) rhs[k] = block_sum(rhs[k], s_red);
    // Cholesky of the (deg+1) x (deg+1) moment matrix A[j][k] = mom[j + k]
    double L[kFlatMaxDeg + 1][kFlatMaxDeg + 1];
    const int m = deg + 1;
    for (int j = 0; j < m; ++j) {
        for (int k = 0; k <= j; ++k) {
            double s = mom[j + k];
            for (int q = 0; q < k; ++q) s -= L[j][q] * L[k][q];
            L[j][k] = (j == k) ? sqrt(s) : s / L[k][k];
        }
    }
    double yv[kFlatMaxDeg + 1];
    for (int j = 0; j < m; ++j) {
        double s = rhs[j];
        for (int q = 0; q < j; ++q) s -= L[j][q] * yv[q];
        yv[j] = s / L[j][j];
    }
    for (int j = m - 1; j >= 0; --j) {
        double s = yv[j];
        for (int q = j + 1; q < m; ++q) s -= L[q][j] * c[q];
        c[j] = s / L[j][j];
    }
}

__device__ __forceinline__ double poly_eval(const double* c, int deg, double u) {
    double p = c[deg];
    for (int k = deg - 1; k >= 0; --k) p = p * u + c[k];
    return p;
}

__global__ void __launch_bounds__(kFlatThreads)
flatten_kernel(const double* __restrict__ t, const double* __restrict__ f, const double* __restrict__ e,
               const FlatStep* __restrict__ steps, int deg, int niter, double sigclip, int maxwin,
               double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* su = reinterpret_cast<double*>(smem_raw);  // scaled abscissa
    double* sy = su + maxwin;
    double* se = sy + maxwin;                          // errors
    double* sw = se + maxwin;                          // polyfit weights 1 / err^2
    double* soff = sw + maxwin;
    unsigned char* keep = reinterpret_cast<unsigned char*>(soff + maxwin);
    __shared__ double s_red[kFlatThreads / 32];
    const FlatStep st = steps[blockIdx.x];
    const int n1 = max(0, st.blo - st.wlo), n2 = max(0, st.whi - st.bhi), n = n1 + n2;
    for (int i = threadIdx.x; i < n; i += kFlatThreads) {
        const int g = i < n1 ? st.wlo + i : st.bhi + (i - n1);
        su[i] = (t[g] - st.xc) / st.h;
        sy[i] = f[g];
        se[i] = e[g];
        sw[i] = 1.0 / (e[g] * e[g]);
        keep[i] = 1;
    }
    __syncthreads();
    double c[kFlatMaxDeg + 1];
    fit_poly(su, sy, sw, keep, n, deg, c, s_red);
    for (int it = 0; it < niter; ++it) {
        double osum = 0.0, npass = 0.0;
        for (int i = threadIdx.x; i < n; i += kFlatThreads) {
            const double o = fabs(sy[i] - poly_eval(c, deg, su[i])) / se[i];
            soff[i] = o;
            osum += o;
            npass += (o < sigclip) ? 1.0 : 0.0;
        }
        osum = block_sum(osum, s_red);
        npass = block_sum(npass, s_red);
        const bool enough = (int)npass > (int)(0.8 * n);
        const double thr = enough ? sigclip : osum / n;
        for (int i = threadIdx.x; i < n; i += kFlatThreads) keep[i] = soff[i] < thr;
        __syncthreads();
        fit_poly(su, sy, sw, keep, n, deg, c, s_red);
    }
    for (int g = st.olo + threadIdx.x; g < st.ohi; g += kFlatThreads)
        out[g] = f[g] - poly_eval(c, deg, (t[g] - st.xc) / st.h) + 1;
}

}  // namespace nmb
