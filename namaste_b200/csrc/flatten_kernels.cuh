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
#include <limits.h>

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

// ---- the batch form: windows, boxes and owned ranges of every detrending step of every candidate, on the device --
// formwindow (k2flatten.py:19-40) for step s of candidate c: the window is `winsize` wide around the step centre
// cent = s / nsteps * lenlc + stepsize / 2 (:66-67); if it runs into a gap wider than `gap` on one side only it is
// shifted to start (or end) at the gap, keeping its width.  searchsorted(side='left') = first index with t >= x.
__device__ __forceinline__ int lower_bound_t(const double* __restrict__ t, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ void step_box(const double* __restrict__ t, int n, long long s, long long nsteps, double lenlc,
                                         double stepsize, double& cent, int& blo, int& bhi) {
    cent = (double)s / (double)nsteps * lenlc + stepsize / 2.;
    blo = lower_bound_t(t, n, cent - stepsize / 2.);
    bhi = lower_bound_t(t, n, cent + stepsize / 2.);
}
// step_off [ncand + 1]: first step of each candidate; steps [step_off[ncand]] out, indices are offsets into the flat
// [ncand][stride] arrays; info[0] = largest window (atomicMax), info[1] = 1 + first step whose window is too small
__global__ void flat_windows_kernel(const double* __restrict__ t, const long long* __restrict__ n, long long stride, int ncand,
                                    const long long* __restrict__ step_off, double winsize, double stepsize, double gap, int deg,
                                    FlatStep* __restrict__ steps, int* __restrict__ info) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= step_off[ncand]) return;
    int c = 0;
    {   // candidate of this step: last c with step_off[c] <= g
        int lo = 0, hi = ncand;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (step_off[mid] <= g) lo = mid; else hi = mid;
        }
        c = lo;
    }
    const long long s = g - step_off[c], nsteps = step_off[c + 1] - step_off[c];
    const int nc = (int)n[c];
    const double* tc = t + (long long)c * stride;
    const double lenlc = tc[nc - 1];
    double cent;
    int blo, bhi;
    step_box(tc, nc, s, nsteps, lenlc, stepsize, cent, blo, bhi);
    const double lo_edge = cent - winsize / 2., hi_edge = cent + winsize / 2.;
    int wlo = lower_bound_t(tc, nc, lo_edge), whi = lower_bound_t(tc, nc, hi_edge);
    if (whi == nc) whi -= 1;
    const bool gap_above = tc[whi] < hi_edge - gap;
    const bool gap_below = tc[wlo] > lo_edge + gap;
    if (gap_above && !gap_below) wlo = lower_bound_t(tc, nc, tc[whi] - winsize);
    else if (gap_below && !gap_above) whi = lower_bound_t(tc, nc, tc[wlo] + winsize);
    // the part of the box that survives the reference's sequential loop: a later step overwrites an earlier one
    // where boxes overlap, i.e. from the first later non-empty box's lower bound on
    int nxt = INT_MAX;
    for (long long s2 = s + 1; s2 < nsteps; ++s2) {
        double c2;
        int b0, b1;
        step_box(tc, nc, s2, nsteps, lenlc, stepsize, c2, b0, b1);
        if (b1 > b0) { nxt = b0; break; }
    }
    FlatStep st;
    const int base = (int)((long long)c * stride);
    const int n1 = max(0, blo - wlo), n2 = max(0, whi - bhi);
    st.wlo = base + wlo; st.blo = base + blo; st.bhi = base + bhi; st.whi = base + whi;
    st.olo = base + blo; st.ohi = base + max(blo, min(bhi, nxt));
    if (wlo > blo) st.wlo = st.blo;   // (empty lower part)
    if (whi < bhi) st.whi = st.bhi;   // (empty upper part)
    const double xa = n1 > 0 ? tc[wlo] : tc[min(bhi, nc - 1)], xb = n2 > 0 ? tc[whi - 1] : tc[max(blo - 1, 0)];
    st.xc = 0.5 * (xa + xb);
    st.h = fmax(0.5 * (xb - xa), 1e-12);
    steps[g] = st;
    atomicMax(info, n1 + n2);
    if (n1 + n2 < deg + 1) atomicCAS(info + 1, 0, (int)min(g + 1, (long long)INT_MAX));
}

// ---- AnomCutDiff (namaste/namaste.py:1459-1467), one block per candidate ------------------------------------------
// keep[0] = keep[n-1] = 1; an interior point is dropped only when it differs from BOTH neighbours by at least
// `thresh` times the median of |diff(flux[1:])| and the two steps have opposite signs.  The median (mean of the
// two middle order statistics for an even count, like np.median) comes from an 8-pass radix selection on the
// order-preserving integer image of the differences, recomputed on the fly.
constexpr int kCutThreads = 256;
__global__ void __launch_bounds__(kCutThreads)
anomcut_kernel(const double* __restrict__ flux, const long long* __restrict__ n, long long stride, double thresh,
               unsigned char* __restrict__ keep) {
    __shared__ int hist[2][256];
    __shared__ unsigned long long prefix[2];
    __shared__ int rank[2];
    const double* f = flux + blockIdx.x * stride;
    unsigned char* kp = keep + blockIdx.x * stride;
    const int nc = (int)n[blockIdx.x];
    const int m = nc - 2;  // number of differences diff(flux[1:])
    if (m < 1) {
        for (int i = threadIdx.x; i < nc; i += kCutThreads) kp[i] = 1;
        return;
    }
    if (threadIdx.x < 2) {
        prefix[threadIdx.x] = 0;
        rank[threadIdx.x] = threadIdx.x == 0 ? (m - 1) / 2 : m / 2;  // the two middle order statistics
    }
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 512; i += kCutThreads) (&hist[0][0])[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += kCutThreads) {
            const unsigned long long k = (unsigned long long)__double_as_longlong(fabs(f[i + 2] - f[i + 1]));  // >= 0: ordered as is
            const int digit = (int)((k >> shift) & 0xff);
            for (int j = 0; j < 2; ++j)
                if (pass == 0 || (k >> (shift + 8)) == prefix[j]) atomicAdd(&hist[j][digit], 1);
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int j = threadIdx.x;
            int r = rank[j], b = 0;
            while (b < 255 && r >= hist[j][b]) { r -= hist[j][b]; ++b; }
            rank[j] = r;
            prefix[j] = (prefix[j] << 8) | (unsigned long long)b;
        }
        __syncthreads();
    }
    const double lo = __longlong_as_double((long long)prefix[0]), hi = __longlong_as_double((long long)prefix[1]);
    const double med = (m & 1) ? hi : (lo + hi) / 2.0;  // np.median: mean of the middle pair
    for (int i = threadIdx.x; i < nc; i += kCutThreads) {
        unsigned char k = 1;
        if (i > 0 && i < nc - 1) {
            const double ahead = (f[i + 1] - f[i]) / med, behind = (f[i] - f[i - 1]) / med;
            k = (ahead * behind > 0) || (fabs(ahead) < thresh) || (fabs(behind) < thresh);
        }
        kp[i] = k;
    }
}

// ---- window mask of RunMCMC (namaste/namaste.py:550-556) for a batch ----------------------------------------------
__global__ void window_mask_kernel(const double* __restrict__ t, const long long* __restrict__ n, long long stride, int ncand,
                                   const double* __restrict__ tcen, const double* __restrict__ width, int two_sided,
                                   unsigned char* __restrict__ keep) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= (long long)ncand * stride) return;
    const int c = (int)(g / stride);
    const long long i = g - (long long)c * stride;
    unsigned char k = 0;
    if (i < n[c]) {
        const double d = t[g] - tcen[c];
        // the reference takes abs() of a boolean: everything before tcen + width is kept (quirk Q6)
        k = two_sided ? (fabs(d) < width[c]) : (d < width[c]);
    }
    keep[g] = k;
}

}  // namespace nmb
