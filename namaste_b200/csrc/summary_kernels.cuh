// Posterior reduction on the device-resident chain (namaste/namaste.py:597-608 trimming, percentile
// summaries as used at :643-652): exact order statistics by 8-pass radix selection on an
// order-preserving integer image of the doubles, so a [W, T, ndim] chain never has to leave HBM.
//   walker_select_kernel : one block per walker, order statistics of lnprobability[w, ncut:]
//                          (the 50th / 95th percentile inputs of the failed-walker filter);
//   column_hist_kernel + column_pick_kernel : order statistics of every column of
//                          chain[good, ncut:, :] (+ the lnprob column), all queries at once.
// The interpolation between the two neighbouring order statistics (NumPy's "linear" method) and the
// median over walkers are done by the host on these few numbers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmb {

__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

constexpr int kSelThreads = 256;
constexpr int kSelMaxQ = 8;  // order statistics per walker row

// values: row pointer = base + row * stride + first; L values per row; ranks[nq] (0-based, ascending
// order); out[row][nq]
__global__ void __launch_bounds__(kSelThreads)
walker_select_kernel(const double* __restrict__ base, long long stride, long long first, int L, int nq,
                     const int* __restrict__ ranks, double* __restrict__ out) {
    __shared__ int hist[kSelMaxQ][256];
    __shared__ unsigned long long prefix[kSelMaxQ];
    __shared__ int rank[kSelMaxQ];
    const double* row = base + (long long)blockIdx.x * stride + first;
    if (threadIdx.x < nq) { prefix[threadIdx.x] = 0; rank[threadIdx.x] = ranks[threadIdx.x]; }
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < nq * 256; i += kSelThreads) (&hist[0][0])[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += kSelThreads) {
            const unsigned long long k = order_key(row[i]);
            const int digit = (int)((k >> shift) & 0xff);
            for (int j = 0; j < nq; ++j)
                if (pass == 0 || (k >> (shift + 8)) == prefix[j]) atomicAdd(&hist[j][digit], 1);
        }
        __syncthreads();
        if (threadIdx.x < nq) {
            const int j = threadIdx.x;
            int r = rank[j], b = 0;
            while (b < 255 && r >= hist[j][b]) { r -= hist[j][b]; ++b; }
            rank[j] = r;
            prefix[j] = (prefix[j] << 8) | (unsigned long long)b;
        }
        __syncthreads();
    }
    if (threadIdx.x < nq) out[(long long)blockIdx.x * nq + threadIdx.x] = key_value(prefix[threadIdx.x]);
}

struct ColQuery {
    unsigned long long prefix;
    long long rank;
    int col;
    int pad;
};

// one pass of the selection over chain[good, ncut:, :] (+ lnprob as column ndim): every block
// histograms the current digit of the elements that still match each query's prefix
__global__ void __launch_bounds__(kSelThreads)
column_hist_kernel(const double* __restrict__ chain, const double* __restrict__ lnp, const unsigned char* __restrict__ good,
                   long long rows, long long tcap, long long ncut, long long tlen, int ndim, int pass,
                   const ColQuery* __restrict__ queries, int nqueries, unsigned int* __restrict__ ghist) {
    extern __shared__ unsigned int shist[];  // [nqueries][256]
    for (int i = threadIdx.x; i < nqueries * 256; i += kSelThreads) shist[i] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const long long L = tlen - ncut, total = rows * L;
    for (long long e = blockIdx.x * (long long)kSelThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kSelThreads) {
        const long long w = e / L, tt = ncut + (e - w * L);
        if (!good[w]) continue;
        const double* rowp = chain + (w * tcap + tt) * ndim;
        const double lv = lnp[w * tcap + tt];
        for (int j = 0; j < nqueries; ++j) {
            const ColQuery q = queries[j];
            const unsigned long long k = order_key(q.col < ndim ? rowp[q.col] : lv);
            if (pass == 0 || (k >> (shift + 8)) == q.prefix) atomicAdd(&shist[j * 256 + (int)((k >> shift) & 0xff)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nqueries * 256; i += kSelThreads)
        if (shist[i]) atomicAdd(&ghist[i], shist[i]);
}

__global__ void column_pick_kernel(ColQuery* __restrict__ queries, int nqueries, unsigned int* __restrict__ ghist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nqueries) return;
    long long r = queries[j].rank;
    int b = 0;
    unsigned int* h = ghist + j * 256;
    while (b < 255 && r >= (long long)h[b]) { r -= h[b]; ++b; }
    queries[j].rank = r;
    queries[j].prefix = (queries[j].prefix << 8) | (unsigned long long)b;
    for (int i = 0; i < 256; ++i) h[i] = 0;
}

}  // namespace nmb
