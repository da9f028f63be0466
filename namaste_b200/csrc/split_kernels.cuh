// One ensemble split by walker over several GPUs (BASELINE config 5): the exchange of the half-ensemble
// slices that each rank updated, device to device over NVLink, without NCCL and without the host.
//
// The reference has no counterpart (its only parallelism is emcee's process pool, namaste/namaste.py:583/590);
// what must be preserved is emcee's half-ensemble dependency: the proposals of half-step h+1 read the
// positions that ALL walkers of half h have after half-step h (namaste.py:590-594 -> emcee stretch move).
//
// Every rank keeps a replica of the ensemble (pos, lnp) in one allocation that its peers map through CUDA
// IPC.  After a rank has updated its slice of half h, `split_exchange_kernel`
//   1. stores the slice's rows into every peer's replica (plain coalesced stores to peer addresses),
//   2. fences at system scope; the last block to finish publishes the epoch 2*step + half + 1 in every
//      peer's arrival-flag array (st.release.sys), and
//   3. waits (ld.acquire.sys on its own flag array) until every peer's epoch has arrived, i.e. until every
//      other slice of half h has landed in this rank's replica.
// The kernel is launched in the sampler's stream between the half-steps, chained by programmatic dependent
// launch like the rest of the pipeline; the next half-step's kernels read the replica only after it has
// completed.  A peer can be at most one half-step ahead (it needs this rank's flag to go on), and in
// half-step h a rank reads only rows of half 1-h while peers write only rows of half h, so no row is
// read and written concurrently.  Ranks run on different GPUs (one process per GPU): the spin never waits
// for a kernel that needs the same device.  A time-out (10 s of globaltimer) sets an error word instead of
// hanging the device if a peer dies.
#pragma once
#include <stdint.h>

namespace nmb {

constexpr int kMaxPeers = 16;

struct SplitPeers {
    int world, rank;
    double* pos[kMaxPeers];  // peer replicas (null for self)
    double* lnp[kMaxPeers];
    int* flags[kMaxPeers];   // peer arrival-flag arrays: flags[p][r] = last epoch rank r completed
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// rows [row0, row0 + nk) of candidate-major pos [rows][ndim] / lnp [rows] -> every peer; then flags.
// my_flags: this rank's flag array [kMaxPeers], followed by [kMaxPeers] ticket and [kMaxPeers + 1] error.
__global__ void __launch_bounds__(256)
split_exchange_kernel(SplitPeers pp, const double* __restrict__ pos, const double* __restrict__ lnp, int ndim,
                      long long row0, int nk, int epoch, int* my_flags, unsigned long long timeout_ns) {
    pdl_wait();  // the half-step's accept/reject writes are complete and visible
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x, gid = (long long)blockIdx.x * blockDim.x + tid;
    const long long nel = (long long)nk * ndim, base = row0 * ndim;
    for (int p = 0; p < pp.world; ++p) {
        if (p == pp.rank) continue;
        double* __restrict__ dpos = pp.pos[p];
        double* __restrict__ dlnp = pp.lnp[p];
        for (long long i = gid; i < nel; i += gsz) dpos[base + i] = pos[base + i];
        for (long long i = gid; i < nk; i += gsz) dlnp[row0 + i] = lnp[row0 + i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const int t = atomicAdd(my_flags + kMaxPeers, 1);
        s_last = (t == (int)gridDim.x - 1);
        if (s_last) my_flags[kMaxPeers] = 0;
    }
    __syncthreads();
    pdl_trigger();
    if (!s_last) return;
    __threadfence_system();
    if (tid < pp.world && tid != pp.rank) {
        st_release_sys(pp.flags[tid] + pp.rank, epoch);
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(my_flags + tid) < epoch) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                atomicExch(my_flags + kMaxPeers + 1, 1 + tid);
                break;
            }
            __nanosleep(200);
        }
    }
}

// NCCL transport: the rank's slice of half `half` as [nk][ndim + 1] rows (position, lnprob) ...
__global__ void split_pack_kernel(const double* __restrict__ pos, const double* __restrict__ lnp, int ndim, long long row0,
                                  int nk, double* __restrict__ out) {
    const int nd1 = ndim + 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nk * nd1) return;
    const long long r = i / nd1;
    const int d = (int)(i - r * nd1);
    out[i] = d < ndim ? pos[(row0 + r) * ndim + d] : lnp[row0 + r];
}
// ... and the gathered [world][nk][ndim + 1] rows of a half back into the replica (rows half0 + w * nk + r)
__global__ void split_unpack_kernel(double* __restrict__ pos, double* __restrict__ lnp, int ndim, long long half0,
                                    long long nrows, const double* __restrict__ in) {
    const int nd1 = ndim + 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * nd1) return;
    const long long r = i / nd1;
    const int d = (int)(i - r * nd1);
    if (d < ndim) pos[(half0 + r) * ndim + d] = in[i];
    else lnp[half0 + r] = in[i];
}

}  // namespace nmb
