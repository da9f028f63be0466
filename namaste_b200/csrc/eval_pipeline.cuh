// Two-stage ensemble log-probability pipeline (FP64, sm_100a).
//
// Stage 1 finds, per walker, the contiguous index range of timestamps that can be in transit
// and the chi^2 of everything outside it; stage 2 evaluates those ranges with the full
// Mandel-Agol model, load-balanced across the whole GPU; the last item of a walker adds the
// priors and writes lnprob.  Both stages are deterministic (fixed summation trees, integer
// atomics only), so a chain is reproducible bit for bit.
//
//   sweep_classify_kernel (NMB_MODE_BRUTE)   : visits every supersample of every timestamp:
//       grid = (time tile, walker group, candidate); the tile's (t, flux, w) are staged in shared
//       memory by TMA bulk copies and reused by all walkers of the group; 5 FP64 ops per sample
//       (t+off, -tcen, *vel, fma, compare) decide "outside the limb"; out-of-transit chi^2 terms
//       are accumulated on the fly; in-transit timestamps only widen the walker's [lo, hi) range.
//   window_ranges_kernel  (NMB_MODE_WINDOWED): a warp per walker derives a conservative range from
//       the geometry with a 32-ary search over t, and takes the out-of-transit chi^2 from the
//       double-double prefix sums staged with the light curve.
//   eval_ranges_kernel    (both)             : persistent blocks pull (walker, chunk) items; one
//       thread per supersample, bins reduced in NumPy's order, residuals accumulated per item.
//
// Why a range is enough: along the sorted supersample sequence fl(t_i + off_k) - tcen is monotone,
// so fma(vx, vx, b^2) is V-shaped and {q <= threshold} is contiguous.
//
// Reference path: namaste/namaste.py:1155-1164, :1347, :1314-1332; Crossfield_transit.py:690-970.
#pragma once
#include <limits.h>

#include "lnprob_kernels.cuh"
#include "sampler.cuh"

namespace nmb {

struct EvalWs {
    double* partials;    // [CW][ntiles]   out-of-transit chi^2 per (walker, tile)      (brute)
    double* oot;         // [CW]           out-of-transit chi^2 per walker
    double* lp;          // [CW]           log-prior per walker (computed in stage 1)
    WalkerConsts* wcs;   // [CW]
    int* rlo;            // [CW]  first in-transit timestamp   (INT_MAX when idle)
    int* rhi;            // [CW]  one past the last            (0 when idle)
    int* blkstart;       // [CW+1] exclusive scan of items per walker
    int* itemwalker;     // [CW*maxch]
    double* itempart;    // [CW*maxch]
    int* itemsdone;      // [CW]  (0 when idle)
    int* grouptick;      // [C*ngroups] (0 when idle)
    int* counters;       // [0] groups/blocks finished (0 when idle), [1] total items of this launch,
                         // [2] next item of the stage-2 queue
    int* tilectr;        // [C*ntiles] next walker group of each tile (reset by stage 2)
    int ntilectr;
    int maxch;
};

// Every evaluated walker ends here exactly once: plain evaluation writes the outputs, sampler mode
// runs the accept/reject step of the stretch move instead.
__device__ __forceinline__ void finish_walker(const SampleCtx& sc, long long cw, int W, double ll, double lp,
                                              double* __restrict__ lnprob, double* __restrict__ loglike) {
    if (loglike) loglike[cw] = ll;
    if (lnprob) lnprob[cw] = ll + lp;
    if (sc.enabled) accept_stretch(sc, (int)(cw / W), (int)(cw % W), W, ll + lp);
}

template <int S>
__host__ __device__ constexpr int ts_per_block(int threads, int s_rt) {
    return threads / (S > 0 ? S : s_rt);
}

__device__ __forceinline__ int items_for(int n_w, int tsb, int maxch) {
    if (n_w <= 0) return 0;
    const int need = (n_w + tsb - 1) / tsb;
    return need < maxch ? need : maxch;
}

// Exclusive scan over all CW walkers by ONE block (the last one of stage 1); fills blkstart,
// itemwalker and counters[1].  Tiles of THREADS*8 walkers; warp-shuffle scans, no serial part.
template <int THREADS>
__device__ void scan_items(const EvalWs& ws, int CW, int item_ts, int* s_scan /*[THREADS/32 + 1]*/) {
    constexpr int PER = 8, NWARP = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int running = 0;
    for (int base = 0; base < CW; base += THREADS * PER) {
        const int i0 = base + tid * PER;
        int cnt[PER];
        int local = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = i0 + j;
            cnt[j] = (i < CW) ? items_for(__ldcg(ws.rhi + i) - __ldcg(ws.rlo + i), item_ts, ws.maxch) : 0;
            local += cnt[j];
        }
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int woff = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) {
            const int v = s_scan[w];
            if (w < warp) woff += v;
            tile_total += v;
        }
        int run = running + woff + incl - local;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = i0 + j;
            if (i < CW) {
                ws.blkstart[i] = run;
                for (int c = 0; c < cnt[j]; ++c) ws.itemwalker[run + c] = i;
                run += cnt[j];
            }
        }
        running += tile_total;
        __syncthreads();
    }
    if (tid == 0) {
        ws.blkstart[CW] = running;
        ws.counters[1] = running;
        ws.counters[0] = 0;
        ws.counters[2] = 0;
    }
}

// named barriers (id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int S, int CTHREADS, int TPT, int WT>
struct ClassifySmem {
    static constexpr int TT = CTHREADS * TPT;
    double t[TT], f[TT], w[TT];
    double warpsum[2][WT][CTHREADS / 32];
    double off[kMaxS];
    // per-walker classification constants, structure-of-arrays, double buffered across groups
    double vel[2][WT], nvtc[2][WT], b2[2][WT], fbar[2][WT];
    int qhi[2][WT];
    unsigned long long mbar;
    int group[2];
    int s_scan[CTHREADS / 32 + 2];
    int doscan;
};

// Brute-force stage 1, warp specialised.  A block owns one time tile (staged once by TMA) and
// pulls walker groups from the tile's queue.
//   producer warp (the last one): fetches the next group, derives its walkers' constants one
//     group ahead of the consumers, later publishes the group's tile partials, and - when this
//     was the group's last tile - folds the tiles, evaluates the priors and settles walkers that
//     never transit;
//   consumer warps: per thread and timestamp the S fine times ft_k = t_i + off_k are formed once
//     per group and shared by its WT walkers; per walker and fine sample the sweep costs two DFMA
//     and one DSETP:  vx' = fma(vel, ft_k, -vel*tcen), q' = fma(vx', vx', b^2), outside iff
//     q' > qthr_c.  They never wait for constants or reductions of other groups.
template <int S, int CTHREADS, int TPT, int WT, int MINB>
__global__ void __launch_bounds__(CTHREADS + 32, MINB)
sweep_classify_kernel(LcView lc, const double* __restrict__ params, int W, int ngroups,
                      const double* __restrict__ priors, EvalWs ws, int item_ts, double* __restrict__ lnprob,
                      double* __restrict__ loglike, SampleCtx sc) {
    constexpr int TT = CTHREADS * TPT;
    constexpr int NCW = CTHREADS / 32;  // consumer warps
    constexpr int ALL = CTHREADS + 32;
    constexpr int SR = S > 0 ? S : kMaxS;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3;  // +buf
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClassifySmem<S, CTHREADS, TPT, WT>& sm = *reinterpret_cast<ClassifySmem<S, CTHREADS, TPT, WT>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, cand = blockIdx.z;
    const int ntiles = gridDim.x;
    const int Sx = s_of<S>(lc.S);
    const long long n = lc.n[cand];
    const long long base = (long long)tile * TT;
    const bool live = base < n;  // tiles past this candidate's end only contribute zeros
    int* ctr = ws.tilectr + (long long)cand * ntiles + tile;

    if (tid == 0) {
        sm.doscan = 0;
        if (live) {
            mbar_init(&sm.mbar, 1);
            mbar_expect_tx(&sm.mbar, 3u * TT * sizeof(double));
            const long long g = (long long)cand * lc.npad + base;
            tma_bulk_g2s(sm.t, lc.t + g, TT * sizeof(double), &sm.mbar);
            tma_bulk_g2s(sm.f, lc.f + g, TT * sizeof(double), &sm.mbar);
            tma_bulk_g2s(sm.w, lc.w + g, TT * sizeof(double), &sm.mbar);
        }
    }
    if (tid < Sx) sm.off[tid] = lc.off[(long long)cand * Sx + tid];
    __syncthreads();

    if (warp == NCW) {
        // ------------------------------- producer ------------------------------------------
        int pend_group[2] = {-1, -1};
        for (int it = 0;; ++it) {
            const int buf = it & 1;
            if (it >= 2) {
                bar_sync(BAR_EMPTY + buf, ALL);  // consumers finished group it-2: its warp sums are in smem
                const int group = pend_group[buf];
                const int nw_all = (W - group + ngroups - 1) / ngroups;
                const int nw = nw_all < WT ? nw_all : WT;
                if (lane < nw) {
                    double sacc = 0.0;
                    if (live) {
#pragma unroll
                        for (int i = 0; i < NCW; ++i) sacc += sm.warpsum[buf][lane][i];
                    }
                    ws.partials[((long long)cand * W + lane * ngroups + group) * ntiles + tile] = sacc;
                }
                __threadfence();
                __syncwarp();
                int tk = 0;
                if (lane == 0) {
                    int* gt = ws.grouptick + (long long)cand * ngroups + group;
                    tk = atomicAdd(gt, 1);
                    if (tk == ntiles - 1) *gt = 0;
                }
                tk = __shfl_sync(0xffffffffu, tk, 0);
                if (tk == ntiles - 1) {  // last tile of this group: fold tiles, priors, settle
                    __threadfence();
                    if (lane < nw) {
                        const long long cw = (long long)cand * W + lane * ngroups + group;
                        const double* row = ws.partials + cw * ntiles;
                        double sacc = 0.0;
                        for (int i = 0; i < ntiles; ++i) sacc += __ldcg(row + i);
                        const int n_w = __ldcg(ws.rhi + cw) - __ldcg(ws.rlo + cw);
                        if (n_w <= 0) {  // never in transit: done here
                            finish_walker(sc, cw, W, sacc, ws.lp[cw], lnprob, loglike);
                            ws.rlo[cw] = INT_MAX;
                            ws.rhi[cw] = 0;
                        } else {
                            ws.oot[cw] = sacc;
                        }
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) {
                        const int done = atomicAdd(ws.counters, 1);
                        if (done == ngroups * (int)gridDim.z - 1) sm.doscan = 1;  // every group folded
                    }
                }
                pend_group[buf] = -1;
            }
            int group = 0;
            if (lane == 0) group = atomicAdd(ctr, 1);
            group = __shfl_sync(0xffffffffu, group, 0);
            const bool end = group >= ngroups;
            if (!end) {
                const int nw_all = (W - group + ngroups - 1) / ngroups;
                const int nw = nw_all < WT ? nw_all : WT;
                if (lane < WT) {
                    if (live && lane < nw) {
                        const WalkerConsts* wc = ws.wcs + ((long long)cand * W + lane * ngroups + group);
                        sm.vel[buf][lane] = wc->vel; sm.nvtc[buf][lane] = wc->nvtc; sm.b2[buf][lane] = wc->b2;
                        sm.qhi[buf][lane] = wc->qhi; sm.fbar[buf][lane] = wc->fbar;
                    } else {  // padding walker: never in transit, result discarded
                        sm.vel[buf][lane] = 0.0; sm.nvtc[buf][lane] = 0.0; sm.b2[buf][lane] = 0.0;
                        sm.qhi[buf][lane] = -1; sm.fbar[buf][lane] = 1.0;
                    }
                }
                pend_group[buf] = group;
            }
            if (lane == 0) sm.group[buf] = end ? -1 : group;
            __syncwarp();
            bar_arrive(BAR_FULL + buf, ALL);
            if (end) {
                // drain: the other buffer may still hold a group the consumers are working on
                const int ob = buf ^ 1;
                if (pend_group[ob] >= 0) {
                    // reuse the loop body by running one more virtual iteration for that buffer
                    bar_sync(BAR_EMPTY + ob, ALL);
                    const int g2 = pend_group[ob];
                    const int nw_all2 = (W - g2 + ngroups - 1) / ngroups;
                    const int nw2 = nw_all2 < WT ? nw_all2 : WT;
                    if (lane < nw2) {
                        double sacc = 0.0;
                        if (live) {
#pragma unroll
                            for (int i = 0; i < NCW; ++i) sacc += sm.warpsum[ob][lane][i];
                        }
                        ws.partials[((long long)cand * W + lane * ngroups + g2) * ntiles + tile] = sacc;
                    }
                    __threadfence();
                    __syncwarp();
                    int tk = 0;
                    if (lane == 0) {
                        int* gt = ws.grouptick + (long long)cand * ngroups + g2;
                        tk = atomicAdd(gt, 1);
                        if (tk == ntiles - 1) *gt = 0;
                    }
                    tk = __shfl_sync(0xffffffffu, tk, 0);
                    if (tk == ntiles - 1) {
                        __threadfence();
                        if (lane < nw2) {
                            const long long cw = (long long)cand * W + lane * ngroups + g2;
                            const double* row = ws.partials + cw * ntiles;
                            double sacc = 0.0;
                            for (int i = 0; i < ntiles; ++i) sacc += __ldcg(row + i);
                            const int n_w = __ldcg(ws.rhi + cw) - __ldcg(ws.rlo + cw);
                            if (n_w <= 0) {
                                finish_walker(sc, cw, W, sacc, ws.lp[cw], lnprob, loglike);
                                ws.rlo[cw] = INT_MAX;
                                ws.rhi[cw] = 0;
                            } else {
                                ws.oot[cw] = sacc;
                            }
                        }
                        __threadfence();
                        __syncwarp();
                        if (lane == 0) {
                            const int done = atomicAdd(ws.counters, 1);
                            if (done == ngroups * (int)gridDim.z - 1) sm.doscan = 1;
                        }
                    }
                }
                break;
            }
        }
    } else {
        // ------------------------------- consumers -----------------------------------------
        if (live) mbar_wait(&sm.mbar, 0);
        for (int it = 0;; ++it) {
            const int buf = it & 1;
            bar_sync(BAR_FULL + buf, ALL);
            const int group = sm.group[buf];
            if (group < 0) break;
            if (live) {
                double acc[WT];
#pragma unroll
                for (int wl = 0; wl < WT; ++wl) acc[wl] = 0.0;
#pragma unroll 1
                for (int s = 0; s < TPT; ++s) {
                    const int li = s * CTHREADS + tid;
                    const bool valid = (base + li) < n;
                    const double ti = sm.t[li], fi = sm.f[li], wi = sm.w[li];
                    const double r1 = fi - 1.0;
                    const double term1 = wi * (r1 * r1);
                    double ft[SR];
                    if (S > 0) {
#pragma unroll
                        for (int k = 0; k < SR; ++k) ft[k] = ti + sm.off[k];
                    }
#pragma unroll
                    for (int wl = 0; wl < WT; ++wl) {
                        const double vel = sm.vel[buf][wl], nvtc = sm.nvtc[buf][wl], b2 = sm.b2[buf][wl];
                        const int qhi = sm.qhi[buf][wl];
                        int qmin = INT_MAX;  // smallest high word of q' over the S fine samples
                        if (S > 0) {
#pragma unroll
                            for (int k = 0; k < SR; ++k) {
                                const double vx = fma(vel, ft[k], nvtc);
                                qmin = min(qmin, __double2hiint(fma(vx, vx, b2)));
                            }
                        } else {
                            for (int k = 0; k < Sx; ++k) {
                                const double vx = fma(vel, ti + sm.off[k], nvtc);
                                qmin = min(qmin, __double2hiint(fma(vx, vx, b2)));
                            }
                        }
                        const bool hit = (qmin <= qhi) && valid;
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (m != 0 && lane == 0) {  // rare: this warp touches the transit of walker wl
                            const long long cw = (long long)cand * W + wl * ngroups + group;
                            const int w0i = (int)base + s * CTHREADS + warp * 32;
                            atomicMin(ws.rlo + cw, w0i + (__ffs(m) - 1));
                            atomicMax(ws.rhi + cw, w0i + (32 - __clz(m)));
                        }
                        const double fbar = sm.fbar[buf][wl];
                        double term = term1;
                        if (fbar != 1.0) {  // block-uniform, practically never taken
                            const double r = fi - fbar;
                            term = wi * (r * r);
                        }
                        acc[wl] += (valid && !hit) ? term : 0.0;
                    }
                }
#pragma unroll
                for (int wl = 0; wl < WT; ++wl) {
                    const double v = warp_sum(acc[wl]);
                    if (lane == 0) sm.warpsum[buf][wl][warp] = v;
                }
            }
            __threadfence();  // range atomics of this group are visible before the producer's ticket
            __syncwarp();
            bar_arrive(BAR_EMPTY + buf, ALL);
        }
    }
    __syncthreads();
    if (!sm.doscan) return;
    __threadfence();
    scan_items<ALL>(ws, W * (int)gridDim.z, item_ts, sm.s_scan);
}

// ---- per-walker setup for the brute-force sweep: constants + log-prior, one thread per walker ----
__global__ void walker_setup_kernel(LcView lc, const double* __restrict__ params, int W,
                                    const double* __restrict__ priors, EvalWs ws, SampleCtx sc) {
    const int cw = blockIdx.x * blockDim.x + threadIdx.x;
    if (cw >= W * lc.ncand) return;
    const int cand = cw / W;
    const double* par = params;
    if (sc.enabled) {
        propose_stretch(sc, cand, cw % W, W);
        par = sc.q;  // the proposal just written by this thread (not through the read-only path)
    }
    ws.wcs[cw] = make_walker_consts(par + (long long)cw * 6, lc.S);
    ws.lp[cw] = priors ? ln_prior_dev(par + (long long)cw * 6, priors + (long long)cand * 36, 6) : 0.0;
}

// ---- windowed stage 1: one thread per walker ---------------------------------------------------
// first index i in [0, n) with t[i] >= x (UPPER = false) or t[i] > x (UPPER = true)
template <bool UPPER>
__device__ __forceinline__ int bsearch_t(const double* __restrict__ t, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double v = __ldg(t + mid);
        if (UPPER ? (v > x) : (v >= x)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
window_ranges_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors,
                     EvalWs ws, int item_ts, double* __restrict__ lnprob, double* __restrict__ loglike,
                     SampleCtx sc) {
    __shared__ int s_scan[THREADS / 32 + 1];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int CW = W * lc.ncand;
    const int cw = blockIdx.x * THREADS + tid;
    const int Sx = lc.S;
    if (cw < CW) {
        const int cand = cw / W;
        const int n = (int)lc.n[cand];
        const double* par = params;
        if (sc.enabled) {
            propose_stretch(sc, cand, cw % W, W);
            par = sc.q;  // the proposal just written by this thread (not through the read-only path)
        }
        const WalkerConsts wc = make_walker_consts(par + (long long)cw * 6, Sx);
        const double span = lc.off[(long long)cand * Sx + Sx - 1];
        const double hx2 = wc.qthr - wc.b2;
        bool full = !(wc.fbar == 1.0);  // prefix sums assume the out-of-transit model is exactly 1
        bool empty = false;
        double tlo = 0.0, thi = 0.0;
        if (!full) {
            if (hx2 < 0.0) {
                empty = true;  // |b| beyond the limb: never in transit
            } else if (!(hx2 >= 0.0) || !(fabs(wc.vel) > 0.0)) {
                full = true;  // NaN somewhere, or vel == 0 (z == |b| for every sample)
            } else {
                const double ht = sqrt(hx2) / fabs(wc.vel) * (1.0 + 1e-9);
                const double slack = 1e-9 * (fabs(wc.tcen) + 1.0) + fabs(span) * 1e-6;
                tlo = wc.tcen - ht - span - slack;
                thi = wc.tcen + ht + slack;
                if (!(tlo == tlo) || !(thi == thi) || isinf(tlo) || isinf(thi)) full = true;
            }
        }
        int ilo, ihi;
        if (full) {
            ilo = 0; ihi = n;
        } else if (empty) {
            ilo = 0; ihi = 0;
        } else {
            // O(1) lookup in the uniform time grid, widened by one cell on each side
            const double t0 = lc.t0[cand], ginv = lc.ginv[cand];
            const int gc = lc.gcount[cand];
            const int* __restrict__ grid = lc.tgrid + (long long)cand * lc.gstride;
            const double jl = floor((tlo - t0) * ginv) - 1.0, jh = floor((thi - t0) * ginv) + 2.0;
            const int j0 = jl < 0.0 ? 0 : (jl > (double)gc ? gc : (int)jl);
            const int j1 = jh < 0.0 ? 0 : (jh > (double)gc ? gc : (int)jh);
            ilo = __ldg(grid + j0);
            ihi = (jh > (double)gc) ? n : __ldg(grid + j1);
            if (ihi < ilo) ihi = ilo;
        }
        // out-of-transit part: pre[ilo] + (pre[n] - pre[ihi]) in double-double
        const double* ph = lc.pre_hi + (long long)cand * (lc.npad + 1);
        const double* pl = lc.pre_lo + (long long)cand * (lc.npad + 1);
        double sa, ea, sb, eb;
        two_sum(ph[n], -ph[ihi], sa, ea);
        ea += pl[n] - pl[ihi];
        two_sum(sa, ph[ilo], sb, eb);
        eb += ea + pl[ilo];
        const double oot = sb + eb;
        const double lp = priors ? ln_prior_dev(par + (long long)cw * 6, priors + (long long)cand * 36, 6) : 0.0;
        if (ihi - ilo <= 0) {
            finish_walker(sc, cw, W, oot, lp, lnprob, loglike);
            ws.rlo[cw] = INT_MAX;
            ws.rhi[cw] = 0;
        } else {
            ws.oot[cw] = oot;
            ws.lp[cw] = lp;
            ws.wcs[cw] = wc;
            ws.rlo[cw] = ilo;
            ws.rhi[cw] = ihi;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int done = atomicAdd(ws.counters, 1);
        s_last = (done == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    scan_items<THREADS>(ws, CW, item_ts, s_scan);
}

// ---- stage 2: evaluate the in-transit ranges -------------------------------------------------
// Each WARP pulls (walker, chunk) items on its own (no block barriers): it walks the chunk in
// groups of TSW timestamps, one supersample per lane and round, parks the fluxes in its private
// shared-memory scratch, lets the first TSW lanes reduce the bins and accumulate residuals, and
// publishes the item's partial sum; the warp that completes a walker's last item folds the items
// (fixed order), adds the priors and writes lnprob.
constexpr int kMaxTSW = 32;

// PATH: 0 = general IEEE FP64, 1 = FP64 hot path (default), 2 = FP32 hot path (NMB_F32)
template <int PATH>
__device__ __forceinline__ double sample_flux_eval(double ti, double offk, const WalkerConsts& wc,
                                                   const QuadConstsF& cf);

__device__ __noinline__ double sample_flux_slowpath(double z, const QuadConsts* q) {
    return occultquad_point<false, IeeeOps>(z, *q, nullptr, nullptr, nullptr);
}

template <>
__device__ __forceinline__ double sample_flux_eval<1>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF&) {
    const double ft = ti + offk;
    const double x = ft - wc.tcen;
    const double vx = wc.vel * x;
    const double q = wc.b2 + vx * vx;
    const double z = FastOps::sqrt(q);
    if (z > wc.q.onep) return wc.q.fout;
    double F;
    // hot cases with branch-free Newton-Raphson quotients/roots; anything else (other cases,
    // operands outside the normal range -> non-finite z or F) takes the general IEEE path
    if (!occultquad_hot(z, wc.q, F) || !(fabs(F) <= 1e300)) {
        const double zz = ::sqrt(q);
        F = (zz > wc.q.onep) ? wc.q.fout : sample_flux_slowpath(zz, &wc.q);
    }
    return F;
}
template <>
__device__ __forceinline__ double sample_flux_eval<0>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF&) {
    return sample_flux(ti, offk, wc);
}
template <>
__device__ __forceinline__ double sample_flux_eval<2>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF& cf) {
    const double ft = ti + offk;
    const double x = ft - wc.tcen;          // FP64 difference of the absolute times, then FP32
    const float vx = (float)wc.vel * (float)x;
    const float zf = sqrtf(fmaf(vx, vx, (float)wc.b2));
    if (zf > cf.onep * 1.000001f) return wc.q.fout;
    float depth;
    if (occultquad_hot_f32(zf, cf, depth) && fabsf(depth) <= 1e30f) return 1.0 - (double)depth;
    // anything else (other cases, exact limb, non-finite): the FP64 general path
    const double vxd = wc.vel * x;
    const double zz = ::sqrt(wc.b2 + vxd * vxd);
    return (zz > wc.q.onep) ? wc.q.fout : sample_flux_slowpath(zz, &wc.q);
}

template <int S, int THREADS>
struct EvalSmem {
    double scratch[THREADS / 32][kMaxTSW * ((S > 0 ? S : kMaxS) | 1)];
};

struct ItemMeta {
    int cw, first, nch, lo, hi;
};
__device__ __forceinline__ ItemMeta load_item_meta(const EvalWs& ws, int item, int total) {
    ItemMeta m;
    m.cw = -1; m.first = 0; m.nch = 1; m.lo = 0; m.hi = 0;
    if (item < total) {
        m.cw = __ldg(ws.itemwalker + item);
        m.first = __ldcg(ws.blkstart + m.cw);
        m.nch = __ldcg(ws.blkstart + m.cw + 1) - m.first;
        m.lo = __ldcg(ws.rlo + m.cw);
        m.hi = __ldcg(ws.rhi + m.cw);
    }
    return m;
}

template <int S, int THREADS, int MINB, int PATH>
__global__ void __launch_bounds__(THREADS, MINB)
eval_ranges_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors, EvalWs ws,
                   int TSW, double* __restrict__ lnprob, double* __restrict__ loglike, SampleCtx sc) {
    constexpr int NWARP = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalSmem<S, THREADS>& sm = *reinterpret_cast<EvalSmem<S, THREADS>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Sx = s_of<S>(lc.S);
    const int SP = Sx | 1;
    const int total = ws.counters[1];
    double* scratch = sm.scratch[warp];
    // stage 1 is complete: rearm its per-tile group queues for the next launch
    for (int i = blockIdx.x * THREADS + threadIdx.x; i < ws.ntilectr; i += gridDim.x * THREADS) ws.tilectr[i] = 0;

    // dynamic item queue, software pipelined: the index two items ahead is being fetched and the
    // metadata of the next item is in flight while the current item is evaluated
    int* queue = ws.counters + 2;
    int idx_cur = 0, idx_nxt = 0;
    if (lane == 0) {
        idx_cur = atomicAdd(queue, 1);
        idx_nxt = atomicAdd(queue, 1);
    }
    idx_cur = __shfl_sync(0xffffffffu, idx_cur, 0);
    idx_nxt = __shfl_sync(0xffffffffu, idx_nxt, 0);
    ItemMeta nxt = load_item_meta(ws, idx_cur, total);
    while (idx_cur < total) {
        const int item = idx_cur;
        const ItemMeta cur = nxt;
        int idx_nn = 0;
        if (lane == 0) idx_nn = atomicAdd(queue, 1);
        nxt = load_item_meta(ws, idx_nxt, total);
        const int cw = cur.cw;
        const int cand = cw / W;
        const WalkerConsts wc = ws.wcs[cw];
        const QuadConstsF cf = (PATH == 2) ? to_f32(wc.q) : QuadConstsF{};
        const int ci = item - cur.first;
        const int n_w = cur.hi - cur.lo;
        int clen = (n_w + cur.nch - 1) / cur.nch;
        clen = (clen + TSW - 1) / TSW * TSW;
        const int cstart = cur.lo + ci * clen;
        const int cend = min(cur.hi, cstart + clen);
        const double* __restrict__ gt = lc.t + (long long)cand * lc.npad;
        const double* __restrict__ gf = lc.f + (long long)cand * lc.npad;
        const double* __restrict__ gw = lc.w + (long long)cand * lc.npad;
        const double* __restrict__ goff = lc.off + (long long)cand * Sx;

        double acc = 0.0;
        for (int sub = cstart; sub < cend; sub += TSW) {
            const int nts = min(TSW, cend - sub);
            const int nsamp = nts * Sx;
            for (int j = lane; j < nsamp; j += 32) {
                const int h = j / Sx, k = j - h * Sx;
                scratch[h * SP + k] = sample_flux_eval<PATH>(__ldg(gt + sub + h), __ldg(goff + k), wc, cf);
            }
            __syncwarp();
            if (lane < nts) {
                const double m = (S > 0) ? bin_mean<(S > 0 ? S : 1)>(scratch + lane * SP)
                                         : bin_mean_rt(scratch + lane * SP, Sx);
                const double r = __ldg(gf + sub + lane) - m;
                acc += __ldg(gw + sub + lane) * (r * r);
            }
            __syncwarp();
        }
        acc = warp_sum(acc);
        bool fold = (cur.nch == 1);
        double s = acc;
        if (!fold) {  // several items per walker: publish, and let the last one fold them in order
            int done = 0;
            if (lane == 0) {
                ws.itempart[item] = acc;
                __threadfence();
                done = atomicAdd(ws.itemsdone + cw, 1);
            }
            done = __shfl_sync(0xffffffffu, done, 0);
            if (done == cur.nch - 1) {
                __threadfence();
                s = 0.0;
                for (int c = lane; c < cur.nch; c += 32) s += __ldcg(ws.itempart + cur.first + c);
                s = warp_sum(s);
                fold = true;
            }
        }
        if (fold && lane == 0) {
            finish_walker(sc, cw, W, ws.oot[cw] + s, ws.lp[cw], lnprob, loglike);
            ws.itemsdone[cw] = 0;
            ws.rlo[cw] = INT_MAX;
            ws.rhi[cw] = 0;
        }
        idx_cur = idx_nxt;
        idx_nxt = __shfl_sync(0xffffffffu, idx_nn, 0);
    }
}

__global__ void init_ranges_kernel(int* rlo, int* rhi, int* itemsdone, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { rlo[i] = INT_MAX; rhi[i] = 0; itemsdone[i] = 0; }
}

}  // namespace nmb
