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

namespace nmb {

struct EvalWs {
    double* partials;    // [CW][ntiles]   out-of-transit chi^2 per (walker, tile)      (brute)
    double* oot;         // [CW]           out-of-transit chi^2 per walker
    WalkerConsts* wcs;   // [CW]
    int* rlo;            // [CW]  first in-transit timestamp   (INT_MAX when idle)
    int* rhi;            // [CW]  one past the last            (0 when idle)
    int* blkstart;       // [CW+1] exclusive scan of items per walker
    int* itemwalker;     // [CW*maxch]
    double* itempart;    // [CW*maxch]
    int* itemsdone;      // [CW]  (0 when idle)
    int* grouptick;      // [C*ngroups] (0 when idle)
    int* counters;       // [0] groups/blocks finished (0 when idle), [1] total items of this launch
    int maxch;
};

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
    }
}

template <int S, int THREADS, int TPT>
struct ClassifySmem {
    static constexpr int TT = THREADS * TPT;
    double t[TT], f[TT], w[TT];
    double warpsum[kWTMax][THREADS / 32];
    double off[kMaxS];
    WalkerConsts wc[kWTMax];
    unsigned long long mbar;
    int s_scan[THREADS / 32 + 1];
    int last;
};

template <int S, int THREADS, int TPT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
sweep_classify_kernel(LcView lc, const double* __restrict__ params, int W, int WT, const double* __restrict__ priors,
                      EvalWs ws, int item_ts, double* __restrict__ lnprob, double* __restrict__ loglike) {
    constexpr int TT = THREADS * TPT;
    constexpr int NWARP = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClassifySmem<S, THREADS, TPT>& sm = *reinterpret_cast<ClassifySmem<S, THREADS, TPT>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, group = blockIdx.y, cand = blockIdx.z;
    const int ngroups = gridDim.y, ntiles = gridDim.x;
    const int Sx = s_of<S>(lc.S);
    const long long n = lc.n[cand];
    const long long base = (long long)tile * TT;
    // walkers are dealt to groups round-robin so that neighbouring (similar) walkers do not pile
    // up in one block: walker = wl * ngroups + group
    const int nw = (W - group + ngroups - 1) / ngroups < WT ? (W - group + ngroups - 1) / ngroups : WT;

    if (base < n) {
        if (tid == 0) {
            mbar_init(&sm.mbar, 1);
            mbar_expect_tx(&sm.mbar, 3u * TT * sizeof(double));
            const long long g = (long long)cand * lc.npad + base;
            tma_bulk_g2s(sm.t, lc.t + g, TT * sizeof(double), &sm.mbar);
            tma_bulk_g2s(sm.f, lc.f + g, TT * sizeof(double), &sm.mbar);
            tma_bulk_g2s(sm.w, lc.w + g, TT * sizeof(double), &sm.mbar);
        }
        if (tid < Sx) sm.off[tid] = lc.off[(long long)cand * Sx + tid];
        if (tid < nw) {
            const long long cw = (long long)cand * W + tid * ngroups + group;
            sm.wc[tid] = make_walker_consts(params + cw * 6, Sx);
            if (tile == 0) ws.wcs[cw] = sm.wc[tid];
        }
        __syncthreads();
        mbar_wait(&sm.mbar, 0);

        double ti[TPT], fi[TPT], wi[TPT];
        bool valid[TPT];
#pragma unroll
        for (int s = 0; s < TPT; ++s) {
            const int li = s * THREADS + tid;
            valid[s] = (base + li) < n;
            ti[s] = sm.t[li];
            fi[s] = sm.f[li];
            wi[s] = sm.w[li];
        }
        double off[S > 0 ? S : 1];
        if (S > 0) {
#pragma unroll
            for (int k = 0; k < S; ++k) off[k] = sm.off[k];
        }

        for (int wl = 0; wl < nw; ++wl) {
            const double tcen = sm.wc[wl].tcen, vel = sm.wc[wl].vel, b2 = sm.wc[wl].b2, qthr = sm.wc[wl].qthr;
            const double fbar = sm.wc[wl].fbar;
            double acc = 0.0;
            int hmin = INT_MAX, hmax = -1;
#pragma unroll
            for (int s = 0; s < TPT; ++s) {
                bool hit = false;
                if (S > 0) {
#pragma unroll
                    for (int k = 0; k < S; ++k) {
                        const double vx = vel * ((ti[s] + off[k]) - tcen);
                        hit |= !(fma(vx, vx, b2) > qthr);
                    }
                } else {
                    for (int k = 0; k < Sx; ++k) {
                        const double vx = vel * ((ti[s] + sm.off[k]) - tcen);
                        hit |= !(fma(vx, vx, b2) > qthr);
                    }
                }
                if (valid[s]) {
                    if (hit) {
                        const int li = s * THREADS + tid;
                        hmin = min(hmin, li);
                        hmax = max(hmax, li);
                    } else {
                        const double r = fi[s] - fbar;
                        acc += wi[s] * (r * r);
                    }
                }
            }
            if (__any_sync(0xffffffffu, hmax >= 0)) {  // rare: this warp touches the transit
                const int wmin = __reduce_min_sync(0xffffffffu, hmin);
                const int wmax = __reduce_max_sync(0xffffffffu, hmax);
                if (lane == 0) {
                    const long long cw = (long long)cand * W + wl * ngroups + group;
                    atomicMin(ws.rlo + cw, (int)base + wmin);
                    atomicMax(ws.rhi + cw, (int)base + wmax + 1);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sm.warpsum[wl][warp] = acc;
        }
        __syncthreads();
        if (tid < nw) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < NWARP; ++i) s += sm.warpsum[tid][i];
            ws.partials[((long long)cand * W + tid * ngroups + group) * ntiles + tile] = s;
        }
    } else {
        if (tid < nw) ws.partials[((long long)cand * W + tid * ngroups + group) * ntiles + tile] = 0.0;
    }

    // ---- the last tile-block of this (candidate, group) folds the tiles of its walkers ----------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int* tk = ws.grouptick + (long long)cand * ngroups + group;
        const int t = atomicAdd(tk, 1);
        sm.last = (t == ntiles - 1);
        if (sm.last) *tk = 0;
    }
    __syncthreads();
    if (!sm.last) return;
    __threadfence();
    for (int wl = warp; wl < nw; wl += NWARP) {
        const long long cw = (long long)cand * W + wl * ngroups + group;
        const double* row = ws.partials + cw * ntiles;
        double s = 0.0;
        for (int i = lane; i < ntiles; i += 32) s += __ldcg(row + i);
        s = warp_sum(s);
        if (lane == 0) {
            const int n_w = __ldcg(ws.rhi + cw) - __ldcg(ws.rlo + cw);
            if (n_w <= 0) {  // never in transit: done here
                if (loglike) loglike[cw] = s;
                lnprob[cw] = s + (priors ? ln_prior_dev(params + cw * 6, priors + (long long)cand * 36, 6) : 0.0);
                ws.rlo[cw] = INT_MAX;
                ws.rhi[cw] = 0;
            } else {
                ws.oot[cw] = s;
            }
        }
    }
    // ---- the globally last group builds the item list for stage 2 --------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int done = atomicAdd(ws.counters, 1);
        sm.last = (done == (int)(gridDim.y * gridDim.z) - 1);
    }
    __syncthreads();
    if (!sm.last) return;
    __threadfence();
    scan_items<THREADS>(ws, W * (int)gridDim.z, item_ts, sm.s_scan);
}

// ---- windowed stage 1 -----------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
window_ranges_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors,
                     EvalWs ws, int item_ts, double* __restrict__ lnprob, double* __restrict__ loglike) {
    __shared__ int s_scan[THREADS / 32 + 1];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CW = W * lc.ncand;
    const int cw = blockIdx.x * (THREADS / 32) + warp;
    const int Sx = lc.S;
    if (cw < CW) {
        const int cand = cw / W;
        const long long n = lc.n[cand];
        const double* __restrict__ gt = lc.t + (long long)cand * lc.npad;
        const WalkerConsts wc = make_walker_consts(params + (long long)cw * 6, Sx);
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
        long long ilo, ihi;
        if (full) {
            ilo = 0; ihi = n;
        } else if (empty) {
            ilo = 0; ihi = 0;
        } else {
            ilo = max(0LL, warp_bound<false>(gt, n, tlo, lane) - 1);
            ihi = min(n, warp_bound<true>(gt, n, thi, lane) + 1);
            if (ihi < ilo) ihi = ilo;
        }
        if (lane == 0) {
            // out-of-transit part: pre[ilo] + (pre[n] - pre[ihi]) in double-double
            const double* ph = lc.pre_hi + (long long)cand * (lc.npad + 1);
            const double* pl = lc.pre_lo + (long long)cand * (lc.npad + 1);
            double s, e, s2, e2;
            two_sum(ph[n], -ph[ihi], s, e);
            e += pl[n] - pl[ihi];
            two_sum(s, ph[ilo], s2, e2);
            e2 += e + pl[ilo];
            const double oot = s2 + e2;
            if (ihi - ilo <= 0) {
                if (loglike) loglike[cw] = oot;
                lnprob[cw] = oot + (priors ? ln_prior_dev(params + (long long)cw * 6, priors + (long long)cand * 36, 6) : 0.0);
                ws.rlo[cw] = INT_MAX;
                ws.rhi[cw] = 0;
            } else {
                ws.oot[cw] = oot;
                ws.wcs[cw] = wc;
                ws.rlo[cw] = (int)ilo;
                ws.rhi[cw] = (int)ihi;
            }
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

template <bool FAST>
__device__ __forceinline__ double sample_flux_eval(double ti, double offk, const WalkerConsts& wc);

__device__ __noinline__ double sample_flux_slowpath(double z, const QuadConsts* q) {
    return occultquad_point<false, IeeeOps>(z, *q, nullptr, nullptr, nullptr);
}

template <>
__device__ __forceinline__ double sample_flux_eval<true>(double ti, double offk, const WalkerConsts& wc) {
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
__device__ __forceinline__ double sample_flux_eval<false>(double ti, double offk, const WalkerConsts& wc) {
    return sample_flux(ti, offk, wc);
}

template <int S, int THREADS>
struct EvalSmem {
    double scratch[THREADS / 32][kMaxTSW * ((S > 0 ? S : kMaxS) | 1)];
};

template <int S, int THREADS, int MINB, bool FAST>
__global__ void __launch_bounds__(THREADS, MINB)
eval_ranges_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors, EvalWs ws,
                   int TSW, double* __restrict__ lnprob, double* __restrict__ loglike) {
    constexpr int NWARP = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalSmem<S, THREADS>& sm = *reinterpret_cast<EvalSmem<S, THREADS>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Sx = s_of<S>(lc.S);
    const int SP = Sx | 1;
    const int total = ws.counters[1];
    double* scratch = sm.scratch[warp];
    const int gwarp = blockIdx.x * NWARP + warp, nwarps = gridDim.x * NWARP;

    for (int item = gwarp; item < total; item += nwarps) {
        const int cw = __ldg(ws.itemwalker + item);
        const int cand = cw / W;
        const int first = __ldcg(ws.blkstart + cw);
        const int nch = __ldcg(ws.blkstart + cw + 1) - first;
        const int lo = __ldcg(ws.rlo + cw), hi = __ldcg(ws.rhi + cw);
        const WalkerConsts wc = ws.wcs[cw];
        const int ci = item - first;
        const int n_w = hi - lo;
        int clen = (n_w + nch - 1) / nch;
        clen = (clen + TSW - 1) / TSW * TSW;
        const int cstart = lo + ci * clen;
        const int cend = min(hi, cstart + clen);
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
                scratch[h * SP + k] = sample_flux_eval<FAST>(__ldg(gt + sub + h), __ldg(goff + k), wc);
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
        int done = 0;
        if (lane == 0) {
            ws.itempart[item] = acc;
            __threadfence();
            done = atomicAdd(ws.itemsdone + cw, 1);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == nch - 1) {  // this warp completed the walker: fold its items, add the priors
            __threadfence();
            double s = 0.0;
            for (int c = lane; c < nch; c += 32) s += __ldcg(ws.itempart + first + c);
            s = warp_sum(s);
            if (lane == 0) {
                const double ll = ws.oot[cw] + s;
                if (loglike) loglike[cw] = ll;
                lnprob[cw] = ll + (priors ? ln_prior_dev(params + (long long)cw * 6, priors + (long long)cand * 36, 6) : 0.0);
                ws.itemsdone[cw] = 0;
                ws.rlo[cw] = INT_MAX;
                ws.rhi[cw] = 0;
            }
        }
    }
}

__global__ void init_ranges_kernel(int* rlo, int* rhi, int* itemsdone, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { rlo[i] = INT_MAX; rhi[i] = 0; itemsdone[i] = 0; }
}

}  // namespace nmb
