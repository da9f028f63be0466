// Two-stage ensemble log-probability pipeline (FP64, sm_100a).
//
// Stage 1 finds, per walker, the contiguous index range of timestamps that can be in transit
// and the chi^2 of everything outside it; stage 2 evaluates those ranges with the full
// Mandel-Agol model, load-balanced across the whole GPU; the last item of a walker adds the
// priors and writes lnprob.  Both stages are deterministic (fixed summation trees, integer
// atomics only), so a chain is reproducible bit for bit.
//
//   sweep_walkers_kernel  (NMB_MODE_BRUTE)   : visits every supersample of every timestamp, one
//       walker per thread; time chunks stream through shared memory by TMA bulk copies; two DFMA
//       and a third of an integer min per sample decide "outside the limb"; out-of-transit chi^2
//       terms are accumulated on the fly; in-transit timestamps only widen the walker's [lo, hi).
//   window_ranges_kernel  (NMB_MODE_WINDOWED): a thread per walker derives a conservative range from
//       the geometry with an O(1) lookup in a uniform time grid, and takes the out-of-transit chi^2
//       from the double-double prefix sums staged with the light curve.
//   eval_slots_kernel     (both)             : resident warps share the queue of slots (passes of
//       one walker's range); one supersample per lane and round, bins reduced in NumPy's order,
//       slot partials stored and folded in slot order by the warp that completes a walker.
//   eval_mixed_kernel     (both, mid-sized ensembles): the same queue workers plus one block per walker
//       that evaluates a short range on the spot from what stage 1 left in the workspace; stage 1 then
//       queues only the long ranges.
//   window_direct_kernel  (NMB_MODE_WINDOWED, ensembles of one resident wave): one block per walker does
//       proposal, constants, window, flux rounds, fold, priors and accept without leaving the SM.
// Stage 1 pushes its walkers' slots with one atomicAdd per block; kernels are chained with
// programmatic dependent launch so the next one is resident when its predecessor drains.  Whatever the
// path, slot boundaries and summation trees are the same, so a walker's lnprob has the same bits.
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

// A *slot* is the unit of work and of summation in stage 2: a run of TSW-timestamp passes of one
// walker's in-transit range (one pass unless the walker needs more than `maxslots` slots).  Slot
// boundaries depend only on the walker's own range and on S, never on the launch shape, and each
// slot's chi^2 partial is stored, so any partition of the slots over warps gives the same bits.
struct SlotRec {
    int cw;      // walker (candidate * W + walker)
    int t0, t1;  // timestamps [t0, t1) of this slot
    int first;   // first slot of the walker (its slots are contiguous in the queue)
};

struct EvalWs {
    double* partials;    // [C][ngran][W]  out-of-transit chi^2 per (granule, walker)   (brute)
    double* oot;         // [CW]           out-of-transit chi^2 per walker
    WalkerConsts* wcs;   // [CW]
    SweepConsts* swc;    // [CW]  classification constants of the brute-force sweep, when sweep_consts_kernel ran first
    int pre_consts;      // 1: the sweep's blocks load swc instead of deriving the walker's constants themselves
    int* rlo;            // [CW]  first in-transit timestamp   (INT_MAX when idle)
    int* rhi;            // [CW]  one past the last            (0 when idle)
    int* nslots;         // [CW]  slots of each walker with work in stage 2
    SlotRec* slots;      // [slotcap]  queue of slots (spans allocated by atomicAdd in stage 1)
    double* slotpart;    // [slotcap]  chi^2 partial of each slot
    int* slotsdone;      // [CW]  (0 when idle)
    int* grouptick;      // [C*walker blocks] finished time blocks per walker group (0 when idle)
    int* counters;       // [5] set when a pushed walker has slots of several passes (non-uniform queue);
                         // [4] direct blocks of window_direct_kernel that have decided (evaluate / finish / queue);
                         // [1] slots pushed by stage 1 (queue tail), [2] next dynamically assigned chunk,
                         // [3] stage-2 blocks finished; the last stage-2 block zeroes them
    int maxslots;        // slots per walker are capped; longer ranges put several passes in a slot
    int tsw;             // timestamps per slot: as many whole bins as fit in kEvalPass supersamples
    int merge;           // consecutive slots of a walker a warp evaluates in one pass (<= kEvalMerge)
    int direct_cap;      // > 0: walkers with at most this many single-pass slots are not queued; a block of
                         // eval_mixed_kernel evaluates each of them on the spot (0: every walker is queued)
    double* modelbuf;    // GP mode only (else null): [CW][npad] binned model of the in-transit timestamps;
                         // stage 2 stores it instead of reducing residuals and nobody finishes walkers
    unsigned long long* trace;  // diagnostics (NMB_TRACE): [16] phase envelopes in globaltimer ns, else null
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// slot 0/8: earliest block start of stage 1/2 (atomicMin); other slots: latest time a phase ended
__device__ __forceinline__ void trace_min(const EvalWs& ws, int slot) {
    if (ws.trace) atomicMin(ws.trace + slot, gtimer());
}
__device__ __forceinline__ void trace_max(const EvalWs& ws, int slot) {
    if (ws.trace) atomicMax(ws.trace + slot, gtimer());
}

// line prefetch into L1 (no register, no dependency): used where an address is known several dependent
// global round trips before its data is needed
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Every evaluated walker ends here exactly once: plain evaluation writes the outputs, sampler mode
// runs the accept/reject step of the stretch move instead.
__device__ __forceinline__ void finish_walker(const SampleCtx& sc, long long cw, int W, double ll, double lp,
                                              double* __restrict__ lnprob, double* __restrict__ loglike) {
    if (loglike) loglike[cw] = ll;
    if (lnprob) lnprob[cw] = ll + lp;
    if (sc.enabled) accept_stretch(sc, (int)(cw / W), (int)(cw % W), W, ll + lp);
}

// number of slots and passes per slot for an in-transit range of n_w timestamps
__device__ __forceinline__ int slots_for(int n_w, int tsw, int maxslots, int& passes_per_slot) {
    passes_per_slot = 1;
    if (n_w <= 0) return 0;
    const int npass = (n_w + tsw - 1) / tsw;
    passes_per_slot = (npass + maxslots - 1) / maxslots;
    return (npass + passes_per_slot - 1) / passes_per_slot;
}

// Block-cooperative push onto the stage-2 queue: every thread brings the in-transit range of its
// walker (empty for none); the block takes one contiguous span of the queue with a single atomicAdd
// and each walker writes its slot records.  The order of spans in the queue depends on block
// timing, results do not (see SlotRec).
template <int THREADS>
__device__ __forceinline__ void push_slots(const EvalWs& ws, long long cw, int lo, int hi, int* s_scan /*[THREADS/32+1]*/) {
    constexpr int NWARP = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int pps;
    int mine = slots_for(hi - lo, ws.tsw, ws.maxslots, pps);
    if (pps == 1 && mine <= ws.direct_cap) mine = 0;  // left to a direct block of eval_mixed_kernel
    if (pps > 1) atomicOr(ws.counters + 5, 1);  // slots of several passes: the queue is not homogeneous (rare)
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
#pragma unroll
        for (int i = 0; i < NWARP; ++i) {
            const int v = s_scan[i];
            s_scan[i] = tot;
            tot += v;
        }
        s_scan[NWARP] = tot ? atomicAdd(ws.counters + 1, tot) : 0;
    }
    __syncthreads();
    const int first = s_scan[NWARP] + s_scan[warp] + incl - mine;
    const int step = pps * ws.tsw;
    auto put = [&](int cwi, int lo_, int hi_, int first_, int step_, int i) {
        SlotRec r;
        r.cw = cwi; r.t0 = lo_ + i * step_; r.t1 = min(hi_, lo_ + (i + 1) * step_); r.first = first_;
        *reinterpret_cast<int4*>(ws.slots + first_ + i) = *reinterpret_cast<const int4*>(&r);
    };
    if (mine > 0) {
        ws.nslots[cw] = mine;
        for (int i = 0; i < min(mine, 8); ++i) put((int)cw, lo, hi, first, step, i);
    }
    // walkers with long ranges (rare): the whole warp writes their remaining records
    unsigned big = __ballot_sync(0xffffffffu, mine > 8);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int cw_b = __shfl_sync(0xffffffffu, (int)cw, src), lo_b = __shfl_sync(0xffffffffu, lo, src);
        const int hi_b = __shfl_sync(0xffffffffu, hi, src), first_b = __shfl_sync(0xffffffffu, first, src);
        const int step_b = __shfl_sync(0xffffffffu, step, src), n_b = __shfl_sync(0xffffffffu, mine, src);
        for (int i = 8 + lane; i < n_b; i += 32) put(cw_b, lo_b, hi_b, first_b, step_b, i);
    }
}

// ---- brute-force stage 1, walker-per-thread --------------------------------------------------
// A block owns WPB*32 walkers (one per thread, constants in registers) and a span of `gpb`
// reduction granules of one light curve.  The span streams through shared memory in chunks of CH
// timestamps: two TMA bulk copies per chunk (t and the precomputed chi^2 terms w*(f-1)^2) into a
// double buffer.  The inner loop costs, per walker and fine sample, ONE DFMA
//   y = fma(V, t_i, c_k),  V = vel s,  c_k = fma(vel, off_k, -vel*tcen) s,  s = 2 / X  (S registers per walker)
// -- the walker's position along the chord in units of half the half-width X of its in-transit window (SweepConsts,
// lnprob_kernels.cuh) -- and half of a 3-input LOP3: a sample can be inside the limb only if |y| < 2, i.e. bit 30 of
// hi32(y) is clear, so "some sample may be in transit" is the AND of the high words.
// (Round 1 and most of round 2 squared the position, q' = fma(vx', vx', b^2), and took the integer minimum of the
// high words against (1 + p)^2: two DFMA and half a VIMNMX3 -- a two-cycle instruction on this part -- per sample.
// The interval test classifies the same samples with half the FP64 work and full-rate logic.)  t_i is one
// broadcast LDS per timestamp.  A DFMA holds the issue port for two cycles on this part and every other instruction
// for one, so the loop is written to keep non-FP64 instructions per sample to a minimum.  (The
// run-time-S instantiation reads a table ft[i][k] = t_i + off_k built per chunk instead.)  A timestamp whose S samples
// all fall outside the window is provably outside the limb (see make_sweep_consts) and adds its
// precomputed term; the others only widen the walker's in-transit range, kept in registers and
// merged with one atomicMin/Max per thread.  Per-(granule, walker) partial sums go to HBM
// (coalesced) and are folded in granule order by the last block of each walker group, which makes
// the result independent of the launch shape.
template <int S, int CH>
struct WalkSmem {
    static constexpr int SR = S > 0 ? S : kMaxS;
    static constexpr int SP = (SR + 1) & ~1;  // even row stride: 16-byte aligned rows for LDS.128
    double ftab[2][S > 0 ? 2 : CH * SP];  // fine-time table: run-time-S instantiation only
    double traw[2][CH];
    double term[2][CH];
    double tsum[2][CH / kSweepTrip];  // the terms of each trip, summed in timestamp order when the light curve was staged
    double off[kMaxS];
    unsigned long long mbar[2];
    int s_scan[40];
    int flag;
};

// AND over the S supersamples of one timestamp of hi32(y), y = fma(V, t_i, ck[k]): S DFMA and ceil(S / 2)
// three-input logic operations.  Bit 30 of the result is clear iff some sample has |y| < 2.
template <int S>
__device__ __forceinline__ unsigned sweep_yand(double vs, double ti, const double (&ck)[S > 0 ? S : 1]) {
    unsigned h[S > 0 ? S : 1];
#pragma unroll
    for (int k = 0; k < S; ++k) h[k] = (unsigned)__double2hiint(fma(vs, ti, ck[k]));
    unsigned a = (S & 1) ? h[S - 1] : (h[S - 2] & h[S - 1]);
#pragma unroll
    for (int k = 0; k + 1 < S - ((S & 1) ? 0 : 2); k += 2) a &= h[k] & h[k + 1];
    return a;
}
// One thread per walker: proposal (sampler), model constants and the sweep's classification constants, once, for
// ensembles whose walker groups are swept by many time blocks (launched in front of sweep_walkers_kernel).
static __global__ void __launch_bounds__(128)
sweep_consts_kernel(LcView lc, const double* __restrict__ params, int W, EvalWs ws, SampleCtx sc) {
    const long long cw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Sx = lc.S;
    const bool active = cw < (long long)W * lc.ncand;
    const int cand = active ? (int)(cw / W) : 0, w = active ? (int)(cw % W) : 0;
    const double span = lc.off[(long long)cand * Sx + Sx - 1];
    pdl_wait();     // walker positions come from the previous kernel
    pdl_trigger();  // the sweep may be scheduled; it waits for this grid to finish before reading
    if (!active) return;
    double par[6];
    if (sc.enabled) {
        const Proposal pr = propose_compute(sc, cand, w);
#pragma unroll
        for (int d = 0; d < 6; ++d) par[d] = pr.q[d];
        propose_store(sc, pr, cand, w, W);
    } else {
#pragma unroll
        for (int d = 0; d < 6; ++d) par[d] = params[cw * 6 + d];
    }
    const WalkerConsts wc = make_walker_consts(par, Sx);
    ws.wcs[cw] = wc;
    const SweepConsts c = make_sweep_consts(wc, span);
    ws.swc[cw] = c;
}

template <int S, int THREADS, int CH, int MINB, bool SUB>
__global__ void __launch_bounds__(THREADS, MINB)
sweep_walkers_kernel(LcView lc, const double* __restrict__ params, int W, int gpb,
                     const double* __restrict__ priors, EvalWs ws, double* __restrict__ lnprob,
                     double* __restrict__ loglike, SampleCtx sc) {
    using Smem = WalkSmem<S, CH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int tb = blockIdx.x, wb = blockIdx.y, cand = blockIdx.z;
    const int ntb = gridDim.x, nwb = gridDim.y;
    const int Sx = s_of<S>(lc.S);
    const int SP = S > 0 ? Smem::SP : ((Sx + 1) & ~1);
    const int ngran = lc.ngran;
    const int cpg = lc.gran / CH;  // chunks per granule
    const int n = (int)lc.n[cand];
    const int g0 = tb * gpb, g1 = min(g0 + gpb, ngran);
    const int c_begin = g0 * cpg, c_end = min(g1 * cpg, (n + CH - 1) / CH);  // live chunks of this block
    // Walker of this lane.  A warp whose share of the ensemble is 16 walkers or fewer (the last warp of a 100-walker
    // ensemble holds 4) gives each walker F = 2, 4 or 8 lanes, which split the trips of a chunk between them: the
    // classification -- all of the sweep's arithmetic -- then costs that warp 1/F of a full pass instead of a full
    // pass for a handful of live lanes.  The walker's first lane (`leader`) still adds the chunk's terms in
    // timestamp order, so the partial sums have the bits of the one-lane-per-walker layout.
    // (Rotating the light warp's place in the block from block to block, so that it does not always land on the same
    // sub-partition, was measured and loses: 0.166 -> 0.185 ms on the candidate batch.)  SUB = false compiles the
    // one-lane-per-walker layout only: ensembles without such a tail warp run the kernel they always ran.
    const int lane = tid & 31, wbase = wb * THREADS + (tid & ~31);
    int F = 1;
    if (S > 0 && SUB) {
        const int rem = W - wbase;
        F = (rem >= 17 || rem <= 0) ? 1 : rem >= 9 ? 2 : rem >= 5 ? 4 : 8;
    }
    const int sub = lane & (F - 1);
    const int w = wbase + lane / F;
    const bool active = w < W;
    const bool leader = active && sub == 0;
    const long long cw = (long long)cand * W + w;
    const double* __restrict__ gt = lc.t + (long long)cand * lc.npad;
    const double* __restrict__ gterm = lc.term + (long long)cand * lc.npad;
    const double* __restrict__ gtsum = lc.tsum + (long long)cand * (lc.npad / kSweepTrip);

    if (tid == 0) {
        trace_min(ws, 0);
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        if (c_begin < c_end) {
            mbar_expect_tx(&sm.mbar[0], (2u * CH + CH / kSweepTrip) * sizeof(double));
            tma_bulk_g2s(sm.traw[0], gt + (long long)c_begin * CH, CH * sizeof(double), &sm.mbar[0]);
            tma_bulk_g2s(sm.term[0], gterm + (long long)c_begin * CH, CH * sizeof(double), &sm.mbar[0]);
            tma_bulk_g2s(sm.tsum[0], gtsum + (long long)c_begin * (CH / kSweepTrip), (CH / kSweepTrip) * sizeof(double), &sm.mbar[0]);
        }
    }
    if (tid < Sx) sm.off[tid] = lc.off[(long long)cand * Sx + tid];
    pdl_wait();  // everything above reads only the staged light curve; walkers come from the previous kernel

    // per-walker constants (every block of the walker group derives them; the first one publishes
    // what stage 2 and the finalisation need: proposal and model constants)
    SweepConsts swc = {0.0, 4.0, 1.0, 0.0};  // padding lanes never hit
    const double span = lc.off[(long long)cand * Sx + Sx - 1];  // staged light curve: read before the grid dependency
    if (active && ws.pre_consts) {
        // many time blocks per walker group: sweep_consts_kernel derived the constants once (32 B per walker to load
        // here instead of a chain of ~150 FP64 instructions that queues behind the other warps' DFMA streams)
        const double2 v0 = __ldcg(reinterpret_cast<const double2*>(ws.swc + cw));
        const double2 v1 = __ldcg(reinterpret_cast<const double2*>(ws.swc + cw) + 1);
        swc.vel = v0.x; swc.nvtc = v0.y; swc.s = v1.x; swc.vs = v1.y;
    } else if (active) {
        double par[6];
        if (sc.enabled) {
            const Proposal pr = propose_compute(sc, cand, w);
#pragma unroll
            for (int d = 0; d < 6; ++d) par[d] = pr.q[d];
            if (tb == 0 && sub == 0) propose_store(sc, pr, cand, w, W);
        } else {
#pragma unroll
            for (int d = 0; d < 6; ++d) par[d] = params[cw * 6 + d];
        }
        const WalkerConsts wc = make_walker_consts(par, Sx);
        swc = make_sweep_consts(wc, span);
        if (tb == 0 && sub == 0) ws.wcs[cw] = wc;
    }
    __syncthreads();  // mbarrier init and offsets visible
    if (tid == 0) trace_max(ws, 1);  // constants ready
    // compiled-in S: the walker's S sample offsets live in registers, ck = (vel * off_k - vel * tcen) * s, so a
    // fine sample is y = fma(V, t_i, ck) with t_i one broadcast load per timestamp and no table
    double ck[S > 0 ? S : 1];
    if (S > 0) {
#pragma unroll
        for (int k = 0; k < S; ++k) ck[k] = fma(swc.vel, sm.off[k], swc.nvtc) * swc.s;
    }

    const double vs = swc.vs, c0s = swc.nvtc * swc.s;
    double acc = 0.0;
    int lo = INT_MAX, hi = 0;
    double* __restrict__ prow = ws.partials + (long long)cand * ngran * W + w;
    for (int c = c_begin; c < c_end; ++c) {
        const int it = c - c_begin, buf = it & 1;
        mbar_wait(&sm.mbar[buf], (it >> 1) & 1);
        if (S == 0) {  // run-time S: fine-time table of this chunk, ft[i][k] = t_i + off_k
            const double* tr = sm.traw[buf];
            double* ft = sm.ftab[buf];
            for (int e = tid; e < CH * Sx; e += THREADS) {
                const int i = e / Sx, k = e - i * Sx;
                ft[i * SP + k] = tr[i] + sm.off[k];
            }
        }
        // every thread is past the previous chunk (and the table is complete).  (A ring of four buffers with
        // per-buffer "empty" barriers instead of this block-wide one -- warps free to drift two chunks apart -- was
        // measured and is slower: 0.474 against 0.448 ms on the 65 k x 4096 sweep, profiles/r02a2_ab_sweep_ring.txt.)
        __syncthreads();
        if (tid == 0 && c + 1 < c_end) {
            mbar_expect_tx(&sm.mbar[buf ^ 1], (2u * CH + CH / kSweepTrip) * sizeof(double));
            tma_bulk_g2s(sm.traw[buf ^ 1], gt + (long long)(c + 1) * CH, CH * sizeof(double), &sm.mbar[buf ^ 1]);
            tma_bulk_g2s(sm.term[buf ^ 1], gterm + (long long)(c + 1) * CH, CH * sizeof(double), &sm.mbar[buf ^ 1]);
            tma_bulk_g2s(sm.tsum[buf ^ 1], gtsum + (long long)(c + 1) * (CH / kSweepTrip), (CH / kSweepTrip) * sizeof(double),
                         &sm.mbar[buf ^ 1]);
        }
        const double* tr = sm.traw[buf];
        const double* tm = sm.term[buf];
        const int base = c * CH;
        // stage 2 reads flux and weight of the in-transit timestamps, which this kernel never touches: the
        // first walker group asks L2 for the chunk's lines now (16 B per timestamp, once per light curve)
        if (wb == 0 && tid < 2 * (CH * 8 / 128)) {
            const double* src = (tid & 1 ? lc.w : lc.f) + (long long)cand * lc.npad + base + (tid >> 1) * 16;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
        }
        if (S > 0 && SUB && F > 1) {
            // several lanes per walker: lane `sub` classifies the trips sub, sub + F, ... of the chunk and marks
            // the timestamps that may be in transit; the marks of the walker's lanes are combined and its leader
            // adds the other timestamps' terms in timestamp order
            static_assert(CH <= 64, "one mark per timestamp of a chunk in a 64-bit word");
            constexpr int NT = kSweepTrip;
            unsigned long long mask = 0ull;
#pragma unroll 1
            for (int i0 = sub * NT; i0 < CH; i0 += F * NT) {
                unsigned q[NT];
                unsigned qall = 0xffffffffu;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    q[j] = sweep_yand<S>(vs, tr[i0 + j], ck);
                    qall &= q[j];
                }
                if (!(qall & kSweepOutBit)) {
#pragma unroll
                    for (int j = 0; j < NT; ++j)
                        if (!(q[j] & kSweepOutBit)) mask |= 1ull << (i0 + j);
                }
            }
            for (int o = 1; o < F; o <<= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
            if (sub == 0) {   // the same association as the one-lane layout: trip sums, single terms inside marked trips
                const double* ts = sm.tsum[buf];
                for (int tr_ = 0; tr_ < CH / NT; ++tr_) {
                    const unsigned long long mt = (mask >> (tr_ * NT)) & ((1ull << NT) - 1ull);
                    if (mt == 0ull) {
                        acc += ts[tr_];
                    } else {
                        for (int j = 0; j < NT; ++j) {
                            const int i = tr_ * NT + j;
                            if ((mt >> j) & 1ull) {
                                lo = min(lo, base + i);
                                hi = base + i + 1;
                            } else {
                                acc += tm[i];
                            }
                        }
                    }
                }
            }
        } else if (S > 0) {
            // NT timestamps per trip.  The common case -- none of them can be in transit -- costs, beside the
            // S DFMA and S / 2 logic operations per timestamp, one more AND, a bit test and a branch per
            // trip and ONE DADD: the trip's terms were summed (in timestamp order) when the light curve was staged.
            // The per-timestamp book-keeping, and the terms themselves, are touched only on the rare trips that
            // reach the transit; a walker's partial sum is therefore "trip sums, and single terms inside the trips
            // around its transit", the same association on every launch shape and lane layout.
            constexpr int NT = kSweepTrip;
            const double* ts = sm.tsum[buf];
#pragma unroll 1
            for (int i0 = 0; i0 < CH; i0 += NT) {
                double tt[NT];
                unsigned q[NT];
#pragma unroll
                for (int j = 0; j < NT; j += 2) {
                    const double2 tv = *reinterpret_cast<const double2*>(tr + i0 + j);
                    tt[j] = tv.x; tt[j + 1] = tv.y;
                }
                const double tsum = ts[i0 / NT];
                unsigned qall = 0xffffffffu;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    q[j] = sweep_yand<S>(vs, tt[j], ck);
                    qall &= q[j];
                }
                if (qall & kSweepOutBit) {
                    acc += tsum;
                } else {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        if (q[j] & kSweepOutBit) {
                            acc += tm[i0 + j];
                        } else {
                            lo = min(lo, base + i0 + j);
                            hi = base + i0 + j + 1;
                        }
                    }
                }
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < CH; ++i) {
                unsigned yand = 0xffffffffu;
                const double* row = sm.ftab[buf] + i * SP;
                for (int k = 0; k < Sx; ++k) yand &= (unsigned)__double2hiint(fma(vs, row[k], c0s));
                const bool hit = !(yand & kSweepOutBit);
                const double term = tm[i];
                if (!hit) acc += term;
                if (hit) {
                    lo = min(lo, base + i);
                    hi = base + i + 1;
                }
            }
        }
        if ((c + 1) % cpg == 0 || c + 1 == c_end) {  // granule complete: publish its partial sum
            if (leader) prow[(long long)(c / cpg) * W] = acc;
            acc = 0.0;
        }
    }
    pdl_trigger();  // stage 2 may be scheduled; it waits for this grid to finish before reading
    {   // granules of this span that lie entirely beyond the light curve
        const int gdone = (c_end > c_begin) ? (c_end + cpg - 1) / cpg : g0;
        if (leader)
            for (int g = gdone; g < g1; ++g) prow[(long long)g * W] = 0.0;
    }
    if (leader && hi > lo) {
        atomicMin(ws.rlo + cw, lo);
        atomicMax(ws.rhi + cw, min(hi, n));
    }

    // ---- last block of this walker group: fold granules in order, settle walkers that never transit
    if (tid == 0) trace_max(ws, 2);  // sweep done
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int* tk = ws.grouptick + (long long)cand * nwb + wb;
        const int t = atomicAdd(tk, 1);
        sm.flag = (t == ntb - 1);
        if (sm.flag) *tk = 0;
    }
    __syncthreads();
    if (!sm.flag) return;
    __threadfence();
    // (one walker per thread again, whatever the lane layout of the sweep was)
    int lo_w = 0, hi_w = 0;
    const int wf = wb * THREADS + tid;
    const long long cwf = (long long)cand * W + wf;
    if (wf < W) {
        const double* col = ws.partials + (long long)cand * ngran * W + wf;
        double sacc = 0.0;
        int g = 0;
        for (; g + 16 <= ngran; g += 16) {  // sixteen loads in flight, summed in granule order
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __ldcg(col + (long long)(g + j) * W);
#pragma unroll
            for (int j = 0; j < 16; ++j) sacc += v[j];
        }
        for (; g < ngran; ++g) sacc += __ldcg(col + (long long)g * W);
        lo_w = __ldcg(ws.rlo + cwf);
        hi_w = __ldcg(ws.rhi + cwf);
        if (hi_w - lo_w <= 0) {  // never in transit: done here
            const double* par = (sc.enabled ? sc.q : params) + cwf * 6;
            const double lp = priors ? ln_prior_dev(par, priors + (long long)cand * 36, 6) : 0.0;
            finish_walker(sc, cwf, W, sacc, lp, lnprob, loglike);
            ws.rlo[cwf] = INT_MAX;
            ws.rhi[cwf] = 0;
            lo_w = hi_w = 0;
        } else {
            ws.oot[cwf] = sacc;
        }
    }
    push_slots<THREADS>(ws, cwf, lo_w, hi_w, sm.s_scan);
    if (tid == 0) trace_max(ws, 3);  // fold + queue push done
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

// chi^2 of the timestamps outside [ilo, ihi): pre[ilo] + (pre[n] - pre[ihi]) in double-double.  The six
// prefix-sum words are loaded by the constructor and combined by value(), so a caller can issue the loads
// early and pay their latency under other work.
struct window_oot {
    double hn, hh, hl, ln, lh, ll;
    __device__ __forceinline__ window_oot(const LcView& lc, int cand, int n, int ilo, int ihi) {
        const double* ph = lc.pre_hi + (long long)cand * (lc.npad + 1);
        const double* pl = lc.pre_lo + (long long)cand * (lc.npad + 1);
        hn = ph[n]; hh = ph[ihi]; hl = ph[ilo];
        ln = pl[n]; lh = pl[ihi]; ll = pl[ilo];
    }
    __device__ __forceinline__ double value() const {
        double sa, ea, sb, eb;
        two_sum(hn, -hh, sa, ea);
        ea += ln - lh;
        two_sum(sa, hl, sb, eb);
        eb += ea + ll;
        return sb + eb;
    }
};

// Conservative in-transit index range [ilo, ihi) of one walker from the geometry (O(1): two lookups in
// the uniform time grid) and the chi^2 of everything outside it from the double-double prefix sums.
// Exact: the range is a superset of the in-transit timestamps and stage 2 classifies every sample.
// Per-candidate scalars of the window lookup.  They belong to the staged light curve, so a kernel loads them
// before its grid dependency resolves and before the walker's own chain of dependent loads starts.
struct CandGeom {
    int n, gc;
    double span, t0, ginv;
};
__device__ __forceinline__ CandGeom load_cand_geom(const LcView& lc, int cand, int Sx) {
    CandGeom g;
    g.n = (int)lc.n[cand];
    g.gc = lc.gcount[cand];
    g.span = lc.off[(long long)cand * Sx + Sx - 1];
    g.t0 = lc.t0[cand];
    g.ginv = lc.ginv[cand];
    return g;
}
__device__ __forceinline__ void walker_range(const LcView& lc, int cand, const WalkerConsts& wc, const CandGeom& geo,
                                             int& ilo, int& ihi) {
    const int n = geo.n;
    const double span = geo.span;
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
    if (full) {
        ilo = 0; ihi = n;
    } else if (empty) {
        ilo = 0; ihi = 0;
    } else {
        // O(1) lookup in the uniform time grid, widened by one cell on each side
        const double t0 = geo.t0, ginv = geo.ginv;
        const int gc = geo.gc;
        const int* __restrict__ grid = lc.tgrid + (long long)cand * lc.gstride;
        const double jl = floor((tlo - t0) * ginv) - 1.0, jh = floor((thi - t0) * ginv) + 2.0;
        const int j0 = jl < 0.0 ? 0 : (jl > (double)gc ? gc : (int)jl);
        const int j1 = jh < 0.0 ? 0 : (jh > (double)gc ? gc : (int)jh);
        ilo = __ldg(grid + j0);
        ihi = (jh > (double)gc) ? n : __ldg(grid + j1);
        if (ihi < ilo) ihi = ilo;
#ifndef NMB_NO_WINDOW_TRIM
        // The cell lookup lands one to two cadences outside [tlo, thi] on either side: three timestamps of ~30 whose
        // samples all leave the flux function at once but still take lanes in the first and last rounds of the range.
        // A timestamp has a sample inside the window only if tlo <= t_i <= thi (tlo already carries the span of the
        // supersampling offsets), so up to two timestamps per side are dropped by looking at their times.
        if (ihi - ilo >= 4) {
            const double* __restrict__ tt = lc.t + (long long)cand * lc.npad;
            const double a0 = __ldg(tt + ilo), a1 = __ldg(tt + ilo + 1);
            const double b0 = __ldg(tt + ihi - 1), b1 = __ldg(tt + ihi - 2);
            if (a0 < tlo) ilo += (a1 < tlo) ? 2 : 1;
            if (b0 > thi) ihi -= (b1 > thi) ? 2 : 1;
        }
#endif
    }
}
__device__ __forceinline__ void walker_window(const LcView& lc, int cand, const WalkerConsts& wc, const CandGeom& geo,
                                              int& ilo, int& ihi, double& oot) {
    walker_range(lc, cand, wc, geo, ilo, ihi);
    oot = window_oot(lc, cand, geo.n, ilo, ihi).value();
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
window_ranges_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors,
                     EvalWs ws, double* __restrict__ lnprob, double* __restrict__ loglike, SampleCtx sc) {
    __shared__ int s_scan[THREADS / 32 + 1];
    const int tid = threadIdx.x;
    const int CW = W * lc.ncand;
    int lo_w = 0, hi_w = 0;
    const int cw = blockIdx.x * THREADS + tid;
    const int Sx = lc.S;
    CandGeom geo = CandGeom{};
    if (cw < CW) geo = load_cand_geom(lc, cw / W, Sx);
    pdl_wait();     // walker positions come from the previous kernel
    pdl_trigger();  // stage 2 may be scheduled; it waits for this grid to finish before reading
    if (tid == 0) trace_min(ws, 0);
    if (cw < CW) {
        const int cand = cw / W;
        const double* par = params;
        if (sc.enabled) {
            propose_stretch(sc, cand, cw % W, W);
            par = sc.q;  // the proposal just written by this thread (not through the read-only path)
        }
        const WalkerConsts wc = make_walker_consts(par + (long long)cw * 6, Sx);
        int ilo, ihi;
        double oot;
        walker_window(lc, cand, wc, geo, ilo, ihi, oot);
        if (ihi - ilo <= 0) {  // never in transit: done here (the GP solve reads the idle range instead)
            if (!ws.modelbuf) {
                const double lp = priors ? ln_prior_dev(par + (long long)cw * 6, priors + (long long)cand * 36, 6) : 0.0;
                finish_walker(sc, cw, W, oot, lp, lnprob, loglike);
            }
        } else {
            ws.oot[cw] = oot;
            ws.wcs[cw] = wc;
            ws.rlo[cw] = ilo;
            ws.rhi[cw] = ihi;
            lo_w = ilo;
            hi_w = ihi;
        }
    }
    push_slots<THREADS>(ws, cw, lo_w, hi_w, s_scan);
    if (tid == 0) trace_max(ws, 3);
}

// ---- stage 2: evaluate the in-transit ranges -------------------------------------------------
// The queue of slots is split evenly and statically over all resident warps (no work-fetching
// atomics).  A warp walks its slots pass by pass: one supersample per lane and round, fluxes parked
// in its private shared-memory scratch, bins reduced in NumPy's order by one lane per timestamp,
// residuals accumulated; each slot's partial is stored.  When the warp leaves a walker it reports
// how many of the walker's slots it finished; the warp that completes the walker folds the slot
// partials in slot order, evaluates the priors (one parameter per lane) and writes lnprob or runs
// the sampler's accept step.
#ifndef NMB_EVALPASS
#define NMB_EVALPASS 96
#endif
constexpr int kEvalPass = NMB_EVALPASS;  // supersamples of one slot (the unit of the stored partial sums)
#ifndef NMB_EVALMERGE
#define NMB_EVALMERGE 4
#endif
constexpr int kEvalMerge = NMB_EVALMERGE;  // most consecutive slots of a walker a warp evaluates in one pass

// PATH: 0 = general IEEE FP64, 1 = FP64 hot path in the reference's rounding sequence (NMB_F64_STRICT),
//       2 = FP32 geometry + FP64 occultation (NMB_F32), 3 = relaxed FP64 hot path (NMB_F64, the default),
//       4 = FP32 occultation (NMB_F32_PURE; queue path only)
template <int PATH>
__device__ __forceinline__ double sample_flux_eval(double ti, double offk, const WalkerConsts& wc,
                                                   const QuadConstsF& cf, const QuadConsts* gq);

static __device__ __noinline__ double sample_flux_slowpath(double z, const QuadConsts* q) {
    return occultquad_point<false, IeeeOps>(z, *q, nullptr, nullptr, nullptr);
}

template <>
__device__ __forceinline__ double sample_flux_eval<1>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF&, const QuadConsts* gq) {
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
        F = (zz > wc.q.onep) ? wc.q.fout : sample_flux_slowpath(zz, gq);
    }
    return F;
}
template <>
__device__ __forceinline__ double sample_flux_eval<3>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF&, const QuadConsts* gq) {
    // geometry exactly as in the strict path (z is the correctly rounded root, so the case a sample
    // falls in is the reference's); only the occultation arithmetic is relaxed
    const double ft = ti + offk;
    const double x = ft - wc.tcen;
    const double vx = wc.vel * x;
    const double q = wc.b2 + vx * vx;
    const double z = FastOps::sqrt(q);
    if (z > wc.q.onep) return wc.q.fout;
    double F;
    if (!occultquad_relaxed(z, wc.q, F) || !(fabs(F) <= 1e300)) {
        const double zz = ::sqrt(q);
        F = (zz > wc.q.onep) ? wc.q.fout : sample_flux_slowpath(zz, gq);
    }
    return F;
}
template <>
__device__ __forceinline__ double sample_flux_eval<0>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF&, const QuadConsts*) {
    return sample_flux(ti, offk, wc);
}
// NMB_F32: FP32 where the 1e-5 tolerance on lnprob allows it -- the geometry's square root and with it the case a
// sample falls in -- and FP64 (the relaxed arithmetic of the default) where it does not: K, E, Pi and their
// combination lambda_d lose a factor ~60 to cancellation, and the limb's acos / sqrt amplify argument rounding by
// 1 / sqrt(distance to contact).  The occultation is the FP64 function evaluated at the float-rounded z
// (|dz| <= 3e-8 z, |dF| <= 1e-8).  lnprob stays within ~1e-7 of the FP64 result on every fixture.
template <>
__device__ __forceinline__ double sample_flux_eval<2>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF&, const QuadConsts* gq) {
    const double ft = ti + offk;
    const double x = ft - wc.tcen;          // FP64 difference of the absolute times (FP32 here costs 2e-4 on ll)
    const double vx = wc.vel * x;
    const double q = fma(vx, vx, wc.b2);
    const double z = (double)sqrtf((float)q);
    if (z > wc.q.onep) return wc.q.fout;
    double F;
    if (!occultquad_relaxed(z, wc.q, F) || !(fabs(F) <= 1e300)) {
        const double zz = ::sqrt(q);
        F = (zz > wc.q.onep) ? wc.q.fout : sample_flux_slowpath(zz, gq);
    }
    return F;
}
// NMB_F32_PURE: the two hot cases entirely in FP32 (depth space).  Reported with its error: 1e-5 ... 1e-4 on
// lnprob, because FP32 rounding of K, E and Pi (1e-7 each) is amplified ~60x by the cancellation in lambda_d.
template <>
__device__ __forceinline__ double sample_flux_eval<4>(double ti, double offk, const WalkerConsts& wc,
                                                      const QuadConstsF& cf, const QuadConsts* gq) {
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
    return (zz > wc.q.onep) ? wc.q.fout : sample_flux_slowpath(zz, gq);
}

template <int THREADS>
struct EvalSmem {
    // per warp: TSW bins of S fluxes at an odd stride; TSW * S <= kEvalPass
    // ... for up to kEvalMerge slots sharing one pass, followed by the chi^2 terms of that pass
    double scratch[THREADS / 32][kEvalMerge * (2 * kEvalPass + 8) + kEvalMerge * kEvalPass];
    // the current walker's constants, one copy per warp: operands come from shared memory when
    // needed instead of pinning ~40 registers for the whole pass
    WalkerConsts wc[THREADS / 32];
};

template <int S, int THREADS, int PATH, bool PACKED>
__device__ __forceinline__ void
eval_slots_body(EvalSmem<THREADS>& sm, const int nblocks, const int bid, const LcView& lc, const double* __restrict__ params, int W,
                const double* __restrict__ priors, const EvalWs& ws, int nsm, double* __restrict__ lnprob,
                double* __restrict__ loglike, const SampleCtx& sc) {
    constexpr int NWARP = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Sx = s_of<S>(lc.S);
    const int SP = Sx | 1;
    const int TSW = ws.tsw;
    pdl_wait();  // stage 1 complete, its queue visible
    const int total = __ldcg(ws.counters + 1);
    double* scratch = sm.scratch[warp];
    if (threadIdx.x == 0) trace_min(ws, 8);

    // warp index that runs over SMs first, then sub-partitions, then resident blocks of an SM
    // (blocks b, b + nsm, ... share an SM in a one-wave launch), so a short queue still touches
    // every sub-partition of the GPU
    const int nwarps = nblocks * NWARP;
    // (PACKED, eval_mixed_kernel: block after block instead, so that a short queue keeps few blocks busy and
    // the others leave their SM slots to the direct blocks at once)
    const int gw = PACKED ? bid * NWARP + warp : ((bid / nsm) * NWARP + warp) * nsm + bid % nsm;
    // chunks of slots: the first chunk of a warp is its own index (no atomic on the way to the first
    // sample), further chunks come from a shared counter, fetched one chunk ahead.  Small queues are
    // dealt out in one chunk per warp; large ones in chunks of at least one merged pass.
    const int per = (total + nwarps - 1) / nwarps;
    // slots a warp evaluates in one pass: the staged default, one more when the queue is short (a few slots per
    // warp: the warps' latency chains -- slot record, walker constants, completion count -- weigh more than the
    // balance of the last chunks).  A queue of up to two passes per warp is dealt out statically, one chunk per
    // warp and no work-fetching atomic; measured on the 2048-walker half-steps of the 65 k-point light curve,
    // S = 15: 67 -> 61 us with three slots per pass, -> 52 us with the static deal up to two passes
    // (profiles/r02v_ab_stage2_short_queue.txt); its 4096-walker evaluation with prior-box walkers, 40 slots per
    // warp, prefers two slots per pass and fetched chunks.  Results do not depend on either (see SlotRec).
    const int merge = per <= 8 ? min(kEvalMerge, ws.merge + 1) : ws.merge;
    const int chunk = per <= 2 * merge ? max(1, per) : max(merge, total / (nwarps * 8));
    // The static deal is balanced: every warp takes floor(total / nwarps) slots and the first total mod nwarps warps
    // one more (the warp index runs over the SMs first, so the extra slots spread over the GPU), instead of
    // ceil(total / nwarps) slots for as many warps as it takes: with 4.15 slots per warp -- the 2048-walker half-steps
    // of the 65 k-point light curve -- no sub-partition carries five warps of five slots while others idle.
    // (eval_mixed_kernel's workers, PACKED, keep the block-after-block deal: their aim is to leave blocks free.)
#ifndef NMB_STATIC_MAX_PER
#define NMB_STATIC_MAX_PER 24
#endif
    // A longer queue -- up to 24 slots per warp -- is dealt statically too while every slot is a single pass (stage 1
    // raises counters[5] when it pushes a walker whose range exceeds the per-walker slot cap: slots of several
    // passes, very unequal work).  A warp then stays with one or two walkers (constants, completion count) and
    // fetches nothing.  Measured (profiles/r02i2_ab_stage2_static_uniform.txt): half-step of a 1024-walker share of
    // the 500 k-point light curve 137 -> 130 us, of the 256-candidate batch 179 -> 169 us; at 50 slots per warp (the
    // batch's full evaluation, candidates of unequal cost) the static deal loses 8 %, and with the prior-box walkers'
    // multi-pass slots it loses 20-60 %, hence the two conditions.
    const bool uniform = __ldcg(ws.counters + 5) == 0;
    const bool balanced = !PACKED && (per <= 2 * merge || (uniform && per <= NMB_STATIC_MAX_PER));
    const int base = total / nwarps, extra = total - base * nwarps;
    const int nchunks = balanced ? min(total, nwarps) : (total + chunk - 1) / chunk;
    int* queue = ws.counters + 2;
    int ch_cur = gw, ch_nxt = INT_MAX;
    if (nchunks > nwarps) {
        if (lane == 0) ch_nxt = nwarps + atomicAdd(queue, 1);
        ch_nxt = __shfl_sync(0xffffffffu, ch_nxt, 0);
    }

    int cur_cw = -1, cur_first = 0, seg = 0, cand = 0, cur_nsl = 0;
    WalkerConsts& wc = sm.wc[warp];
    QuadConstsF cf = QuadConstsF{};
    const double *gt = nullptr, *gf = nullptr, *gw_ = nullptr, *goff = nullptr;
    double* terms = scratch + kEvalMerge * (2 * kEvalPass + 8);  // chi^2 terms of a merged pass

    // leaving a walker: report `seg` finished slots; the warp that completes it folds the slot partials
    // in slot order, evaluates the priors and finishes the walker
    auto leave_walker = [&]() {
        if (cur_cw < 0 || ws.modelbuf) return;
        int done = 0;
        const int nsl = cur_nsl;
        if (lane == 0 && seg < nsl) {
            __threadfence();
            done = atomicAdd(ws.slotsdone + cur_cw, seg);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done + seg == nsl) {
            if (seg < nsl) __threadfence();
            __syncwarp();
            const double* part = ws.slotpart + cur_first;
            double ssum = 0.0;
            for (int c = lane; c < nsl; c += 32) ssum += __ldcg(part + c);
            ssum = warp_sum(ssum);
            const double* par = (sc.enabled ? sc.q : params) + (long long)cur_cw * 6;
            const double lp = priors ? ln_prior_warp(par, priors + (long long)(cur_cw / W) * 36, 6, lane) : 0.0;
            if (lane == 0) {
                finish_walker(sc, cur_cw, W, ws.oot[cur_cw] + ssum, lp, lnprob, loglike);
                ws.slotsdone[cur_cw] = 0;
                ws.rlo[cur_cw] = INT_MAX;
                ws.rhi[cur_cw] = 0;
            }
        }
    };

    if (lane == 0) trace_max(ws, 9);
    while (ch_cur < nchunks) {
      const int s0 = balanced ? ch_cur * base + min(ch_cur, extra) : ch_cur * chunk;
      const int s1 = balanced ? s0 + base + (ch_cur < extra ? 1 : 0) : min(total, s0 + chunk);
      int ch_nn = INT_MAX;
      if (lane == 0 && ch_nxt < nchunks) ch_nn = nwarps + atomicAdd(queue, 1);
      int s = s0;
      while (s < s1) {
        // the next kEvalMerge slot records, one per lane
        SlotRec mine;
        mine.cw = -2; mine.t0 = 0; mine.t1 = 0; mine.first = 0;
        if (lane < kEvalMerge && s + lane < s1)
            *reinterpret_cast<int4*>(&mine) = __ldcg(reinterpret_cast<const int4*>(ws.slots + s + lane));
        const int cw0 = __shfl_sync(0xffffffffu, mine.cw, 0);
        const int t0 = __shfl_sync(0xffffffffu, mine.t0, 0);
        int tend = __shfl_sync(0xffffffffu, mine.t1, 0);
        if (!ws.modelbuf) {
            // the pass's time / flux / weight lines (<= 32 timestamps of them) start travelling now, under
            // the fetch of the walker's constants and the flux rounds
            const long long at = (long long)(cw0 / W) * lc.npad + min(t0 + lane, (int)lc.npad - 1);
            prefetch_l1(lc.t + at);
            prefetch_l1(lc.f + at);
            prefetch_l1(lc.w + at);
        }
        if (cw0 != cur_cw) {
            leave_walker();
            cur_cw = cw0;
            cur_first = __shfl_sync(0xffffffffu, mine.first, 0);
            seg = 0;
            cand = cur_cw / W;
            cur_nsl = __ldcg(ws.nslots + cur_cw);  // needed when the walker is left: requested here
            {
                static_assert(sizeof(WalkerConsts) % 8 == 0, "WalkerConsts is copied as doubles");
                const double* src = reinterpret_cast<const double*>(ws.wcs + cur_cw);
                double* dst = reinterpret_cast<double*>(&wc);
                __syncwarp();
                for (int i = lane; i < (int)(sizeof(WalkerConsts) / 8); i += 32) dst[i] = __ldcg(src + i);
                __syncwarp();
            }
            if (PATH == 4) cf = to_f32(wc.q);
            gt = lc.t + (long long)cand * lc.npad;
            gf = lc.f + (long long)cand * lc.npad;
            gw_ = lc.w + (long long)cand * lc.npad;
            goff = lc.off + (long long)cand * Sx;
        }
        // how many consecutive single-pass slots of this walker can share one pass
        int run = 1;
        const bool single = (tend - t0 <= TSW);
        if (single) {
#pragma unroll
            for (int l = 1; l < kEvalMerge; ++l) {
                const int cwl = __shfl_sync(0xffffffffu, mine.cw, l);
                const int t0l = __shfl_sync(0xffffffffu, mine.t0, l), t1l = __shfl_sync(0xffffffffu, mine.t1, l);
                if (run == l && l < merge && cwl == cur_cw && t0l == tend && t1l - t0l <= TSW) { run = l + 1; tend = t1l; }
            }
        }
        if (single) {
            // ---- one pass over up to kEvalMerge slots: every lane stays busy through the flux rounds,
            // bins and residuals are formed once, then each slot's terms are summed in its own canonical
            // lane order so the partials do not depend on how slots were merged
            const int nts = tend - t0;
            const int nsamp = nts * Sx;
            {   // the time and offset of the NEXT round are requested before this round's flux is evaluated
                int h = lane / Sx, k = lane - h * Sx;
                double ti = lane < nsamp ? __ldg(gt + t0 + h) : 0.0, ok_ = lane < nsamp ? __ldg(goff + k) : 0.0;
                for (int j = lane; j < nsamp; j += 32) {
                    const int slot = h * SP + k;
                    const double ti_c = ti, ok_c = ok_;
                    const int jn = j + 32;
                    if (jn < nsamp) {
                        h = jn / Sx; k = jn - h * Sx;
                        ti = __ldg(gt + t0 + h); ok_ = __ldg(goff + k);
                    }
                    scratch[slot] = sample_flux_eval<PATH>(ti_c, ok_c, wc, cf, &ws.wcs[cur_cw].q);
                }
            }
            __syncwarp();
            for (int h = lane; h < nts; h += 32) {
                const double m = (S > 0) ? bin_mean<(S > 0 ? S : 1)>(scratch + h * SP)
                                         : bin_mean_rt(scratch + h * SP, Sx);
                if (ws.modelbuf) {
                    ws.modelbuf[(long long)cur_cw * lc.npad + t0 + h] = m;
                } else {
                    const double r = __ldg(gf + t0 + h) - m;
                    terms[h] = __ldg(gw_ + t0 + h) * (r * r);
                }
            }
            __syncwarp();
            if (!ws.modelbuf) {
                for (int q = 0; q < run; ++q) {
                    const int b0 = q * TSW, nq = min(TSW, nts - b0);
                    double acc = 0.0;
                    for (int h = lane; h < nq; h += 32) acc += terms[b0 + h];
                    acc = warp_sum(acc);
                    if (lane == 0) ws.slotpart[s + q] = acc;
                }
            }
            __syncwarp();
        } else {
            // ---- a slot of several passes (walkers whose range exceeds the per-walker slot cap)
            double acc = 0.0;
            for (int sub = t0; sub < tend; sub += TSW) {
                const int nts = min(TSW, tend - sub);
                const int nsamp = nts * Sx;
                for (int j = lane; j < nsamp; j += 32) {
                    const int h = j / Sx, k = j - h * Sx;
                    scratch[h * SP + k] = sample_flux_eval<PATH>(__ldg(gt + sub + h), __ldg(goff + k), wc, cf, &ws.wcs[cur_cw].q);
                }
                __syncwarp();
                for (int h = lane; h < nts; h += 32) {
                    const double m = (S > 0) ? bin_mean<(S > 0 ? S : 1)>(scratch + h * SP)
                                             : bin_mean_rt(scratch + h * SP, Sx);
                    if (ws.modelbuf) {
                        ws.modelbuf[(long long)cur_cw * lc.npad + sub + h] = m;
                    } else {
                        const double r = __ldg(gf + sub + h) - m;
                        acc += __ldg(gw_ + sub + h) * (r * r);
                    }
                }
                __syncwarp();
            }
            acc = warp_sum(acc);
            if (lane == 0) ws.slotpart[s] = acc;
        }
        seg += run;
        s += run;
        if (lane == 0) trace_max(ws, 10);  // slots evaluated
      }
      leave_walker();
      cur_cw = -1;
      ch_cur = ch_nxt;
      ch_nxt = __shfl_sync(0xffffffffu, ch_nn, 0);
    }
    if (lane == 0) trace_max(ws, 11);  // warp done
    pdl_trigger();
    // the last block to finish re-arms the queue counters for the next evaluation
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(ws.counters + 3, 1);
        if (done == nblocks - 1) {
            ws.counters[1] = 0;
            ws.counters[2] = 0;
            ws.counters[3] = 0;
            ws.counters[4] = 0;
            ws.counters[5] = 0;
        }
    }
}

template <int S, int THREADS, int MINB, int PATH>
__global__ void __launch_bounds__(THREADS, MINB)
eval_slots_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors, EvalWs ws,
                  int nsm, double* __restrict__ lnprob, double* __restrict__ loglike, SampleCtx sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    eval_slots_body<S, THREADS, PATH, false>(*reinterpret_cast<EvalSmem<THREADS>*>(smem_raw), (int)gridDim.x, (int)blockIdx.x, lc, params, W,
                                             priors, ws, nsm, lnprob, loglike, sc);
}

// ---- windowed mode, direct evaluation: one thread block per walker ------------------------------
// For small and medium ensembles the two-kernel pipeline is a chain of dependent global-memory round
// trips (ranges -> queue tail -> slot records -> walker constants -> partials -> completion count),
// each ~1 us, around a few microseconds of arithmetic.  Here a block does everything for its walker
// without leaving the SM: every thread derives the walker's constants and window (uniform work, no
// exchange needed), each of the block's warps evaluates one pass of up to `merge` slots, the slot
// partials meet in shared memory and warp 0 folds them, adds the priors and finishes the walker
// (lnprob store or the sampler's accept step).  Slot boundaries, the summation order inside a slot
// and the fold over slots are those of eval_slots_kernel, so the result has the same bits.  Walkers
// whose range needs more than one pass per warp (slow or grazing walkers; every walker of the long
// light curves) are pushed to the slot queue and evaluated by eval_slots_kernel, launched behind
// this kernel as before.
template <int THREADS>
struct DirectSmem {
    double scratch[THREADS / 32][kEvalMerge * (2 * kEvalPass + 8) + kEvalMerge * kEvalPass];
    WalkerConsts wc[THREADS / 32];
    double part[(THREADS / 32) * kEvalMerge];
    int first;
};

// The direct evaluation proper, shared by window_direct_kernel (range and constants derived in the block)
// and eval_mixed_kernel (range and constants left in the workspace by stage 1).  On entry sm.wc[warp]
// holds the walker's constants; warp w takes the consecutive slots [w * per, (w + 1) * per) as one pass.
// `ln_prior_w0()` / `oot_w0()` are evaluated by warp 0 only.
template <int S, int THREADS, int PATH, class LnPrior, class Oot>
__device__ __forceinline__ void
direct_eval_tail(DirectSmem<THREADS>& sm, const LcView& lc, const EvalWs& ws, const SampleCtx& sc, int cw, int cand, int W,
                 int ilo, int ihi, int nsl, double* __restrict__ lnprob, double* __restrict__ loglike, LnPrior ln_prior_w0,
                 Oot oot_w0) {
    constexpr int NWARP = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Sx = s_of<S>(lc.S);
    const int SP = Sx | 1;
    const int TSW = ws.tsw;
    WalkerConsts& wc = sm.wc[warp];
    QuadConstsF cf = QuadConstsF{};
    if (PATH == 4) cf = to_f32(wc.q);
    const int per = (nsl + NWARP - 1) / NWARP;
    const int s0 = warp * per, s1 = min(nsl, s0 + per);
    if (s0 < s1) {
        double* scratch = sm.scratch[warp];
        double* terms = scratch + kEvalMerge * (2 * kEvalPass + 8);
        const double* gt = lc.t + (long long)cand * lc.npad;
        const double* gf = lc.f + (long long)cand * lc.npad;
        const double* gw_ = lc.w + (long long)cand * lc.npad;
        const double* goff = lc.off + (long long)cand * Sx;
        const int t0 = ilo + s0 * TSW, tend = min(ihi, ilo + s1 * TSW);
        const int nts = tend - t0;
        const int nsamp = nts * Sx;
        // flux and weight of the first 32 timestamps of the pass: requested before the flux rounds
        const double f_pre = lane < nts ? __ldg(gf + t0 + lane) : 0.0;
        const double w_pre = lane < nts ? __ldg(gw_ + t0 + lane) : 0.0;
        {   // the time and offset of the NEXT round are requested before this round's flux is evaluated
            int h = lane / Sx, k = lane - h * Sx;
            double ti = lane < nsamp ? __ldg(gt + t0 + h) : 0.0, ok_ = lane < nsamp ? __ldg(goff + k) : 0.0;
            for (int j = lane; j < nsamp; j += 32) {
                const int slot = h * SP + k;
                const double ti_c = ti, ok_c = ok_;
                const int jn = j + 32;
                if (jn < nsamp) {
                    h = jn / Sx; k = jn - h * Sx;
                    ti = __ldg(gt + t0 + h); ok_ = __ldg(goff + k);
                }
                scratch[slot] = sample_flux_eval<PATH>(ti_c, ok_c, wc, cf, &wc.q);
            }
        }
        __syncwarp();
        if (lane == 0) trace_max(ws, 14);  // flux rounds of a direct pass done
        for (int h = lane; h < nts; h += 32) {
            const double m = (S > 0) ? bin_mean<(S > 0 ? S : 1)>(scratch + h * SP) : bin_mean_rt(scratch + h * SP, Sx);
            const double r = (h == lane ? f_pre : __ldg(gf + t0 + h)) - m;
            terms[h] = (h == lane ? w_pre : __ldg(gw_ + t0 + h)) * (r * r);
        }
        __syncwarp();
        for (int q = 0; q < s1 - s0; ++q) {
            const int b0 = q * TSW, nq = min(TSW, nts - b0);
            double acc = 0.0;
            for (int h = lane; h < nq; h += 32) acc += terms[b0 + h];
            acc = warp_sum(acc);
            if (lane == 0) sm.part[s0 + q] = acc;
        }
    }
    __syncthreads();
    if (warp == 0) {
        double ssum = 0.0;
        for (int c = lane; c < nsl; c += 32) ssum += sm.part[c];
        ssum = warp_sum(ssum);
        if (lane == 0) trace_max(ws, 15);  // slot partials of a direct block folded
        const double lp = ln_prior_w0();
        if (lane == 0) {
            finish_walker(sc, cw, W, oot_w0() + ssum, lp, lnprob, loglike);
            trace_max(ws, 7);  // walker of a direct block finished
        }
    }
}

// ---- stage 2 for ensembles of about one resident wave: queue workers and direct blocks in one launch ----
// Blocks [0, nq) (one resident wave, dispatched first) are eval_slots workers on the queue, which stage 1
// filled only with walkers of more than `direct_cap` slots (slow or grazing walkers; every walker of a long
// light curve).  The queue is dealt to them block after block, so with a short queue most of them find
// nothing and exit at once.  Block nq + cw then takes a free SM slot and looks at walker cw: if stage 1 left
// it a short in-transit range, the block evaluates it on the spot (one pass per warp, partials folded in
// shared memory) with two dependent global round trips (range + constants, then the pass's light-curve
// lines) where the queue needs eight; if the walker was queued or finished, the block exits.  The two kinds
// of block do not communicate; results have the same bits either way.
template <int S, int THREADS, int MINB, int PATH>
__global__ void __launch_bounds__(THREADS, MINB)
eval_mixed_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors, EvalWs ws,
                  int nsm, int nq, double* __restrict__ lnprob, double* __restrict__ loglike, SampleCtx sc) {
    constexpr int NWARP = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if ((int)blockIdx.x < nq) {
        eval_slots_body<S, THREADS, PATH, true>(*reinterpret_cast<EvalSmem<THREADS>*>(smem_raw), nq, (int)blockIdx.x, lc, params, W, priors, ws,
                                                nsm, lnprob, loglike, sc);
        return;
    }
    DirectSmem<THREADS>& sm = *reinterpret_cast<DirectSmem<THREADS>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cw = (int)blockIdx.x - nq;
    const int cand = cw / W;
    // the finishing warp's row of the prior table does not depend on stage 1: load it first
    double pr_lo = 0.0, pr_hi = 0.0, pr_kind = 0.0, pr_mu = 0.0, pr_sg = 1.0;
    if (priors && warp == 0 && lane < 6) {
        const double* row = priors + (long long)cand * 36 + lane * 6;
        pr_lo = __ldg(row); pr_hi = __ldg(row + 1); pr_kind = __ldg(row + 2); pr_mu = __ldg(row + 3); pr_sg = __ldg(row + 4);
    }
    pdl_wait();  // stage 1 complete: ranges, constants and out-of-transit sums visible
    if (tid == 0) { trace_min(ws, 12); trace_max(ws, 13); }  // first / last direct block started
    const int ilo = __ldcg(ws.rlo + cw), ihi = __ldcg(ws.rhi + cw);
    int pps;
    const int nsl = slots_for(ihi - ilo, ws.tsw, ws.maxslots, pps);
    if (nsl == 0 || pps > 1 || nsl > ws.direct_cap) return;  // finished by stage 1, or on the queue
    {   // the walker's constants into this warp's shared-memory copy
        const double* src = reinterpret_cast<const double*>(ws.wcs + cw);
        double* dst = reinterpret_cast<double*>(&sm.wc[warp]);
        for (int i = lane; i < (int)(sizeof(WalkerConsts) / 8); i += 32) dst[i] = __ldcg(src + i);
    }
    double oot = 0.0, x_lane = 0.0;
    if (warp == 0) {
        oot = __ldcg(ws.oot + cw);
        if (lane < 6) x_lane = __ldcg((sc.enabled ? sc.q : params) + (long long)cw * 6 + lane);
    }
    __syncwarp();
    if (lane == 0) trace_max(ws, 5);  // constants of a direct block in shared memory
    auto ln_prior_w0 = [&]() {
        if (!priors) return 0.0;
        double wall = 0.0, bump = 0.0;
        if (lane < 6) ln_prior_terms_v(x_lane, pr_lo, pr_hi, (int)pr_kind, pr_mu, pr_sg, wall, bump);
        double lp = 0.0;
        for (int n = 0; n < 6; ++n) lp = (lp - __shfl_sync(0xffffffffu, wall, n)) + __shfl_sync(0xffffffffu, bump, n);
        return lp;
    };
    direct_eval_tail<S, THREADS, PATH>(sm, lc, ws, sc, cw, cand, W, ilo, ihi, nsl, lnprob, loglike, ln_prior_w0,
                                       [&]() { return oot; });
    if (tid == 0) {  // leave the workspace idle for the next evaluation (as the queue path does)
        ws.rlo[cw] = INT_MAX;
        ws.rhi[cw] = 0;
    }
}

template <int S, int THREADS, int MINB, int PATH>
__global__ void __launch_bounds__(THREADS, MINB)
window_direct_kernel(LcView lc, const double* __restrict__ params, int W, const double* __restrict__ priors, EvalWs ws,
                     int nsm, double* __restrict__ lnprob, double* __restrict__ loglike, SampleCtx sc) {
    constexpr int NWARP = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ndirect = W * lc.ncand;
    if ((int)blockIdx.x >= ndirect) {
        // ---- queue worker (the last blocks of the grid, dispatched after the direct blocks): waits until every
        // direct block has decided what to do with its walker -- a block decides a few microseconds after it starts,
        // long before its flux rounds --, then drains the queue like eval_slots_kernel.  Once an ensemble has left
        // its initial prior-box walkers behind the queue is empty and the workers exit at once; the separate
        // eval_slots launch that used to find it empty cost ~3 us of every half-step.  Direct blocks never wait
        // for workers, and a worker only holds one of the SM's block slots, so the wait cannot deadlock.
        pdl_wait();
        if (threadIdx.x == 0) {
            int spins = 0;
            while (ld_acquire_gpu(ws.counters + 4) < ndirect && ++spins < (1 << 22)) __nanosleep(64);
        }
        __syncthreads();
        eval_slots_body<S, THREADS, PATH, false>(*reinterpret_cast<EvalSmem<THREADS>*>(smem_raw), (int)gridDim.x - ndirect,
                                                 (int)blockIdx.x - ndirect, lc, params, W, priors, ws, nsm, lnprob, loglike, sc);
        return;
    }
    DirectSmem<THREADS>& sm = *reinterpret_cast<DirectSmem<THREADS>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cw = blockIdx.x;
    const int cand = cw / W;
    const int Sx = s_of<S>(lc.S);
    const int SP = Sx | 1;
    const int TSW = ws.tsw;
    // the finishing warp's row of the prior table does not depend on the previous kernel: load it first
    double pr_lo = 0.0, pr_hi = 0.0, pr_kind = 0.0, pr_mu = 0.0, pr_sg = 1.0;
    if (priors && warp == 0 && lane < 6) {
        const double* row = priors + (long long)cand * 36 + lane * 6;
        pr_lo = __ldg(row); pr_hi = __ldg(row + 1); pr_kind = __ldg(row + 2); pr_mu = __ldg(row + 3); pr_sg = __ldg(row + 4);
    }
    const CandGeom geo = load_cand_geom(lc, cand, Sx);
    // the random draws of the proposal and the walker's own position do not depend on the previous half-step
    // (it updates the other half of the ensemble): both are ready when the grid dependency resolves
    Draw dr = Draw{};
    double s_own[kMaxDim];
    if (sc.enabled) {
        dr = propose_draw(sc, cand, cw % W);  // a pure function: every thread gets the same
        const double* s = sc.pos + ((long long)cand * sc.W + dr.i) * sc.ndim;
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d) s_own[d] = d < sc.ndim ? s[d] : 0.0;
    }
    pdl_wait();     // the complementary walkers' positions come from the previous kernel
    pdl_trigger();  // the next kernel may be scheduled; it waits for this grid to finish before reading
    if (tid == 0) trace_min(ws, 0);

    double p6[6];
    if (sc.enabled) {
        const Proposal pr = propose_from(sc, cand, dr, s_own);
        if (tid == 0) propose_store(sc, pr, cand, cw % W, W);
#pragma unroll
        for (int d = 0; d < 6; ++d) p6[d] = pr.q[d];
    } else {
#pragma unroll
        for (int d = 0; d < 6; ++d) p6[d] = __ldg(params + (long long)cw * 6 + d);
    }
    double x_lane = 0.0;  // lane d of the finishing warp keeps parameter d for the priors
#pragma unroll
    for (int d = 0; d < 6; ++d)
        if (lane == d) x_lane = p6[d];
    const WalkerConsts wcr = make_walker_consts(p6, Sx);
    int ilo, ihi;
    walker_range(lc, cand, wcr, geo, ilo, ihi);
    int pps;
    const int nsl = slots_for(ihi - ilo, TSW, ws.maxslots, pps);
    // out-of-transit chi^2: the loads go out now, the finishing warp combines them after its pass
    const window_oot oot_words(lc, cand, geo.n, warp == 0 ? ilo : 0, warp == 0 ? ihi : 0);

    // priors of the walker, one parameter per lane of warp 0, summed in the reference's order
    auto ln_prior_w0 = [&]() {
        if (!priors) return 0.0;
        double wall = 0.0, bump = 0.0;
        if (lane < 6) ln_prior_terms_v(x_lane, pr_lo, pr_hi, (int)pr_kind, pr_mu, pr_sg, wall, bump);
        double lp = 0.0;
        for (int n = 0; n < 6; ++n) lp = (lp - __shfl_sync(0xffffffffu, wall, n)) + __shfl_sync(0xffffffffu, bump, n);
        return lp;
    };

    const bool queued = nsl > 0 && (pps > 1 || nsl > ws.direct_cap);
    // decided (nothing to publish unless the walker goes onto the queue): tell the queue workers
    if (!queued && tid == 0) atomicAdd(ws.counters + 4, 1);
    if (nsl == 0) {  // never in transit
        if (warp == 0) {
            const double lp = ln_prior_w0();
            if (lane == 0) finish_walker(sc, cw, W, oot_words.value(), lp, lnprob, loglike);
        }
        if (tid == 0) trace_max(ws, 3);
        return;
    }
    if (queued) {  // too long for one block: onto the queue
        if (tid == 0) {
            if (pps > 1) atomicOr(ws.counters + 5, 1);
            sm.first = atomicAdd(ws.counters + 1, nsl);
            ws.oot[cw] = oot_words.value();
            ws.wcs[cw] = wcr;
            ws.rlo[cw] = ilo;
            ws.rhi[cw] = ihi;
            ws.nslots[cw] = nsl;
        }
        __syncthreads();
        const int first = sm.first, step = pps * TSW;
        for (int i = tid; i < nsl; i += THREADS) {
            SlotRec r;
            r.cw = cw; r.t0 = ilo + i * step; r.t1 = min(ihi, ilo + (i + 1) * step); r.first = first;
            *reinterpret_cast<int4*>(ws.slots + first + i) = *reinterpret_cast<const int4*>(&r);
        }
        __syncthreads();  // every slot record of this walker is written ...
        if (tid == 0) {
            __threadfence();  // ... and visible before the workers see the count
            atomicAdd(ws.counters + 4, 1);
            trace_max(ws, 3);
        }
        return;
    }

    // ---- direct: warp w takes the consecutive slots [w * per, (w + 1) * per) as one pass
    if (lane == 0) sm.wc[warp] = wcr;
    __syncwarp();
    direct_eval_tail<S, THREADS, PATH>(sm, lc, ws, sc, cw, cand, W, ilo, ihi, nsl, lnprob, loglike, ln_prior_w0,
                                       [&]() { return oot_words.value(); });
}

static __global__ void init_ranges_kernel(int* rlo, int* rhi, int* itemsdone, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { rlo[i] = INT_MAX; rhi[i] = 0; itemsdone[i] = 0; }
}

}  // namespace nmb
