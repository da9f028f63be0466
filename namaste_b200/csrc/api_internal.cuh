// Internals shared by the translation units of the C-ABI library: error plumbing, the staged light-curve
// handle, and the launch helpers.  The launch helpers are templates on the supersampling factor S; each S is
// compiled in its own translation unit (inst_s*.cu, explicit instantiation) so the library builds in parallel.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/namaste_b200.h"
#include "eval_pipeline.cuh"

extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;
// live light-curve handles: a sampler that is destroyed after its light curve must not touch it
extern std::mutex g_live_mu;
extern std::set<const void*> g_live_lc;

inline int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
    return fail(NMB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
    } while (0)

constexpr int kThreads = 128;  // sweep / window block size
constexpr int kTPT = 4;        // timestamps per thread in the sweep -> 512-timestamp tiles
constexpr int kTile = kThreads * kTPT;
constexpr int kSweepChunk = 64;  // timestamps per TMA chunk of the walker-per-thread sweep
#ifndef NMB_EVAL_MINB
#define NMB_EVAL_MINB 5          // resident stage-2 blocks per SM of the default arithmetic (96 registers)
#endif
#ifndef NMB_SWEEP_MINB
#define NMB_SWEEP_MINB 4         // resident 128-walker sweep blocks per SM (124-126 registers, sixteen timestamps per trip).
                                 // With one DFMA per fine sample and the trips' terms summed at staging time, the
                                 // 65 k x 4096 sweep (S = 15) / the 500 k x 2048 one (S = 10) / the 256-candidate batch
                                 // (S = 11) take 0.379 / 0.990 / 0.156 ms at 4 blocks per SM, 0.418 / 1.153 / 0.158 at 5,
                                 // 0.428 / 1.072 / 0.160 at 6 and 0.396 / 0.945 / 0.169 at 3
                                 // (profiles/r02z3_ab_sweep_tripsums_minb.txt): four warps per sub-partition are enough to
                                 // keep the issue port busy (tools/micro/sweep_mix.cu) and with 128 registers ptxas
                                 // keeps the shared-memory loads and the logic off the DFMA stream's back.
#endif

// Per-stage device timing (nmb_stage_timing / nmb_stage_times): CUDA events around stage 1 and
// stage 2 of every evaluation on the launching stream.  The event between the two kernels keeps
// them from overlapping, so these are per-kernel durations, not a decomposition of the step time.
struct StageTimer {
    bool on = false;
    static constexpr int kRing = 64;
    cudaEvent_t e[kRing][3] = {};
    int head = 0, pending = 0;
    double acc1 = 0, acc2 = 0;
    long n = 0;
    void enable(bool v) {
        if (v && !e[0][0])
            for (auto& r : e) for (auto& x : r) cudaEventCreate(&x);
        on = v;
    }
    void mark(int i, cudaStream_t st) { if (on) cudaEventRecord(e[head][i], st); }
    void drain(int keep) {
        while (pending > keep) {
            const int idx = (head - pending + 2 * kRing) % kRing;
            cudaEventSynchronize(e[idx][2]);
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, e[idx][0], e[idx][1]);
            cudaEventElapsedTime(&b, e[idx][1], e[idx][2]);
            acc1 += a; acc2 += b; ++n;
            --pending;
        }
    }
    void done() {
        if (!on) return;
        head = (head + 1) % kRing;
        ++pending;
        drain(kRing - 1);
    }
    ~StageTimer() {
        if (e[0][0]) for (auto& r : e) for (auto& x : r) cudaEventDestroy(x);
    }
};

struct nmb_lc {
    const char* last_stage1 = "";  // kernels of the last nmb_lnprob evaluation (nmb_last_kernels)
    const char* last_stage2 = "";
    int64_t ncand = 0, nmax = 0, npad = 0;
    int S = 0;
    int gran = 64, ngran = 0;
    int device = 0;
    std::vector<int64_t> n;
    std::vector<double> cad;
    double *d_t = nullptr, *d_f = nullptr, *d_w = nullptr, *d_term = nullptr, *d_tsum = nullptr, *d_var = nullptr, *d_pre_hi = nullptr, *d_pre_lo = nullptr, *d_off = nullptr;
    long long* d_n = nullptr;
    int *d_tgrid = nullptr, *d_gcount = nullptr;
    double *d_t0 = nullptr, *d_ginv = nullptr;
    int64_t gstride = 0;
    // reduction workspace of the fused kernel (grown on demand)
    double* d_partials = nullptr;
    int* d_tickets = nullptr;
    size_t partials_cap = 0, tickets_cap = 0;
    // workspace of the two-stage pipeline (grown on demand)
    nmb::EvalWs ws{};
    size_t ws_cw = 0, ws_partials = 0;
    // staging for the host-buffer entry point
    double *h_pin = nullptr, *d_stage = nullptr;
    size_t pin_cap = 0, stage_cap = 0;
    cudaStream_t own_stream = nullptr;
    StageTimer timer;
    // One handle owns one reduction workspace.  Evaluations issued on different streams (the caller's, the
    // host-buffer entry point's own stream, a sampler's) are ordered against each other: when the stream
    // changes, an event recorded behind everything queued on the previous stream is waited for first.
    // Host entry points additionally serialise on `mu` (pinned staging is shared).
    cudaStream_t ws_last_stream = nullptr;
    bool ws_used = false;
    cudaEvent_t ws_event = nullptr;
    std::recursive_mutex mu;

    // call at the top of every evaluation that touches the workspace, with the stream it will run on
    int order_on(cudaStream_t st) {
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess || dev != device) {
            g_err = "this light curve was staged on CUDA device " + std::to_string(device) + " but the current device is " +
                    std::to_string(dev);
            return NMB_ERR_ARG;
        }
        if (ws_used && st != ws_last_stream) {
            if (!ws_event && cudaEventCreateWithFlags(&ws_event, cudaEventDisableTiming) != cudaSuccess) return NMB_ERR_CUDA;
            if (cudaEventRecord(ws_event, ws_last_stream) != cudaSuccess || cudaStreamWaitEvent(st, ws_event, 0) != cudaSuccess) {
                cudaGetLastError();
                if (cudaDeviceSynchronize() != cudaSuccess) {
                    g_err = "nmb: could not order the evaluation behind the previous stream";
                    return NMB_ERR_CUDA;
                }
            }
        }
        ws_used = true;
        ws_last_stream = st;
        return NMB_OK;
    }
    // GP objective: binned in-transit model [CW][npad] and the repacked transit parameters [CW][6]
    double *d_modelbuf = nullptr, *d_meanpar = nullptr;
    size_t modelbuf_cap = 0, meanpar_cap = 0;

    nmb::LcView view() const {
        nmb::LcView v;
        v.t = d_t; v.f = d_f; v.w = d_w; v.term = d_term; v.tsum = d_tsum; v.gran = gran; v.ngran = ngran; v.pre_hi = d_pre_hi; v.pre_lo = d_pre_lo; v.off = d_off;
        v.n = d_n; v.npad = npad; v.ncand = (int)ncand; v.S = S;
        v.tgrid = d_tgrid; v.t0 = d_t0; v.ginv = d_ginv; v.gcount = d_gcount; v.gstride = gstride;
        return v;
    }
};


template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NMB_OK;
}

constexpr int kEvalThreads = 128;

// Launch with programmatic dependent launch enabled (see pdl_wait / pdl_trigger): the kernel may be
// scheduled while its predecessor in the stream drains.  NMB_NO_PDL=1 restores plain launches.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool off = std::getenv("NMB_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}



inline int ensure_ws(nmb_lc* lc, size_t CW, size_t ntiles, cudaStream_t st) {
    nmb::EvalWs& ws = lc->ws;
    if (CW > lc->ws_cw) {
        cudaFree(ws.oot); cudaFree(ws.wcs); cudaFree(ws.swc); cudaFree(ws.rlo); cudaFree(ws.rhi); cudaFree(ws.nslots);
        cudaFree(ws.slots); cudaFree(ws.slotpart); cudaFree(ws.slotsdone); cudaFree(ws.grouptick);
        cudaFree(ws.counters); cudaFree(ws.trace);
        cudaFree(ws.partials);
        ws = nmb::EvalWs{};
        lc->ws_cw = 0; lc->ws_partials = 0;
        // timestamps per pass: as many whole bins as fit in kEvalPass supersamples; slots per walker:
        // one pass each up to a pool of 4 M slots over the ensemble (>= 64 per walker)
        ws.tsw = std::max(1, nmb::kEvalPass / lc->S);
        // slots a warp evaluates in one pass: ~270 supersamples, two slots when the bins are long
        // (measured: S = 15 prefers 2, S = 10 / 11 prefer 3-4; differences of a few per cent)
        ws.merge = lc->S >= 13 ? 2 : 3;
        if (const char* m = std::getenv("NMB_EVAL_MERGE")) ws.merge = std::max(1, std::min(nmb::kEvalMerge, std::atoi(m)));
        const size_t full = (size_t)(lc->npad + ws.tsw - 1) / ws.tsw;
        ws.maxslots = (int)std::min<size_t>(full, std::max<size_t>(64, (size_t)4194304 / CW));
        const size_t cap = CW * ws.maxslots;
        CU(cudaMalloc(&ws.oot, CW * sizeof(double)));
        CU(cudaMalloc(&ws.wcs, CW * sizeof(nmb::WalkerConsts)));
        CU(cudaMalloc(&ws.swc, CW * sizeof(nmb::SweepConsts)));
        CU(cudaMalloc(&ws.rlo, CW * sizeof(int)));
        CU(cudaMalloc(&ws.rhi, CW * sizeof(int)));
        CU(cudaMalloc(&ws.slotsdone, CW * sizeof(int)));
        CU(cudaMalloc(&ws.nslots, CW * sizeof(int)));
        CU(cudaMalloc(&ws.slots, cap * sizeof(nmb::SlotRec)));
        CU(cudaMalloc(&ws.slotpart, cap * sizeof(double)));
        CU(cudaMalloc(&ws.grouptick, CW * sizeof(int)));
        CU(cudaMalloc(&ws.counters, 8 * sizeof(int)));
        if (std::getenv("NMB_TRACE")) {
            CU(cudaMalloc(&ws.trace, 16 * sizeof(unsigned long long)));
            CU(cudaMemsetAsync(ws.trace, 0, 16 * sizeof(unsigned long long), st));
        }
        CU(cudaMemsetAsync(ws.grouptick, 0, CW * sizeof(int), st));
        CU(cudaMemsetAsync(ws.counters, 0, 8 * sizeof(int), st));
        nmb::init_ranges_kernel<<<(unsigned)((CW + 255) / 256), 256, 0, st>>>(ws.rlo, ws.rhi, ws.slotsdone, (int)CW);
        g_launches++;
        CU(cudaGetLastError());
        lc->ws_cw = CW;
    }
    if (CW * ntiles > lc->ws_partials) {
        cudaFree(ws.partials);
        ws.partials = nullptr; lc->ws_partials = 0;
        CU(cudaMalloc(&ws.partials, CW * ntiles * sizeof(double)));
        lc->ws_partials = CW * ntiles;
    }
    return NMB_OK;
}

template <int S, int MINB, int PATH>
int launch_eval_mb(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                   cudaStream_t st, const nmb::SampleCtx& sc) {
    auto kern = nmb::eval_slots_kernel<S, kEvalThreads, MINB, PATH>;
    const size_t smem = sizeof(nmb::EvalSmem<kEvalThreads>);
    static thread_local int per_sm = 0, sms = 0;
    if (!per_sm) {
        if (int rc = set_smem(kern, smem)) return rc;
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEvalThreads, smem));
        per_sm = std::max(per_sm, 1);
    }
    const int grid = sms * per_sm;  // one resident wave
    CU(launch_pdl(kern, dim3(grid), dim3(kEvalThreads), smem, st, lc->view(), params, W, priors, lc->ws, sms, lnprob,
                  loglike, sc));
    g_launches++;
    return NMB_OK;
}

template <int S>
int launch_eval(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                cudaStream_t st, const nmb::SampleCtx& sc, int dtype) {
    if (dtype == NMB_F32) return launch_eval_mb<S, 5, 2>(lc, params, W, priors, lnprob, loglike, st, sc);
    if (dtype == NMB_F32_PURE) return launch_eval_mb<S, 4, 4>(lc, params, W, priors, lnprob, loglike, st, sc);
    if (dtype == NMB_F64_STRICT) return launch_eval_mb<S, 4, 1>(lc, params, W, priors, lnprob, loglike, st, sc);
    // relaxed path: five blocks (20 warps) per SM.  It fits 96 registers with two spilled words, and the shorter
    // per-warp share of the queue outweighs them (measured: -4 % ... -7 % on every workload against four
    // blocks; six blocks spill inside the flux rounds and lose)
    return launch_eval_mb<S, NMB_EVAL_MINB, 3>(lc, params, W, priors, lnprob, loglike, st, sc);
}

// Stage 2 as one launch of queue workers + one direct block per walker (eval_mixed_kernel).
template <int S, int MINB, int PATH>
int launch_mixed_mb(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                    cudaStream_t st, const nmb::SampleCtx& sc) {
    auto kern = nmb::eval_mixed_kernel<S, kEvalThreads, MINB, PATH>;
    const size_t smem = std::max(sizeof(nmb::EvalSmem<kEvalThreads>), sizeof(nmb::DirectSmem<kEvalThreads>));
    static thread_local int sms = 0, per_sm = 0;
    if (!sms) {
        if (int rc = set_smem(kern, smem)) return rc;
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEvalThreads, smem));
        per_sm = std::max(per_sm, 1);
    }
    // queue workers: one resident wave, like eval_slots_kernel (those that find the queue empty exit at once)
    static const int nq_force = std::getenv("NMB_MIXED_NQ") ? std::max(1, std::atoi(std::getenv("NMB_MIXED_NQ"))) : 0;
    const int nq = sms * (nq_force ? nq_force : per_sm);
    const size_t CW = (size_t)lc->ncand * W;
    CU(launch_pdl(kern, dim3((unsigned)(nq + CW)), dim3(kEvalThreads), smem, st, lc->view(), params, W, priors, lc->ws, sms,
                  nq, lnprob, loglike, sc));
    g_launches++;
    return NMB_OK;
}

template <int S>
int launch_mixed(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                 cudaStream_t st, const nmb::SampleCtx& sc, int dtype) {
    if (dtype == NMB_F32) return launch_mixed_mb<S, 5, 2>(lc, params, W, priors, lnprob, loglike, st, sc);
    if (dtype == NMB_F64_STRICT) return launch_mixed_mb<S, 4, 1>(lc, params, W, priors, lnprob, loglike, st, sc);
    return launch_mixed_mb<S, NMB_EVAL_MINB, 3>(lc, params, W, priors, lnprob, loglike, st, sc);
}

// The mixed stage 2 (eval_mixed_kernel: short ranges evaluated by a block per walker, only long ranges queued) was
// the default for ensembles of 256 ... 1184 walkers until the queue's short-queue deal (three slots per pass, static
// and balanced, eval_slots_body) and the queue workers inside window_direct_kernel arrived.  Measured again on the
// K2-like light curve (profiles/r02h2_scale_mixed_vs_queue.txt; back-to-back / L2 flushed, us):
//   windowed  768 walkers: mixed 30.4 / 37.9, queue 26.1 / 31.8;  1024: 34.3 / 43.1 against 32.3 / 38.9
//   brute     512 walkers: mixed 28.6 / 35.8, queue 29.0 / 33.8;  1024: 38.2 / 48.2 against 40.0 / 46.1
// The queue wins everywhere on a flushed L2 and in windowed mode; the mixed kernel keeps 2 us on back-to-back
// brute-force evaluations of ~1000 walkers, which is not a product path.  It is therefore off by default and kept
// behind NMB_MIXED_MINCW / NMB_MIXED_MAXCW (tests/test_gpu_edges.py runs it and compares its bits).
inline bool use_mixed(const nmb_lc* lc, size_t CW) {
    static const long long lo = std::getenv("NMB_MIXED_MINCW") ? std::atoll(std::getenv("NMB_MIXED_MINCW")) : 256;
    static const long long hi = std::getenv("NMB_MIXED_MAXCW") ? std::atoll(std::getenv("NMB_MIXED_MAXCW")) : 0;
    return !lc->ws.modelbuf && (long long)CW >= lo && (long long)CW <= hi;
}

// Slots a direct block of eval_mixed_kernel takes: two per warp (six flux rounds).  A longer range in one
// block makes it the straggler of the launch while queue workers run beside it and could spread that
// walker over the whole GPU (config 1 with the mixed path forced: 40.4 us with three slots per warp, 34.7 us
// with two).  window_direct_kernel takes one full pass per warp (`merge` slots): there the queue only
// starts after the direct blocks, so fewer queued walkers is the better trade.
inline int direct_cap_slots(const nmb_lc* lc, bool queue_runs_beside) {
    static const int per_warp = std::getenv("NMB_DIRECT_SLOTS") ? std::max(1, std::atoi(std::getenv("NMB_DIRECT_SLOTS"))) : 2;
    return (kEvalThreads / 32) * (queue_runs_beside ? std::min(per_warp, lc->ws.merge) : lc->ws.merge);
}

template <int S, int THREADS, int MINB, bool SUB>
int launch_sweep_t(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                   cudaStream_t st, const nmb::SampleCtx& sc) {
    auto kern = nmb::sweep_walkers_kernel<S, THREADS, kSweepChunk, MINB, SUB>;
    const size_t smem = sizeof(nmb::WalkSmem<S, kSweepChunk>);
    static thread_local int resident = 0;
    if (!resident) {
        if (int rc = set_smem(kern, smem)) return rc;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int dev = 0, sms = 0, per_sm = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
        resident = sms * std::max(per_sm, 1);
    }
    const size_t CW = (size_t)lc->ncand * W;
    const int nwb = (W + THREADS - 1) / THREADS;
    // granules per block: many more blocks than resident slots, so that the hardware block scheduler
    // evens out the SMs (measured: 5-8 % faster than one exactly-resident wave on the 65 k and 500 k
    // light curves), but at least ~256 timestamps per block to amortise its prologue
    static const int force_gpb = std::getenv("NMB_GPB") ? std::atoi(std::getenv("NMB_GPB")) : 0;
    const long long units = (long long)lc->ncand * nwb * lc->ngran;
    const long long by_balance = std::max<long long>(1, units / (4LL * resident));
    const long long by_work = std::max<long long>(1, (256 + lc->gran - 1) / lc->gran);
    int gpb = (int)std::min(by_balance, by_work);
    if (force_gpb > 0) gpb = force_gpb;
    gpb = std::min(gpb, lc->ngran);
    const int ntb = (lc->ngran + gpb - 1) / gpb;
    dim3 grid(ntb, nwb, (unsigned)lc->ncand);
    // walker groups swept by many time blocks: the walkers' constants are derived once, by a kernel in front
    static const int pre_min = std::getenv("NMB_SWEEP_PRE_MIN") ? std::atoi(std::getenv("NMB_SWEEP_PRE_MIN")) : 16;
    lc->ws.pre_consts = (pre_min > 0 && ntb >= pre_min) ? 1 : 0;
    if (lc->ws.pre_consts) {
        CU(launch_pdl(nmb::sweep_consts_kernel, dim3((unsigned)((CW + 127) / 128)), dim3(128), 0, st, lc->view(), params, W,
                      lc->ws, sc));
        g_launches++;
    }
    CU(launch_pdl(kern, grid, dim3(THREADS), smem, st, lc->view(), params, W, gpb, priors, lc->ws, lnprob, loglike, sc));
    g_launches++;
    lc->ws.pre_consts = 0;
    return NMB_OK;
}

template <int S>
int launch_sweep(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                 cudaStream_t st, const nmb::SampleCtx& sc, int dtype) {
    const size_t CW = (size_t)lc->ncand * W;
    if (int rc = ensure_ws(lc, CW, lc->ngran, st)) return rc;
    const bool mixed = use_mixed(lc, CW) && dtype != NMB_F32_PURE;  // (the pure-FP32 variant exists on the queue path only)
    lc->ws.direct_cap = mixed ? direct_cap_slots(lc, true) : 0;
    lc->timer.mark(0, st);
    // 128-walker blocks when they alone fill the GPU, single-warp blocks (finer spread) otherwise
    static const int force_thr = std::getenv("NMB_SWEEP_THREADS") ? std::atoi(std::getenv("NMB_SWEEP_THREADS")) : 0;
    const long long blocks128 = (long long)lc->ncand * ((W + 127) / 128) * lc->ngran;
    const bool wide = force_thr ? (force_thr == 128) : (W >= 128 || blocks128 >= 4 * 148);
    // (the S = 10 loop schedules best with three blocks per SM: 0.945 against 0.990 ms on the 500 k x 2048 sweep)
    constexpr int kMinB = (S == 10 && NMB_SWEEP_MINB == 4) ? 3 : NMB_SWEEP_MINB;
    // a tail warp of 16 walkers or fewer (W = 100: four) takes the several-lanes-per-walker layout
    const int tail = W % 32;
    const bool sub = S > 0 && tail >= 1 && tail <= 16;
    int rc = !wide ? (sub ? launch_sweep_t<S, 32, 16, true>(lc, params, W, priors, lnprob, loglike, st, sc)
                          : launch_sweep_t<S, 32, 16, false>(lc, params, W, priors, lnprob, loglike, st, sc))
                   : (sub ? launch_sweep_t<S, 128, kMinB, true>(lc, params, W, priors, lnprob, loglike, st, sc)
                          : launch_sweep_t<S, 128, kMinB, false>(lc, params, W, priors, lnprob, loglike, st, sc));
    if (rc) return rc;
    lc->timer.mark(1, st);
    rc = mixed ? launch_mixed<S>(lc, params, W, priors, lnprob, loglike, st, sc, dtype)
               : launch_eval<S>(lc, params, W, priors, lnprob, loglike, st, sc, dtype);
    lc->last_stage1 = "sweep_walkers_kernel";
    lc->last_stage2 = mixed ? "eval_mixed_kernel" : "eval_slots_kernel";
    lc->timer.mark(2, st);
    lc->timer.done();
    return rc;
}

template <int S, int PATH, int THREADS>
int launch_direct_t(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                    cudaStream_t st, const nmb::SampleCtx& sc, int* resident_out) {
    auto kern = nmb::window_direct_kernel<S, THREADS, 512 / THREADS, PATH>;
    const size_t smem = std::max(sizeof(nmb::DirectSmem<THREADS>), sizeof(nmb::EvalSmem<THREADS>));
    static thread_local int resident = 0, sms = 0;
    if (!resident) {
        if (int rc = set_smem(kern, smem)) return rc;
        int dev = 0, per_sm = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
        resident = sms * std::max(per_sm, 1);
    }
    if (resident_out) { *resident_out = resident; if (!params && !sc.enabled) return NMB_OK; }
    const size_t CW = (size_t)lc->ncand * W;
    // one queue worker per SM behind the direct blocks, in the same grid (see window_direct_kernel)
    static const int work_bps = std::getenv("NMB_DIRECT_EVAL_BPS") ? std::max(1, std::atoi(std::getenv("NMB_DIRECT_EVAL_BPS"))) : 1;
    const unsigned nwork = (unsigned)(sms * work_bps);
    CU(launch_pdl(kern, dim3((unsigned)CW + nwork), dim3(THREADS), smem, st, lc->view(), params, W, priors, lc->ws, sms,
                  lnprob, loglike, sc));
    g_launches++;
    return NMB_OK;
}

// A block of four warps per walker (one slot per warp for a K2 transit).  Used while the ensemble fits one
// resident wave of such blocks: beyond that the walkers left to the queue would wait for several waves of
// direct blocks, and the two-kernel pipeline balances better (measured: 1024 walkers of BASELINE config 2,
// 58 us direct against 50 us through the queue; 512-walker half-steps 24 us against 34 us).
// *use = false tells the caller to take the queue path.
template <int S, int PATH>
int launch_direct(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                  cudaStream_t st, const nmb::SampleCtx& sc, bool* use) {
    static const int force = std::getenv("NMB_DIRECT") ? std::atoi(std::getenv("NMB_DIRECT")) : -1;
    static thread_local int resident = 0;
    if (!resident) {
        if (int rc = launch_direct_t<S, PATH, 128>(lc, nullptr, W, priors, lnprob, loglike, st, nmb::SampleCtx{}, &resident))
            return rc;
    }
    const size_t CW = (size_t)lc->ncand * W;
    *use = force >= 0 ? force != 0 : CW <= (size_t)resident;
    if (!*use) return NMB_OK;
    return launch_direct_t<S, PATH, 128>(lc, params, W, priors, lnprob, loglike, st, sc, nullptr);
}

template <int S>
int launch_window(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                  cudaStream_t st, const nmb::SampleCtx& sc, int dtype) {
    const size_t CW = (size_t)lc->ncand * W;
    if (int rc = ensure_ws(lc, CW, 0, st)) return rc;
    constexpr int kWinThreads = 64;
    lc->timer.mark(0, st);
    // small ensembles: a block per walker evaluates short ranges on the spot (window_direct_kernel); otherwise,
    // and in the GP model mode, the thread-per-walker kernel fills the queue for eval_slots_kernel
    bool direct = false;
    lc->ws.direct_cap = direct_cap_slots(lc, false);  // read by window_direct_kernel only (it never calls push_slots)
    if (!lc->ws.modelbuf && dtype != NMB_F32_PURE) {
        int rc = dtype == NMB_F32          ? launch_direct<S, 2>(lc, params, W, priors, lnprob, loglike, st, sc, &direct)
                 : dtype == NMB_F64_STRICT ? launch_direct<S, 1>(lc, params, W, priors, lnprob, loglike, st, sc, &direct)
                                           : launch_direct<S, 3>(lc, params, W, priors, lnprob, loglike, st, sc, &direct);
        if (rc) return rc;
    }
    const bool mixed = !direct && use_mixed(lc, CW) && dtype != NMB_F32_PURE;
    if (!direct) {
        lc->ws.direct_cap = mixed ? direct_cap_slots(lc, true) : 0;
        CU(launch_pdl(nmb::window_ranges_kernel<kWinThreads>, dim3((unsigned)((CW + kWinThreads - 1) / kWinThreads)),
                      dim3(kWinThreads), 0, st, lc->view(), params, W, priors, lc->ws, lnprob, loglike, sc));
        g_launches++;
    }
    lc->timer.mark(1, st);
    // (window_direct_kernel carries its own queue workers: nothing to launch behind it)
    const int rc = direct  ? NMB_OK
                   : mixed ? launch_mixed<S>(lc, params, W, priors, lnprob, loglike, st, sc, dtype)
                           : launch_eval<S>(lc, params, W, priors, lnprob, loglike, st, sc, dtype);
    lc->last_stage1 = direct ? "window_direct_kernel" : "window_ranges_kernel";
    lc->last_stage2 = direct ? "window_direct_kernel(queue workers)" : mixed ? "eval_mixed_kernel" : "eval_slots_kernel";
    lc->timer.mark(2, st);
    lc->timer.done();
    return rc;
}

template <int S>
int launch_fused(nmb_lc* lc, const double* params, int W, const double* priors, double* lnprob, double* loglike,
                 cudaStream_t st) {
    const int ntiles = (int)(lc->npad / kTile);
    long long blocks1 = (long long)ntiles * W * lc->ncand;
    int WT = (int)std::min<long long>(nmb::kWTMax, std::max<long long>(1, blocks1 / (148 * 8)));
    const int groups = (W + WT - 1) / WT;
    const size_t need_p = (size_t)lc->ncand * groups * WT * ntiles;
    const size_t need_t = (size_t)lc->ncand * groups;
    if (need_p > lc->partials_cap) {
        cudaFree(lc->d_partials);
        lc->d_partials = nullptr; lc->partials_cap = 0;
        CU(cudaMalloc(&lc->d_partials, need_p * sizeof(double)));
        lc->partials_cap = need_p;
    }
    if (need_t > lc->tickets_cap) {
        cudaFree(lc->d_tickets);
        lc->d_tickets = nullptr; lc->tickets_cap = 0;
        CU(cudaMalloc(&lc->d_tickets, need_t * sizeof(int)));
        CU(cudaMemsetAsync(lc->d_tickets, 0, need_t * sizeof(int), st));
        lc->tickets_cap = need_t;
    }
    auto kern = nmb::lnprob_sweep_kernel<S, kThreads, kTPT, false>;
    const size_t smem = sizeof(nmb::SweepSmem<S, kThreads, kTPT>);
    if (int rc = set_smem(kern, smem)) return rc;
    dim3 grid(ntiles, groups, (unsigned)lc->ncand);
    kern<<<grid, kThreads, smem, st>>>(lc->view(), params, W, WT, priors, lc->d_partials, lc->d_tickets, lnprob,
                                       loglike, nullptr, 0);
    g_launches++;
    CU(cudaGetLastError());
    return NMB_OK;
}

template <int S>
int launch_model(nmb_lc* lc, int cand, const double* params, int W, double* model, cudaStream_t st) {
    const int ntiles = (int)((lc->n[cand] + kTile - 1) / kTile);
    const int WT = 1;
    auto kern = nmb::lnprob_sweep_kernel<S, kThreads, kTPT, true>;
    const size_t smem = sizeof(nmb::SweepSmem<S, kThreads, kTPT>);
    if (int rc = set_smem(kern, smem)) return rc;
    dim3 grid(ntiles, W, 1);
    kern<<<grid, kThreads, smem, st>>>(lc->view(), params, W, WT, nullptr, nullptr, nullptr,
                                       nullptr, nullptr, model, cand);
    g_launches++;
    CU(cudaGetLastError());
    return NMB_OK;
}

#define NMB_DISPATCH_S(S_, CALL)                  \
    switch (S_) {                                 \
        case 1: { constexpr int SS = 1; CALL; } break;   \
        case 10: { constexpr int SS = 10; CALL; } break; \
        case 11: { constexpr int SS = 11; CALL; } break; \
        case 15: { constexpr int SS = 15; CALL; } break; \
        default: { constexpr int SS = 0; CALL; } break;  \
    }


// ---- one translation unit per supersampling factor ------------------------------------------------------
#define NMB_FOR_EACH_S(X) X(0) X(1) X(10) X(11) X(15)
#define NMB_LAUNCH_SIGNATURES(S_, KW)                                                                                   \
    KW template int launch_sweep<S_>(nmb_lc*, const double*, int, const double*, double*, double*, cudaStream_t,        \
                                     const nmb::SampleCtx&, int);                                                       \
    KW template int launch_window<S_>(nmb_lc*, const double*, int, const double*, double*, double*, cudaStream_t,       \
                                      const nmb::SampleCtx&, int);                                                      \
    KW template int launch_fused<S_>(nmb_lc*, const double*, int, const double*, double*, double*, cudaStream_t);       \
    KW template int launch_model<S_>(nmb_lc*, int, const double*, int, double*, cudaStream_t);
#ifdef NMB_INSTANTIATE_S
NMB_LAUNCH_SIGNATURES(NMB_INSTANTIATE_S, )
#else
#define NMB_EXTERN_S(S_) NMB_LAUNCH_SIGNATURES(S_, extern)
NMB_FOR_EACH_S(NMB_EXTERN_S)
#undef NMB_EXTERN_S
#endif

