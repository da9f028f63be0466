/* namaste_b200 -- C ABI of the B200-native Namaste single-transit likelihood path.
 *
 * Drop-in boundary for the MCMC likelihood hot path of hposborn/Namaste.  The
 * reference has no native code; its "plugin API" for this path is three Python
 * contracts.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference checkout):
 *
 *   nmb_lc_create*      <- the `lc[N,3]` argument of MonoOnlyLogProb (namaste/namaste.py:1345)
 *                          plus the loop-invariant part of MonotransitModel.get_value
 *                          (cadence = median(diff(t)), supersampling offsets; :1157-1159)
 *   nmb_lnprob*         <- MonoOnlyLogProb(params, lc, priors, model)       (:1345-1349)
 *                          = get_value (:1155-1164) -> occultquad
 *                            (namaste/Crossfield_transit.py:690-970) -> likelihood (:1347)
 *                            + MonoLnPrior (:1314-1332), for a whole ensemble [W,6]
 *   nmb_model           <- MonotransitModel.get_value(t)                     (:1155-1164)
 *   nmb_occultquad      <- Crossfield_transit.occultquad(z, p, gamma, retall=True) (:690)
 *   nmb_lnprior         <- MonoLnPrior(params, priors)                       (:1314-1332)
 *   nmb_sampler_*       <- emcee.EnsembleSampler(...).run_mcmc(...) as called at
 *                          namaste/namaste.py:590-594 (stretch move, a = 2)
 *
 * Conventions: every function returns NMB_OK (0) or a negative code and records a
 * thread-local message readable through nmb_last_error(); no C++ exception crosses
 * the boundary.  Host buffers are owned by the caller and never written unless named
 * `out`.  Device buffers behind the opaque handles are owned by the library.  A handle
 * must not be used from two threads / two streams concurrently (it carries the
 * reduction workspace).  `stream` is a cudaStream_t passed as void* (NULL = default).
 * All floating-point data is IEEE binary64 ("double"), like the reference.
 */
#ifndef NAMASTE_B200_H
#define NAMASTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMB_OK 0
#define NMB_ERR_ARG (-1)     /* bad argument (null pointer, size, dtype ...)             */
#define NMB_ERR_CUDA (-2)    /* a CUDA runtime call failed; message has the detail       */
#define NMB_ERR_PRECOND (-3) /* light curve violates the supersampling precondition      */
#define NMB_ERR_NAN (-4)     /* NaN log-probability where emcee would raise ValueError   */

/* evaluation strategy (results agree to rounding) */
#define NMB_MODE_BRUTE 0    /* every fine sample of every timestamp is visited            */
#define NMB_MODE_WINDOWED 1 /* exact: out-of-transit chi^2 from prefix sums, window only  */
#define NMB_MODE_FUSED 2    /* brute force in one launch (small problems / diagnostics)     */
/* Within a mode the library picks its kernels by ensemble size (a block per walker for small and mid-sized
 * ensembles, a load-balanced slot queue otherwise; nmb_last_kernels reports the choice).  The choice never
 * changes a result: slot boundaries and summation trees are the same on every path, so a walker's lnprob has
 * the same bits whatever ensemble, launch shape or number of GPUs it is evaluated with. */

/* GP noise-model kernels (namaste/namaste.py:279-314) */
#define NMB_GP_REAL 0  /* celerite RealTerm(log_a, log_c) + JitterTerm(log_sigma)                    */
#define NMB_GP_QUASI 1 /* RotationTerm(log_amp, log_timescale, log_period, log_factor) + JitterTerm  */

/* arithmetic type of the transit model */
#define NMB_F64 64        /* FP64, contracted multiply-adds and <= 1.5-ulp quotients / roots in the two hot cases   */
#define NMB_F64_STRICT 66 /* FP64 in the reference's own rounding sequence (most elements bit-identical to NumPy) */
#define NMB_F32 32        /* FP32 where 1e-5 on lnprob allows (geometry root, case selection), FP64 for K, E, Pi, lambda_d */
#define NMB_F32_PURE 33   /* the two hot occultation cases entirely in FP32: 1e-5 ... 1e-4 on lnprob, reported only   */

#define NMB_NPAR 6       /* tcen, b, vel, RpRs, LD1, LD2  (namaste.py:1153)                */
#define NMB_PRIOR_COLS 6 /* lo, hi, kind(0 none,1 gaussian,2 normlim,3 evans), mu, sigma, pad */

typedef struct nmb_lc nmb_lc;           /* one or more staged light curves, device resident */
typedef struct nmb_sampler nmb_sampler; /* device-resident ensemble sampler state           */

const char* nmb_last_error(void);
const char* nmb_version(void);

/* ---- light-curve staging ------------------------------------------------------------- */
/* Stage one light curve (host pointers, n points each).  `oversamp` is the supersampling
 * factor S (the reference hard-codes 10, namaste.py:1158).  Requires n >= 2, t finite and
 * increasing with t[i+1] >= t[i] + cad*(S-1)/S (otherwise the reference's sort+resize
 * binning mixes cadences; NMB_ERR_PRECOND). */
int nmb_lc_create(const double* t, const double* flux, const double* err, int64_t n, int oversamp, nmb_lc** out);

/* Stage `ncand` independent light curves.  Candidate c occupies t[c*stride .. c*stride+n[c]). */
int nmb_lc_create_batch(const double* t, const double* flux, const double* err, const int64_t* n, int64_t ncand,
                        int64_t stride, int oversamp, nmb_lc** out);
void nmb_lc_destroy(nmb_lc* lc);
int64_t nmb_lc_ncand(const nmb_lc* lc);
int64_t nmb_lc_npoints(const nmb_lc* lc, int64_t cand);
double nmb_lc_cadence(const nmb_lc* lc, int64_t cand);
int nmb_lc_oversamp(const nmb_lc* lc);

/* ---- ensemble log-probability ------------------------------------------------------------
 * params  [ncand][W][6]   device, row = (tcen, b, vel, RpRs, LD1, LD2)
 * priors  [ncand][6][6]   device prior table, may be NULL (=> no prior term, like an empty dict)
 * lnprob  [ncand][W]      device out: loglike + lnprior
 * loglike [ncand][W]      device out, may be NULL
 * Asynchronous on `stream`. */
int nmb_lnprob(nmb_lc* lc, const double* params, int64_t W, const double* priors, double* lnprob, double* loglike,
               int mode, int dtype, void* stream);

/* Same with HOST buffers: copies params (and priors) to the device, launches, copies lnprob (and loglike if
 * non-NULL) back and synchronises.  Ordinary host memory passes through the handle's pinned staging buffer; a
 * `params` / `lnprob` (+ `loglike`) buffer that is itself page-locked (cudaMallocHost, cudaHostRegister) is copied
 * to / from the device directly. */
int nmb_lnprob_host(nmb_lc* lc, const double* params, int64_t W, const double* priors, double* lnprob, double* loglike,
                    int mode, int dtype);

/* model [W][n] device out for candidate `cand`: MonotransitModel.get_value for W parameter rows */
int nmb_model(nmb_lc* lc, int64_t cand, const double* params, int64_t W, double* model, void* stream);

/* out [4][n] device: F, lambda_e, lambda_d, eta_d for z[n] (device), like retall=True */
int nmb_occultquad(const double* z, int64_t n, double rprs, double ld1, double ld2, double* out, void* stream);

/* lnprior [W] device out; params [W][npar] device; priors [npar][6] device */
int nmb_lnprior(const double* params, int64_t W, int npar, const double* priors, double* lnprior, void* stream);

/* ---- device-resident ensemble sampler (emcee 2.x stretch move, namaste/namaste.py:590-594) ----
 * The ensemble, its log-probabilities, the chain [ncand][W][T][6] and lnprobability [ncand][W][T]
 * live in HBM; a half-step is two kernel launches and no host round trip.  Random numbers are
 * Philox4x32-10 keyed by `seed` with counter (step, half, candidate, walker): a chain does not
 * depend on the launch geometry or on how walkers are split across GPUs. */
int nmb_sampler_create(nmb_lc* lc, int64_t W, const double* priors /*host [ncand][6][6] or NULL*/, uint64_t seed,
                       double a /*stretch scale, emcee default 2*/, int mode, nmb_sampler** out);
/* same sampler over the GP objective MonoLogProb (see nmb_gp_lnprob for the walker layout);
 * nwalkers must be even and >= 2 * ndim, ndim = popcount(freemask) + 6; priors [ncand][ndim][6] */
int nmb_sampler_create_gp(nmb_lc* lc, int64_t nwalkers, int kind, const double* kfixed, int freemask,
                          const double* priors, uint64_t seed, double a, nmb_sampler** out);
void nmb_sampler_destroy(nmb_sampler* s);
/* pos host [ncand][W][6]; lnp host [ncand][W] or NULL (then evaluated on the device).
 * NMB_ERR_ARG for NaN/inf parameters, NMB_ERR_NAN if an initial lnprob is NaN (emcee: ValueError). */
int nmb_sampler_set_state(nmb_sampler* s, const double* pos, const double* lnp);
/* nsteps full steps (both half-ensembles); every thin-th step is appended when store != 0.
 * Synchronises at the end; NMB_ERR_NAN if any lnprob evaluation returned NaN. */
int nmb_sampler_run(nmb_sampler* s, int64_t nsteps, int64_t thin, int store);
/* building blocks for the multi-GPU walker split: update walkers [k0, k0+nk) of one half ... */
int nmb_sampler_halfstep(nmb_sampler* s, int half, int64_t k0, int64_t nk, int store);
/* ... then close the step (RNG counter, chain column; append_all appends the whole ensemble) */
int nmb_sampler_advance(nmb_sampler* s, int stored, int append_all);
int nmb_sampler_sync(nmb_sampler* s);
/* ---- one ensemble split by walker over several GPUs of a node, one process per GPU (BASELINE config 5).
 * The reference has no counterpart (its parallelism is emcee's process pool, namaste/namaste.py:583/590); the
 * half-ensemble dependency of the stretch move it drives (:590-594) is what the exchange preserves.
 * Every rank creates the same sampler (same W, priors, seed) and sets the same state; then
 *   export : 64-byte CUDA IPC handle of this rank's ensemble replica (pos, lnp, arrival flags)
 *   attach : handles [world][64] of all ranks in rank order (the own entry is ignored); maps the peers
 *   run_split : nsteps steps; per half-step the rank updates walkers [rank*nk, (rank+1)*nk), nk = W/2/world,
 *               of the active half, stores them into every peer's replica over NVLink and exchanges arrival
 *               flags -- device side, no NCCL, no host round trip; store_all appends the whole ensemble to
 *               this rank's chain every step.  The chain has the bits of the single-GPU chain of that seed.
 * world = 1 needs no attach.  pack_half / unpack_half are the building blocks of the NCCL transport of the
 * same exchange (one all-gather of [nk][ndim + 1] rows per half-step): device buffers, sampler's stream. */
#define NMB_IPC_HANDLE_BYTES 64
int nmb_sampler_split_export(nmb_sampler* s, void* handle_out /*64 bytes*/);
int nmb_sampler_split_attach(nmb_sampler* s, int rank, int world, const void* handles /*[world][64]*/);
int nmb_sampler_run_split(nmb_sampler* s, int64_t nsteps, int store_all);
int nmb_sampler_pack_half(nmb_sampler* s, int half, int64_t k0, int64_t nk, double* out /*[nk][ndim+1]*/);
int nmb_sampler_unpack_half(nmb_sampler* s, int half, const double* in /*[W/2][ndim+1]*/);
int nmb_sampler_reset(nmb_sampler* s);
/* The current CUDA device belongs to the calling host thread (cudaGetDevice / cudaSetDevice): a thread that drives
 * a handle created by another thread selects that thread's device first.  Every entry point that takes a light-curve
 * handle refuses to run on another device than the one the handle was staged on. */
int nmb_get_device(int* device);
int nmb_set_device(int device);

/* The random stream is keyed by (seed; step, half, candidate, walker).  A catalogue of candidates stepped as several
 * batches (one sampler and CUDA stream each -- independent candidates need no lock step, and two such groups keep the
 * GPU busier than one) draws the numbers it would draw as ONE batch when every sampler is told the index of its
 * first candidate.  Call before the first step.  (Reference: one Star.RunMCMC per candidate, namaste.py:526-612.) */
int nmb_sampler_set_candidate_offset(nmb_sampler* s, int64_t cand0);
int64_t nmb_sampler_len(const nmb_sampler* s);   /* stored chain columns */
int64_t nmb_sampler_steps(const nmb_sampler* s); /* steps taken (RNG counter) */
/* what: 0 pos, 1 lnp, 2 chain [ncand][W][len][6], 3 lnprobability [ncand][W][len], 4 naccepted (int64) */
int nmb_sampler_get(nmb_sampler* s, int what, void* out);
/* Posterior reduction without moving the chain (namaste.py:597-608 + percentile summaries :643-652):
 * ncut = min(int(0.25 nsteps), 3000); good[w] = 95th percentile of lnprobability[w, ncut:] above the
 * median over walkers of the per-walker medians; pct[c][i][d] = np.percentile(samples[:, d], q[i]) over
 * chain[good, ncut:, :] with the log-probability as last column (d = ndim).  Host outputs. */
int nmb_sampler_summary(nmb_sampler* s, int64_t nsteps, const double* q, int nq, double* pct, uint8_t* good,
                        int64_t* ncut, int64_t* nsamples);
void* nmb_sampler_devptr(nmb_sampler* s, int what /*0 pos, 1 lnp*/);
void* nmb_sampler_stream(nmb_sampler* s);

/* ---- GP objective: MonoLogProb(params, lc, priors, gp) (namaste/namaste.py:1334-1339) ------------
 * = celerite GP log-likelihood of flux - MonotransitModel(t) under the kernel + MonoLnPrior, for an
 * ensemble.  A walker vector is celerite's GP.get_parameter_vector(): the FREE kernel parameters in
 * kernel order (bit i of `freemask` set = parameter i of the kernel is free), then tcen, b, vel, RpRs,
 * LD1, LD2; frozen kernel parameters are taken from `kfixed` (full kernel vector: 3 values for
 * NMB_GP_REAL, 5 for NMB_GP_QUASI).  params [ncand][W][ndim], priors [ncand][ndim][6] (nullable),
 * ndim = popcount(freemask) + 6.  nmb_gp_lnprob takes device pointers, nmb_gp_lnprob_host host ones. */
int nmb_gp_lnprob(nmb_lc* lc, const double* params, int64_t W, int kind, const double* kfixed, int freemask,
                  const double* priors, double* lnprob, double* loglike, void* stream);
int nmb_gp_lnprob_host(nmb_lc* lc, const double* params, int64_t W, int kind, const double* kfixed, int freemask,
                       const double* priors, double* lnprob, double* loglike);

/* ---- the pre-fit steps for a batch of candidates -----------------------------------------------
 * Namaste prepares every light curve with an outlier cut, two detrending passes and a window around the transit
 * before it fits (namaste/namaste.py:1109, :1143 / Example.ipynb cell 45, :550-556).  One workspace handle owns the
 * device buffers (grown on demand, kept between calls) and a stream; host arrays are [ncand][stride], candidate c
 * holding n[c] valid points.
 *   nmb_anomcut_batch     <- AnomCutDiff(flux, thresh)                       (namaste/namaste.py:1459-1467)
 *                            keep [ncand][stride] out, 1 = keep; flux must be NaN-free (nonan, :1456)
 *   nmb_flatten_batch     <- k2flatten.ReduceNoise(lc, winsize, stepsize, polydegree, niter, sigmaclip, gapthreshold)
 *                            (namaste/k2flatten.py:51-84; formwindow :19-40, dopolyfit :5-17): t shifted to start at 0,
 *                            f and e divided by the median flux; out = detrended flux, 0 where no step owns a point
 *   nmb_window_mask_batch <- the fit window of RunMCMC (namaste/namaste.py:550-556): keep = abs(t - tcen < width)
 *                            (the reference's one-sided quirk) or, two_sided != 0, |t - tcen| < width          */
typedef struct nmb_flat nmb_flat;
int nmb_flat_create(nmb_flat** out);
void nmb_flat_destroy(nmb_flat* w);
int nmb_anomcut_batch(nmb_flat* w, const double* flux, const int64_t* n, int64_t ncand, int64_t stride, double thresh,
                      uint8_t* keep);
int nmb_flatten_batch(nmb_flat* w, const double* t, const double* f, const double* e, const int64_t* n, int64_t ncand,
                      int64_t stride, double winsize, double stepsize, double gapthreshold, int deg, int niter,
                      double sigclip, double* out);
int nmb_window_mask_batch(nmb_flat* w, const double* t, const int64_t* n, int64_t ncand, int64_t stride,
                          const double* tcen, const double* width, int two_sided, uint8_t* keep);

/* ---- measurement helpers (used by bench.py) ------------------------------------------------ */
/* Sustained DFMA throughput of this GPU in TFLOP/s (2 flops per DFMA), CUDA-event timed. */
int nmb_measure_fp64_peak(double* tflops, double* ms);
/* per-kernel device timing of the two pipeline stages (CUDA events on the launching stream):
 * enable, run evaluations, then read the mean stage-1 / stage-2 kernel durations in ms (and reset) */
int nmb_stage_timing(nmb_lc* lc, int enable);
int nmb_stage_times(nmb_lc* lc, double* stage1_ms, double* stage2_ms, int64_t* count);
/* names of the stage-1 / stage-2 kernels the last nmb_lnprob evaluation on this handle launched (static
 * strings): the pipeline picks them by mode and ensemble size, and bench.py reports what actually ran */
int nmb_last_kernels(nmb_lc* lc, const char** stage1, const char** stage2);
/* diagnostics: phase envelopes (globaltimer ns) of the last evaluation; needs NMB_TRACE=1 in the environment */
int nmb_debug_trace(nmb_lc* lc, uint64_t* out16);
/* host utility: position-sensitive 128-bit checksum of `nwords` 64-bit words of a host array (one pass at memory
 * speed).  The Python layer keys its cache of staged light curves on it, so that the reference's call
 * MonoOnlyLogProb(params, lc, priors, model) with the same `lc` array (namaste/namaste.py:590, emcee args=) finds
 * the device copy again and notices in-place edits. */
int nmb_host_checksum(const void* data, int64_t nwords, uint64_t* out2);
/* number of kernels this library has launched so far in this process */
int64_t nmb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NAMASTE_B200_H */
