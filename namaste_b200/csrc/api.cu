// C-ABI entry points (include/namaste_b200.h): light-curve staging, launches, error plumbing.
#include "api_internal.cuh"
// kernels only the entry points launch (not templates on S): defined in this translation unit alone
#include "gp_kernels.cuh"
#include "flatten_kernels.cuh"
#include "summary_kernels.cuh"
#include "split_kernels.cuh"

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::mutex g_live_mu;
std::set<const void*> g_live_lc;

extern "C" const char* nmb_last_error(void) { return g_err.c_str(); }
extern "C" const char* nmb_version(void) { return "namaste_b200 0.1 (sm_100a)"; }
extern "C" int64_t nmb_launch_count(void) { return g_launches.load(); }

extern "C" void nmb_lc_destroy(nmb_lc* lc) {
    if (!lc) return;
    {
        std::lock_guard<std::mutex> g(g_live_mu);
        g_live_lc.erase(lc);
    }
    cudaFree(lc->d_t); cudaFree(lc->d_f); cudaFree(lc->d_w); cudaFree(lc->d_term); cudaFree(lc->d_tsum); cudaFree(lc->d_var); cudaFree(lc->d_modelbuf); cudaFree(lc->d_meanpar); cudaFree(lc->d_pre_hi); cudaFree(lc->d_pre_lo);
    cudaFree(lc->d_tgrid); cudaFree(lc->d_gcount); cudaFree(lc->d_t0); cudaFree(lc->d_ginv);
    cudaFree(lc->d_off); cudaFree(lc->d_n); cudaFree(lc->d_partials); cudaFree(lc->d_tickets); cudaFree(lc->d_stage);
    cudaFree(lc->ws.partials); cudaFree(lc->ws.oot); cudaFree(lc->ws.wcs); cudaFree(lc->ws.swc); cudaFree(lc->ws.rlo); cudaFree(lc->ws.rhi);
    cudaFree(lc->ws.nslots); cudaFree(lc->ws.slots); cudaFree(lc->ws.slotpart); cudaFree(lc->ws.slotsdone);
    cudaFree(lc->ws.grouptick); cudaFree(lc->ws.counters); cudaFree(lc->ws.trace);
    if (lc->h_pin) cudaFreeHost(lc->h_pin);
    if (lc->own_stream) cudaStreamDestroy(lc->own_stream);
    if (lc->ws_event) cudaEventDestroy(lc->ws_event);
    delete lc;
}

extern "C" int nmb_lc_create_batch(const double* t, const double* flux, const double* err, const int64_t* n,
                                   int64_t ncand, int64_t stride, int oversamp, nmb_lc** out) {
    if (!t || !flux || !err || !n || !out) return fail(NMB_ERR_ARG, "nmb_lc_create: null pointer");
    if (ncand < 1) return fail(NMB_ERR_ARG, "nmb_lc_create: ncand < 1");
    if (oversamp < 1 || oversamp > nmb::kMaxS) return fail(NMB_ERR_ARG, "nmb_lc_create: oversamp must be in 1..32");
    const int S = oversamp;
    // np.arange(0, 1.0, 1/S) must have exactly S elements (namaste.py:1159)
    if ((int64_t)std::ceil(1.0 / (1.0 / S)) != S)
        return fail(NMB_ERR_ARG, "nmb_lc_create: arange(0,1,1/oversamp) does not have oversamp elements");
    int64_t nmax = 0;
    for (int64_t c = 0; c < ncand; ++c) {
        if (n[c] < 2) return fail(NMB_ERR_ARG, "nmb_lc_create: each light curve needs >= 2 points");
        if (n[c] > stride) return fail(NMB_ERR_ARG, "nmb_lc_create: n[c] > stride");
        nmax = std::max(nmax, n[c]);
    }
    // reduction granule of the brute-force sweep: a power of two >= 64 timestamps chosen from the
    // light-curve length alone (<= 512 granules), so partial sums do not depend on the launch shape
    int gran = kSweepChunk;
    while ((nmax + gran - 1) / gran > 512) gran *= 2;
    const int64_t padto = std::max<int64_t>(kTile, gran);
    const int64_t npad = (nmax + padto - 1) / padto * padto;
    std::vector<double> ht(ncand * npad, 0.0), hf(ncand * npad, 1.0), hw(ncand * npad, 0.0), hterm(ncand * npad, 0.0), hvar(ncand * npad, 1.0);
    std::vector<double> hph(ncand * (npad + 1), 0.0), hpl(ncand * (npad + 1), 0.0), hoff(ncand * S, 0.0);
    std::vector<double> cads(ncand);
    std::vector<double> d;
    for (int64_t c = 0; c < ncand; ++c) {
        const double* tc = t + c * stride;
        const double* fc = flux + c * stride;
        const double* ec = err + c * stride;
        const int64_t nc = n[c];
        d.resize(nc - 1);
        for (int64_t i = 0; i < nc; ++i)
            if (!std::isfinite(tc[i])) return fail(NMB_ERR_ARG, "nmb_lc_create: non-finite time stamp");
        for (int64_t i = 0; i + 1 < nc; ++i) d[i] = tc[i + 1] - tc[i];
        // np.median(np.diff(t)): mean of the two middle order statistics for an even count
        const int64_t m = nc - 1;
        std::nth_element(d.begin(), d.begin() + m / 2, d.end());
        double cad = d[m / 2];
        if (m % 2 == 0) {
            const double lower = *std::max_element(d.begin(), d.begin() + m / 2);
            cad = (lower + cad) / 2.0;
        }
        cads[c] = cad;
        const double step = 1.0 / S;
        for (int k = 0; k < S; ++k) hoff[c * S + k] = cad * (k * step);
        // precondition for "sort + resize(N, S)" == per-timestamp bins
        const double span = hoff[c * S + S - 1];
        for (int64_t i = 0; i + 1 < nc; ++i) {
            if (!(tc[i + 1] >= tc[i] + span)) {
                char buf[200];
                snprintf(buf, sizeof buf,
                         "light curve %lld: t[%lld+1] - t[%lld] is smaller than cad*(S-1)/S; the reference's "
                         "supersampling bins would mix cadences", (long long)c, (long long)i, (long long)i);
                return fail(NMB_ERR_PRECOND, buf);
            }
        }
        double ph = 0.0, pl = 0.0;  // double-double running sum of w*(f-1)^2
        for (int64_t i = 0; i < nc; ++i) {
            ht[c * npad + i] = tc[i];
            hf[c * npad + i] = fc[i];
            const double w = -2 * std::pow(ec[i], -2.0);  // -2*lc[:,2]**-2   (namaste.py:1347)
            hw[c * npad + i] = w;
            hph[c * (npad + 1) + i] = ph;
            hpl[c * (npad + 1) + i] = pl;
            const double r = fc[i] - 1.0;
            const double term = w * (r * r);
            hterm[c * npad + i] = term;
            hvar[c * npad + i] = ec[i] * ec[i];  // yerr**2 (celerite GP.compute)
            const double s = ph + term;
            const double bb = s - ph;
            const double e = (ph - (s - bb)) + (term - bb);
            pl += e;
            ph = s;
            // renormalise
            const double s2 = ph + pl;
            pl = pl - (s2 - ph);
            ph = s2;
        }
        for (int64_t i = nc; i <= npad; ++i) {
            hph[c * (npad + 1) + i] = ph;
            hpl[c * (npad + 1) + i] = pl;
        }
        for (int64_t i = nc; i < npad; ++i) ht[c * npad + i] = tc[nc - 1];
    }

    // uniform time grid (cell = one cadence, coarser if the light curve has huge gaps)
    std::vector<double> ht0(ncand), hginv(ncand);
    std::vector<int> hgcount(ncand);
    std::vector<double> gwidth(ncand);
    int64_t gstride = 0;
    for (int64_t c = 0; c < ncand; ++c) {
        const double* tc = t + c * stride;
        const int64_t nc = n[c];
        const double range = tc[nc - 1] - tc[0];
        double g = cads[c];
        if (!(g > 0.0) || range / g > 4.0 * (double)nc + 16.0) g = range / (4.0 * (double)nc);
        if (!(g > 0.0)) g = 1.0;
        gwidth[c] = g;
        ht0[c] = tc[0];
        hginv[c] = 1.0 / g;
        hgcount[c] = (int)std::floor(range / g) + 2;
        gstride = std::max<int64_t>(gstride, hgcount[c] + 1);
    }
    std::vector<int> hgrid(ncand * gstride, 0);
    for (int64_t c = 0; c < ncand; ++c) {
        const double* tc = t + c * stride;
        const int64_t nc = n[c];
        int64_t i = 0;
        for (int j = 0; j <= hgcount[c]; ++j) {
            const double edge = ht0[c] + j * gwidth[c];
            while (i < nc && tc[i] < edge) ++i;
            hgrid[c * gstride + j] = (int)i;
        }
    }

    nmb_lc* lc = new nmb_lc;
    lc->ncand = ncand; lc->nmax = nmax; lc->npad = npad; lc->S = S; lc->gstride = gstride;
    lc->gran = gran; lc->ngran = (int)(npad / gran);
    lc->n.assign(n, n + ncand);
    lc->cad = cads;
    cudaGetDevice(&lc->device);
    std::vector<long long> hn(n, n + ncand);
    // the terms of each trip of the brute-force sweep (kSweepTrip consecutive timestamps), added in timestamp order
    std::vector<double> htsum(ncand * (npad / nmb::kSweepTrip), 0.0);
    for (int64_t c = 0; c < ncand; ++c)
        for (int64_t k = 0; k < npad / nmb::kSweepTrip; ++k) {
            double acc = 0.0;
            for (int j = 0; j < nmb::kSweepTrip; ++j) acc += hterm[c * npad + k * nmb::kSweepTrip + j];
            htsum[c * (npad / nmb::kSweepTrip) + k] = acc;
        }
    auto up = [&](double** dst, const std::vector<double>& src) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, src.size() * sizeof(double));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaSuccess;
    if ((e = up(&lc->d_t, ht)) != cudaSuccess || (e = up(&lc->d_f, hf)) != cudaSuccess ||
        (e = up(&lc->d_w, hw)) != cudaSuccess || (e = up(&lc->d_term, hterm)) != cudaSuccess || (e = up(&lc->d_tsum, htsum)) != cudaSuccess || (e = up(&lc->d_var, hvar)) != cudaSuccess || (e = up(&lc->d_pre_hi, hph)) != cudaSuccess ||
        (e = up(&lc->d_pre_lo, hpl)) != cudaSuccess || (e = up(&lc->d_off, hoff)) != cudaSuccess ||
        (e = up(&lc->d_t0, ht0)) != cudaSuccess || (e = up(&lc->d_ginv, hginv)) != cudaSuccess ||
        (e = cudaMalloc(&lc->d_tgrid, hgrid.size() * sizeof(int))) != cudaSuccess ||
        (e = cudaMemcpy(lc->d_tgrid, hgrid.data(), hgrid.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&lc->d_gcount, ncand * sizeof(int))) != cudaSuccess ||
        (e = cudaMemcpy(lc->d_gcount, hgcount.data(), ncand * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&lc->d_n, ncand * sizeof(long long))) != cudaSuccess ||
        (e = cudaMemcpy(lc->d_n, hn.data(), ncand * sizeof(long long), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&lc->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        nmb_lc_destroy(lc);
        return cuda_fail(e, "nmb_lc_create: staging");
    }
    {
        std::lock_guard<std::mutex> g(g_live_mu);
        g_live_lc.insert(lc);
    }
    *out = lc;
    return NMB_OK;
}

extern "C" int nmb_lc_create(const double* t, const double* flux, const double* err, int64_t n, int oversamp,
                             nmb_lc** out) {
    return nmb_lc_create_batch(t, flux, err, &n, 1, n, oversamp, out);
}

extern "C" int64_t nmb_lc_ncand(const nmb_lc* lc) { return lc ? lc->ncand : 0; }
extern "C" int64_t nmb_lc_npoints(const nmb_lc* lc, int64_t c) { return (lc && c >= 0 && c < lc->ncand) ? lc->n[c] : -1; }
extern "C" double nmb_lc_cadence(const nmb_lc* lc, int64_t c) { return (lc && c >= 0 && c < lc->ncand) ? lc->cad[c] : NAN; }
extern "C" int nmb_lc_oversamp(const nmb_lc* lc) { return lc ? lc->S : 0; }


static int lnprob_impl(nmb_lc* lc, const double* params, int64_t W, const double* priors, double* lnprob,
                       double* loglike, int mode, int dtype, void* stream, const nmb::SampleCtx& sc) {
    if (!lc || !params || (!lnprob && !sc.enabled)) return fail(NMB_ERR_ARG, "nmb_lnprob: null pointer");
    if (W < 1 || W > (1 << 24)) return fail(NMB_ERR_ARG, "nmb_lnprob: W out of range");
    if (dtype != NMB_F64 && dtype != NMB_F32 && dtype != NMB_F64_STRICT && dtype != NMB_F32_PURE)
        return fail(NMB_ERR_ARG, "nmb_lnprob: dtype must be NMB_F64, NMB_F64_STRICT, NMB_F32 or NMB_F32_PURE");
    if ((dtype == NMB_F32 || dtype == NMB_F32_PURE) && mode == NMB_MODE_FUSED)
        return fail(NMB_ERR_ARG, "nmb_lnprob: NMB_MODE_FUSED is FP64 only");
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::recursive_mutex> lock(lc->mu);
    if (int rc0 = lc->order_on(st)) return rc0;
    int rc = NMB_OK;
    if (mode == NMB_MODE_BRUTE) {
        NMB_DISPATCH_S(lc->S, rc = launch_sweep<SS>(lc, params, (int)W, priors, lnprob, loglike, st, sc, dtype));
    } else if (mode == NMB_MODE_WINDOWED) {
        NMB_DISPATCH_S(lc->S, rc = launch_window<SS>(lc, params, (int)W, priors, lnprob, loglike, st, sc, dtype));
    } else if (mode == NMB_MODE_FUSED) {
        NMB_DISPATCH_S(lc->S, rc = launch_fused<SS>(lc, params, (int)W, priors, lnprob, loglike, st));
    } else {
        return fail(NMB_ERR_ARG, "nmb_lnprob: unknown mode");
    }
    return rc;
}

extern "C" int nmb_lnprob(nmb_lc* lc, const double* params, int64_t W, const double* priors, double* lnprob,
                          double* loglike, int mode, int dtype, void* stream) {
    nmb::SampleCtx sc{};
    return lnprob_impl(lc, params, W, priors, lnprob, loglike, mode, dtype, stream, sc);
}

// true for page-locked host memory (cudaMallocHost / cudaHostRegister / torch pin_memory): such a buffer is copied
// to or from the device directly, without the handle's staging buffer in between
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int nmb_lnprob_host(nmb_lc* lc, const double* params, int64_t W, const double* priors, double* lnprob,
                               double* loglike, int mode, int dtype) {
    if (!lc || !params || !lnprob) return fail(NMB_ERR_ARG, "nmb_lnprob_host: null pointer");
    if (W < 1) return fail(NMB_ERR_ARG, "nmb_lnprob_host: W < 1");
    std::lock_guard<std::recursive_mutex> lock(lc->mu);  // pinned staging and workspace: one caller at a time
    const size_t C = (size_t)lc->ncand;
    const size_t np = C * W * 6, npri = priors ? C * 36 : 0, nout = C * W * (loglike ? 2 : 1);
    const size_t total = np + npri + nout;
    if (total > lc->pin_cap) {
        if (lc->h_pin) cudaFreeHost(lc->h_pin);
        cudaFree(lc->d_stage);
        lc->h_pin = nullptr; lc->d_stage = nullptr; lc->pin_cap = 0;
        CU(cudaMallocHost(&lc->h_pin, total * sizeof(double)));
        CU(cudaMalloc(&lc->d_stage, total * sizeof(double)));
        lc->pin_cap = total;
    }
    cudaStream_t st = lc->own_stream;
    // walker positions: straight from the caller's buffer when it is page-locked, through the staging buffer otherwise
    const bool p_pinned = host_is_pinned(params);
    const bool o_pinned = host_is_pinned(lnprob) && (!loglike || host_is_pinned(loglike));
    if (p_pinned) {
        CU(cudaMemcpyAsync(lc->d_stage, params, np * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        std::memcpy(lc->h_pin, params, np * sizeof(double));
        CU(cudaMemcpyAsync(lc->d_stage, lc->h_pin, np * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (priors) {
        std::memcpy(lc->h_pin + np, priors, npri * sizeof(double));
        CU(cudaMemcpyAsync(lc->d_stage + np, lc->h_pin + np, npri * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    double* d_out = lc->d_stage + np + npri;
    int rc = nmb_lnprob(lc, lc->d_stage, W, priors ? lc->d_stage + np : nullptr, d_out,
                        loglike ? d_out + C * W : nullptr, mode, dtype, st);
    if (rc) return rc;
    if (o_pinned) {
        CU(cudaMemcpyAsync(lnprob, d_out, C * W * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (loglike) CU(cudaMemcpyAsync(loglike, d_out + C * W, C * W * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return NMB_OK;
    }
    CU(cudaMemcpyAsync(lc->h_pin + np + npri, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(lnprob, lc->h_pin + np + npri, C * W * sizeof(double));
    if (loglike) std::memcpy(loglike, lc->h_pin + np + npri + C * W, C * W * sizeof(double));
    return NMB_OK;
}

// ---- GP noise-model objective (MonoLogProb, namaste.py:1334-1339) -----------------------------------
static int gp_lnprob_impl(nmb_lc* lc, const double* params, int64_t W, int kind, const double* kfixed, int freemask,
                          const double* priors, double* lnprob, double* loglike, void* stream,
                          const nmb::SampleCtx& gsc) {
    if (!lc || !params || !kfixed || (!lnprob && !gsc.enabled)) return fail(NMB_ERR_ARG, "nmb_gp_lnprob: null pointer");
    if (W < 1 || W > (1 << 24)) return fail(NMB_ERR_ARG, "nmb_gp_lnprob: W out of range");
    if (kind != NMB_GP_REAL && kind != NMB_GP_QUASI) return fail(NMB_ERR_ARG, "nmb_gp_lnprob: unknown kernel kind");
    nmb::GpSpec gs{};
    gs.kind = kind == NMB_GP_REAL ? nmb::kGpReal : nmb::kGpQuasi;
    gs.nk = kind == NMB_GP_REAL ? 3 : 5;
    if (freemask < 0 || freemask >= (1 << gs.nk)) return fail(NMB_ERR_ARG, "nmb_gp_lnprob: free mask names parameters the kernel does not have");
    gs.freemask = freemask;
    for (int i = 0; i < gs.nk; ++i) {
        gs.kfixed[i] = kfixed[i];
        gs.nfree += (freemask >> i) & 1;
    }
    gs.ndim = gs.nfree + 6;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::recursive_mutex> lock(lc->mu);
    if (int rc0 = lc->order_on(st)) return rc0;
    const size_t CW = (size_t)lc->ncand * W;
    if (int rc = ensure_ws(lc, CW, 0, st)) return rc;
    const size_t need = CW * (size_t)lc->npad;
    if (need > lc->modelbuf_cap) {
        if (need > ((size_t)1 << 31)) return fail(NMB_ERR_ARG, "nmb_gp_lnprob: ensemble x light curve exceeds the 16 GiB model buffer");
        cudaFree(lc->d_modelbuf);
        lc->d_modelbuf = nullptr; lc->modelbuf_cap = 0;
        CU(cudaMalloc(&lc->d_modelbuf, need * sizeof(double)));
        lc->modelbuf_cap = need;
    }
    if (CW * 6 > lc->meanpar_cap) {
        cudaFree(lc->d_meanpar);
        lc->d_meanpar = nullptr; lc->meanpar_cap = 0;
        CU(cudaMalloc(&lc->d_meanpar, CW * 6 * sizeof(double)));
        lc->meanpar_cap = CW * 6;
    }
    nmb::gp_mean_params_kernel<<<(unsigned)((CW * 6 + 255) / 256), 256, 0, st>>>(params, (long long)CW, gs.ndim, gs.nfree,
                                                                                lc->d_meanpar);
    g_launches++;
    CU(cudaGetLastError());
    // stages 1 and 2 in model mode: ranges + binned in-transit model, nobody finishes walkers
    lc->ws.modelbuf = lc->d_modelbuf;
    nmb::SampleCtx sc{};
    int rc = NMB_OK;
    NMB_DISPATCH_S(lc->S, rc = launch_window<SS>(lc, lc->d_meanpar, (int)W, nullptr, nullptr, nullptr, st, sc, NMB_F64));
    lc->ws.modelbuf = nullptr;
    if (rc) return rc;
    nmb::EvalWs ws = lc->ws;
    ws.modelbuf = lc->d_modelbuf;
    const unsigned blocks = (unsigned)((CW + 31) / 32);
    if (gs.kind == nmb::kGpReal) {
        CU(launch_pdl(nmb::gp_solve_kernel<1, 0>, dim3(blocks), dim3(32), 0, st, lc->view(), (const double*)lc->d_var, params,
                      (int)W, gs, priors, ws, lnprob, loglike, gsc));
    } else {
        CU(launch_pdl(nmb::gp_solve_kernel<1, 1>, dim3(blocks), dim3(32), 0, st, lc->view(), (const double*)lc->d_var, params,
                      (int)W, gs, priors, ws, lnprob, loglike, gsc));
    }
    g_launches++;
    return NMB_OK;
}

extern "C" int nmb_gp_lnprob(nmb_lc* lc, const double* params, int64_t W, int kind, const double* kfixed,
                             int freemask, const double* priors, double* lnprob, double* loglike, void* stream) {
    nmb::SampleCtx off{};
    return gp_lnprob_impl(lc, params, W, kind, kfixed, freemask, priors, lnprob, loglike, stream, off);
}

extern "C" int nmb_gp_lnprob_host(nmb_lc* lc, const double* params, int64_t W, int kind, const double* kfixed,
                                  int freemask, const double* priors, double* lnprob, double* loglike) {
    if (!lc || !params || !kfixed || !lnprob) return fail(NMB_ERR_ARG, "nmb_gp_lnprob_host: null pointer");
    if (W < 1) return fail(NMB_ERR_ARG, "nmb_gp_lnprob_host: W < 1");
    std::lock_guard<std::recursive_mutex> lock(lc->mu);
    const int nk = kind == NMB_GP_REAL ? 3 : 5;
    int nfree = 0;
    for (int i = 0; i < nk; ++i) nfree += (freemask >> i) & 1;
    const int ndim = nfree + 6;
    const size_t C = (size_t)lc->ncand;
    const size_t np = C * W * ndim, npri = priors ? C * ndim * 6 : 0, nout = C * W * (loglike ? 2 : 1);
    const size_t total = np + npri + nout;
    if (total > lc->pin_cap) {
        if (lc->h_pin) cudaFreeHost(lc->h_pin);
        cudaFree(lc->d_stage);
        lc->h_pin = nullptr; lc->d_stage = nullptr; lc->pin_cap = 0;
        CU(cudaMallocHost(&lc->h_pin, total * sizeof(double)));
        CU(cudaMalloc(&lc->d_stage, total * sizeof(double)));
        lc->pin_cap = total;
    }
    cudaStream_t st = lc->own_stream;
    std::memcpy(lc->h_pin, params, np * sizeof(double));
    if (priors) std::memcpy(lc->h_pin + np, priors, npri * sizeof(double));
    CU(cudaMemcpyAsync(lc->d_stage, lc->h_pin, (np + npri) * sizeof(double), cudaMemcpyHostToDevice, st));
    double* d_out = lc->d_stage + np + npri;
    int rc = nmb_gp_lnprob(lc, lc->d_stage, W, kind, kfixed, freemask, priors ? lc->d_stage + np : nullptr, d_out,
                           loglike ? d_out + C * W : nullptr, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(lc->h_pin + np + npri, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(lnprob, lc->h_pin + np + npri, C * W * sizeof(double));
    if (loglike) std::memcpy(loglike, lc->h_pin + np + npri + C * W, C * W * sizeof(double));
    return NMB_OK;
}

extern "C" int nmb_model(nmb_lc* lc, int64_t cand, const double* params, int64_t W, double* model, void* stream) {
    if (!lc || !params || !model) return fail(NMB_ERR_ARG, "nmb_model: null pointer");
    if (cand < 0 || cand >= lc->ncand) return fail(NMB_ERR_ARG, "nmb_model: candidate index out of range");
    if (W < 1 || W > 65535) return fail(NMB_ERR_ARG, "nmb_model: W must be in 1..65535");
    int rc = NMB_OK;
    NMB_DISPATCH_S(lc->S, rc = launch_model<SS>(lc, (int)cand, params, (int)W, model, (cudaStream_t)stream));
    return rc;
}

extern "C" int nmb_occultquad(const double* z, int64_t n, double rprs, double ld1, double ld2, double* out,
                              void* stream) {
    if (!z || !out) return fail(NMB_ERR_ARG, "nmb_occultquad: null pointer");
    if (n < 0) return fail(NMB_ERR_ARG, "nmb_occultquad: n < 0");
    if (n == 0) return NMB_OK;
    const int threads = 128;
    const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 148 * 16);
    nmb::occultquad_array_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(z, n, rprs, ld1, ld2, out);
    g_launches++;
    CU(cudaGetLastError());
    return NMB_OK;
}

extern "C" int nmb_lnprior(const double* params, int64_t W, int npar, const double* priors, double* lnprior,
                           void* stream) {
    if (!params || !priors || !lnprior) return fail(NMB_ERR_ARG, "nmb_lnprior: null pointer");
    if (W < 1 || npar < 1) return fail(NMB_ERR_ARG, "nmb_lnprior: bad sizes");
    nmb::lnprior_kernel<<<(unsigned)((W + 127) / 128), 128, 0, (cudaStream_t)stream>>>(params, W, npar, priors, lnprior);
    g_launches++;
    CU(cudaGetLastError());
    return NMB_OK;
}

// Diagnostics (NMB_TRACE=1): phase envelopes of the last evaluation in globaltimer nanoseconds.
// out[0]/out[8] earliest block start of stage 1/2, other slots the latest end of a phase; the
// slots are re-armed for the next evaluation.
extern "C" int nmb_debug_trace(nmb_lc* lc, uint64_t* out) {
    if (!lc || !out) return fail(NMB_ERR_ARG, "nmb_debug_trace: null pointer");
    if (!lc->ws.trace) return fail(NMB_ERR_ARG, "nmb_debug_trace: tracing is off (set NMB_TRACE=1 before the first evaluation)");
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, lc->ws.trace, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    uint64_t init[16] = {0};
    init[0] = init[8] = ~0ull;
    CU(cudaMemcpy(lc->ws.trace, init, sizeof init, cudaMemcpyHostToDevice));
    return NMB_OK;
}

extern "C" int nmb_last_kernels(nmb_lc* lc, const char** stage1, const char** stage2) {
    if (!lc || !stage1 || !stage2) return fail(NMB_ERR_ARG, "nmb_last_kernels: null pointer");
    *stage1 = lc->last_stage1;
    *stage2 = lc->last_stage2;
    return NMB_OK;
}

extern "C" int nmb_stage_timing(nmb_lc* lc, int enable) {
    if (!lc) return fail(NMB_ERR_ARG, "nmb_stage_timing: null pointer");
    lc->timer.drain(0);
    lc->timer.acc1 = lc->timer.acc2 = 0; lc->timer.n = 0;
    lc->timer.enable(enable != 0);
    return NMB_OK;
}

extern "C" int nmb_stage_times(nmb_lc* lc, double* stage1_ms, double* stage2_ms, int64_t* count) {
    if (!lc || !stage1_ms || !stage2_ms) return fail(NMB_ERR_ARG, "nmb_stage_times: null pointer");
    lc->timer.drain(0);
    const long n = lc->timer.n;
    *stage1_ms = n ? lc->timer.acc1 / n : 0.0;
    *stage2_ms = n ? lc->timer.acc2 / n : 0.0;
    if (count) *count = n;
    lc->timer.acc1 = lc->timer.acc2 = 0; lc->timer.n = 0;
    return NMB_OK;
}

// ---- the pre-fit steps for a batch of candidates: outlier cut, detrending, fit window ---------------------------
// (namaste/namaste.py:1459-1467 AnomCutDiff; namaste/k2flatten.py:51-84 ReduceNoise; namaste/namaste.py:550-556)
// One workspace handle owns the device buffers (grown on demand, never freed between calls) and a stream, so a
// call costs its copies and launches only.  Host arrays are [ncand][stride], candidate c holding n[c] points.
struct nmb_flat {
    int device = 0;
    cudaStream_t st = nullptr;
    double *d_t = nullptr, *d_f = nullptr, *d_e = nullptr, *d_out = nullptr, *d_par = nullptr;
    unsigned char* d_keep = nullptr;
    long long *d_n = nullptr, *d_off = nullptr;
    nmb::FlatStep* d_steps = nullptr;
    int* d_info = nullptr;
    size_t cap_pts = 0, cap_cand = 0, cap_steps = 0;
    std::mutex mu;
};

extern "C" int nmb_flat_create(nmb_flat** out) {
    if (!out) return fail(NMB_ERR_ARG, "nmb_flat_create: null pointer");
    nmb_flat* w = new nmb_flat;
    cudaError_t e = cudaGetDevice(&w->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_info, 2 * sizeof(int));
    if (e != cudaSuccess) {
        delete w;
        return cuda_fail(e, "nmb_flat_create");
    }
    *out = w;
    return NMB_OK;
}

extern "C" void nmb_flat_destroy(nmb_flat* w) {
    if (!w) return;
    if (w->st) cudaStreamSynchronize(w->st);
    cudaFree(w->d_t); cudaFree(w->d_f); cudaFree(w->d_e); cudaFree(w->d_out); cudaFree(w->d_par); cudaFree(w->d_keep);
    cudaFree(w->d_n); cudaFree(w->d_off); cudaFree(w->d_steps); cudaFree(w->d_info);
    if (w->st) cudaStreamDestroy(w->st);
    delete w;
}

namespace {

int flat_reserve(nmb_flat* w, size_t pts, size_t ncand, size_t nsteps) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != w->device) return fail(NMB_ERR_ARG, "nmb_flat: the workspace belongs to another CUDA device");
    if (pts > w->cap_pts) {
        cudaFree(w->d_t); cudaFree(w->d_f); cudaFree(w->d_e); cudaFree(w->d_out); cudaFree(w->d_keep);
        w->d_t = w->d_f = w->d_e = w->d_out = nullptr; w->d_keep = nullptr; w->cap_pts = 0;
        CU(cudaMalloc(&w->d_t, pts * sizeof(double)));
        CU(cudaMalloc(&w->d_f, pts * sizeof(double)));
        CU(cudaMalloc(&w->d_e, pts * sizeof(double)));
        CU(cudaMalloc(&w->d_out, pts * sizeof(double)));
        CU(cudaMalloc(&w->d_keep, pts));
        w->cap_pts = pts;
    }
    if (ncand > w->cap_cand) {
        cudaFree(w->d_n); cudaFree(w->d_off); cudaFree(w->d_par);
        w->d_n = w->d_off = nullptr; w->d_par = nullptr; w->cap_cand = 0;
        CU(cudaMalloc(&w->d_n, ncand * sizeof(long long)));
        CU(cudaMalloc(&w->d_off, (ncand + 1) * sizeof(long long)));
        CU(cudaMalloc(&w->d_par, 2 * ncand * sizeof(double)));
        w->cap_cand = ncand;
    }
    if (nsteps > w->cap_steps) {
        cudaFree(w->d_steps);
        w->d_steps = nullptr; w->cap_steps = 0;
        CU(cudaMalloc(&w->d_steps, nsteps * sizeof(nmb::FlatStep)));
        w->cap_steps = nsteps;
    }
    return NMB_OK;
}

int flat_check_sizes(const int64_t* n, int64_t ncand, int64_t stride, int64_t nmin, const char* who) {
    if (ncand < 1 || stride < 1) return fail(NMB_ERR_ARG, std::string(who) + ": empty batch");
    if ((double)ncand * (double)stride > 2.0e9) return fail(NMB_ERR_ARG, std::string(who) + ": batch exceeds 2^31 points");
    for (int64_t c = 0; c < ncand; ++c)
        if (n[c] < nmin || n[c] > stride) return fail(NMB_ERR_ARG, std::string(who) + ": a light curve is too short or longer than the stride");
    return NMB_OK;
}

}  // namespace

// AnomCutDiff(flux, thresh) for ncand NaN-free flux series; keep [ncand][stride] out (1 = keep)
extern "C" int nmb_anomcut_batch(nmb_flat* w, const double* flux, const int64_t* n, int64_t ncand, int64_t stride,
                                 double thresh, uint8_t* keep) {
    if (!w || !flux || !n || !keep) return fail(NMB_ERR_ARG, "nmb_anomcut_batch: null pointer");
    if (int rc = flat_check_sizes(n, ncand, stride, 3, "nmb_anomcut_batch")) return rc;
    std::lock_guard<std::mutex> lock(w->mu);
    const size_t pts = (size_t)ncand * stride;
    if (int rc = flat_reserve(w, pts, ncand, 0)) return rc;
    std::vector<long long> hn(n, n + ncand);
    CU(cudaMemcpyAsync(w->d_n, hn.data(), ncand * sizeof(long long), cudaMemcpyHostToDevice, w->st));
    CU(cudaMemcpyAsync(w->d_f, flux, pts * sizeof(double), cudaMemcpyHostToDevice, w->st));
    nmb::anomcut_kernel<<<(unsigned)ncand, nmb::kCutThreads, 0, w->st>>>(w->d_f, w->d_n, stride, thresh, w->d_keep);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(keep, w->d_keep, pts, cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    return NMB_OK;
}

// keep = abs(t - tcen < width) (the reference's one-sided quirk) or |t - tcen| < width, per candidate
extern "C" int nmb_window_mask_batch(nmb_flat* w, const double* t, const int64_t* n, int64_t ncand, int64_t stride,
                                     const double* tcen, const double* width, int two_sided, uint8_t* keep) {
    if (!w || !t || !n || !tcen || !width || !keep) return fail(NMB_ERR_ARG, "nmb_window_mask_batch: null pointer");
    if (int rc = flat_check_sizes(n, ncand, stride, 1, "nmb_window_mask_batch")) return rc;
    std::lock_guard<std::mutex> lock(w->mu);
    const size_t pts = (size_t)ncand * stride;
    if (int rc = flat_reserve(w, pts, ncand, 0)) return rc;
    std::vector<long long> hn(n, n + ncand);
    CU(cudaMemcpyAsync(w->d_n, hn.data(), ncand * sizeof(long long), cudaMemcpyHostToDevice, w->st));
    CU(cudaMemcpyAsync(w->d_t, t, pts * sizeof(double), cudaMemcpyHostToDevice, w->st));
    CU(cudaMemcpyAsync(w->d_par, tcen, ncand * sizeof(double), cudaMemcpyHostToDevice, w->st));
    CU(cudaMemcpyAsync(w->d_par + ncand, width, ncand * sizeof(double), cudaMemcpyHostToDevice, w->st));
    nmb::window_mask_kernel<<<(unsigned)((pts + 255) / 256), 256, 0, w->st>>>(w->d_t, w->d_n, stride, (int)ncand, w->d_par,
                                                                              w->d_par + ncand, two_sided, w->d_keep);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(keep, w->d_keep, pts, cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    return NMB_OK;
}

// ReduceNoise for ncand light curves: t (already shifted to start at 0), f and e (already divided by the median
// flux) host [ncand][stride]; out host [ncand][stride], entries no detrending step owns stay 0 like in the reference.
// Windows, boxes and owned ranges are found on the device (flat_windows_kernel), then one thread block per step
// runs the clipped polynomial fit (flatten_kernel).
extern "C" int nmb_flatten_batch(nmb_flat* w, const double* t, const double* f, const double* e, const int64_t* n,
                                 int64_t ncand, int64_t stride, double winsize, double stepsize, double gapthreshold, int deg,
                                 int niter, double sigclip, double* out) {
    if (!w || !t || !f || !e || !n || !out) return fail(NMB_ERR_ARG, "nmb_flatten_batch: null pointer");
    if (int rc = flat_check_sizes(n, ncand, stride, 2, "nmb_flatten_batch")) return rc;
    if (deg < 0 || deg > nmb::kFlatMaxDeg) return fail(NMB_ERR_ARG, "nmb_flatten_batch: polynomial degree must be in 0..5");
    if (niter < 0) return fail(NMB_ERR_ARG, "nmb_flatten_batch: niter < 0");
    if (!(winsize > 0.0) || !(stepsize > 0.0)) return fail(NMB_ERR_ARG, "nmb_flatten_batch: winsize and stepsize must be positive");
    std::lock_guard<std::mutex> lock(w->mu);
    // steps of each candidate: nsteps = ceil(lenlc / stepsize)   (k2flatten.py:65)
    std::vector<long long> off(ncand + 1, 0), hn(n, n + ncand);
    for (int64_t c = 0; c < ncand; ++c) {
        const double lenlc = t[c * stride + n[c] - 1];
        if (!(lenlc > 0.0) || !std::isfinite(lenlc)) return fail(NMB_ERR_ARG, "nmb_flatten_batch: times must start at 0 and increase");
        off[c + 1] = off[c] + (long long)std::ceil(lenlc / stepsize);
    }
    const long long total = off[ncand];
    if (total < 1 || total > 0x7fffffffLL) return fail(NMB_ERR_ARG, "nmb_flatten_batch: number of detrending steps out of range");
    const size_t pts = (size_t)ncand * stride;
    if (int rc = flat_reserve(w, pts, ncand, (size_t)total)) return rc;
    cudaStream_t st = w->st;
    CU(cudaMemcpyAsync(w->d_n, hn.data(), ncand * sizeof(long long), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w->d_off, off.data(), (ncand + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w->d_t, t, pts * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w->d_f, f, pts * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w->d_e, e, pts * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(w->d_out, 0, pts * sizeof(double), st));
    CU(cudaMemsetAsync(w->d_info, 0, 2 * sizeof(int), st));
    nmb::flat_windows_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(w->d_t, w->d_n, stride, (int)ncand, w->d_off, winsize,
                                                                            stepsize, gapthreshold, deg, w->d_steps, w->d_info);
    g_launches++;
    CU(cudaGetLastError());
    int info[2] = {0, 0};
    CU(cudaMemcpyAsync(info, w->d_info, sizeof info, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (info[1]) {
        const long long g = info[1] - 1;
        const int64_t c = std::upper_bound(off.begin(), off.end(), g) - off.begin() - 1;
        char buf[200];
        snprintf(buf, sizeof buf, "light curve %lld, step %lld: the fitting window holds fewer than the %d points a degree-%d "
                 "polynomial needs", (long long)c, (long long)(g - off[c]), deg + 1, deg);
        return fail(NMB_ERR_PRECOND, buf);
    }
    const int maxwin = std::max(1, info[0]);
    const size_t smem = (size_t)maxwin * (5 * sizeof(double) + 1) + 16;
    if (smem > 200 * 1024) return fail(NMB_ERR_ARG, "nmb_flatten_batch: fitting window too large for shared memory");
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(nmb::flatten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nmb::flatten_kernel<<<(unsigned)total, nmb::kFlatThreads, smem, st>>>(w->d_t, w->d_f, w->d_e, w->d_steps, deg, niter, sigclip,
                                                                         maxwin, w->d_out);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, w->d_out, pts * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return NMB_OK;
}

// Host utility of the Python layer's staging cache: a position-sensitive checksum of a host array of 64-bit
// words (sum and sum of prefix sums modulo 2^64, Fletcher style, four interleaved lanes), one pass at memory
// speed.  It lets the literal reference callable MonoOnlyLogProb(params, lc_array, ...) recognise the light
// curve it staged before, including in-place edits of a few flux values.
extern "C" int nmb_host_checksum(const void* data, int64_t nwords, uint64_t* out2) {
    if ((!data && nwords > 0) || !out2 || nwords < 0) return fail(NMB_ERR_ARG, "nmb_host_checksum: bad argument");
    const uint64_t* w = static_cast<const uint64_t*>(data);
    uint64_t s[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 4 <= nwords; i += 4) {
        s[0] += w[i]; t[0] += s[0];
        s[1] += w[i + 1]; t[1] += s[1];
        s[2] += w[i + 2]; t[2] += s[2];
        s[3] += w[i + 3]; t[3] += s[3];
    }
    for (int k = 0; i < nwords; ++i, ++k) { s[k] += w[i]; t[k] += s[k]; }
    out2[0] = s[0] + 3 * s[1] + 5 * s[2] + 7 * s[3];
    out2[1] = t[0] ^ (t[1] * 0x9E3779B97F4A7C15ull) ^ (t[2] * 0xC2B2AE3D27D4EB4Full) ^ (t[3] * 0x165667B19E3779F9ull);
    return NMB_OK;
}

extern "C" int nmb_measure_fp64_peak(double* tflops, double* ms_out) {
    if (!tflops) return fail(NMB_ERR_ARG, "nmb_measure_fp64_peak: null pointer");
    double* d = nullptr;
    CU(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    for (int rep = 0; rep < 2; ++rep) nmb::dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        nmb::dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    g_launches += 7;
    const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return NMB_OK;
}

// =============================================================================================
// Device-resident ensemble sampler (stretch move); see sampler.cuh
// =============================================================================================
struct nmb_sampler {
    nmb_lc* lc = nullptr;
    int64_t W = 0, C = 0;
    int mode = NMB_MODE_WINDOWED;
    int ndim = 6;
    bool gp = false;  // objective: MonoLogProb (GP noise model) instead of MonoOnlyLogProb
    int gp_kind = 0, gp_freemask = 0;
    double gp_kfixed[nmb::kGpMaxK] = {0, 0, 0, 0, 0};
    double a = 2.0;
    unsigned long long seed = 0;
    double *d_pos = nullptr, *d_lnp = nullptr, *d_q = nullptr, *d_zz = nullptr, *d_logu = nullptr;
    double *d_chain = nullptr, *d_lnpchain = nullptr, *d_priors = nullptr;
    long long* d_nacc = nullptr;
    int* d_nan = nullptr;
    int64_t tcap = 0, tlen = 0;
    long long step = 0;
    bool has_state = false;
    cudaStream_t st = nullptr;
    // walker split over several GPUs: pos, lnp and the arrival flags live in ONE allocation (d_xchg) whose
    // IPC handle the peers open; peer_base[r] is rank r's block mapped into this process (null for self)
    double* d_xchg = nullptr;
    int* d_flags = nullptr;    // [kMaxPeers] arrival epochs written by the peers, then [0] ticket, [1] error
    int split_rank = 0, split_world = 1;
    void* peer_base[nmb::kMaxPeers] = {};
    long long split_epoch = 0;
    int cand0 = 0;             // index of the batch's first candidate in the random stream (nmb_sampler_set_candidate_offset)
};

namespace {

__global__ void append_chain_kernel(nmb::SampleCtx sc, long long n) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n) return;
    double* row = sc.chain + (g * sc.tcap + sc.tcol) * sc.ndim;
    for (int d = 0; d < sc.ndim; ++d) row[d] = sc.pos[g * sc.ndim + d];
    sc.lnpchain[g * sc.tcap + sc.tcol] = sc.lnp[g];
}

// bytes of the exchange block: pos [rows][ndim], lnp [rows], then kMaxPeers + 2 ints (flags, ticket, error)
size_t split_block_bytes(size_t rows, int ndim) {
    return (rows * ndim + rows) * sizeof(double) + (nmb::kMaxPeers + 2) * sizeof(int) + 8;
}

int sampler_reserve(nmb_sampler* s, int64_t need) {
    if (need <= s->tcap) return NMB_OK;
    int64_t cap = std::max<int64_t>(need, std::max<int64_t>(16, s->tcap * 2));
    const size_t rows = (size_t)s->C * s->W;
    double *nc = nullptr, *nl = nullptr;
    CU(cudaMalloc(&nc, rows * cap * s->ndim * sizeof(double)));
    CU(cudaMalloc(&nl, rows * cap * sizeof(double)));
    if (s->tlen > 0) {
        CU(cudaMemcpy2DAsync(nc, cap * s->ndim * 8, s->d_chain, s->tcap * s->ndim * 8, s->tlen * s->ndim * 8, rows, cudaMemcpyDeviceToDevice, s->st));
        CU(cudaMemcpy2DAsync(nl, cap * 8, s->d_lnpchain, s->tcap * 8, s->tlen * 8, rows, cudaMemcpyDeviceToDevice, s->st));
        CU(cudaStreamSynchronize(s->st));
    }
    cudaFree(s->d_chain);
    cudaFree(s->d_lnpchain);
    s->d_chain = nc;
    s->d_lnpchain = nl;
    s->tcap = cap;
    return NMB_OK;
}

nmb::SampleCtx sampler_ctx(nmb_sampler* s, int half, int k0, int store) {
    nmb::SampleCtx sc{};
    sc.enabled = 1; sc.W = (int)s->W; sc.half = half; sc.k0 = k0; sc.c0 = s->cand0; sc.store = store; sc.ndim = s->ndim;
    sc.step = s->step; sc.tcol = s->tlen; sc.tcap = s->tcap; sc.seed = s->seed; sc.a = s->a;
    sc.pos = s->d_pos; sc.lnp = s->d_lnp; sc.q = s->d_q; sc.zz = s->d_zz; sc.logu = s->d_logu;
    sc.chain = s->d_chain; sc.lnpchain = s->d_lnpchain; sc.nacc = s->d_nacc; sc.nan_count = s->d_nan;
    return sc;
}

int sampler_check_nan(nmb_sampler* s, const char* what) {
    int nan = 0;
    CU(cudaMemcpyAsync(&nan, s->d_nan, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    if (nan) {
        CU(cudaMemsetAsync(s->d_nan, 0, sizeof(int), s->st));
        return fail(NMB_ERR_NAN, what);
    }
    return NMB_OK;
}

}  // namespace

extern "C" void nmb_sampler_destroy(nmb_sampler* s) {
    if (!s) return;
    if (s->st) {
        cudaStreamSynchronize(s->st);
        // the light curve remembers the last stream that used its workspace: forget this one before it is destroyed
        std::lock_guard<std::mutex> g(g_live_mu);
        if (g_live_lc.count(s->lc) && s->lc->ws_last_stream == s->st) s->lc->ws_used = false;
    }
    for (void* pb : s->peer_base)
        if (pb) cudaIpcCloseMemHandle(pb);
    cudaFree(s->d_xchg);  // pos, lnp and the split flags
    cudaFree(s->d_q); cudaFree(s->d_zz); cudaFree(s->d_logu);
    cudaFree(s->d_chain); cudaFree(s->d_lnpchain); cudaFree(s->d_priors); cudaFree(s->d_nacc); cudaFree(s->d_nan);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

static int sampler_create_impl(nmb_lc* lc, int64_t W, int ndim, const double* priors, uint64_t seed, double a, int mode,
                               nmb_sampler** out) {
    if (!lc || !out) return fail(NMB_ERR_ARG, "nmb_sampler_create: null pointer");
    if (W < 2 || (W % 2) != 0) return fail(NMB_ERR_ARG, "The number of walkers must be even.");
    if (W < 2 * ndim)
        return fail(NMB_ERR_ARG, "The number of walkers needs to be more than twice the dimension of your parameter space.");
    if (!(a > 1.0)) return fail(NMB_ERR_ARG, "nmb_sampler_create: stretch scale a must be > 1");
    if (mode != NMB_MODE_BRUTE && mode != NMB_MODE_WINDOWED) return fail(NMB_ERR_ARG, "nmb_sampler_create: unknown mode");
    nmb_sampler* s = new nmb_sampler;
    s->lc = lc; s->W = W; s->C = lc->ncand; s->mode = mode; s->a = a; s->seed = seed; s->ndim = ndim;
    const size_t rows = (size_t)s->C * W, hrows = (size_t)s->C * (W / 2);
    cudaError_t e = cudaSuccess;
    const size_t xchg_bytes = split_block_bytes(rows, ndim);
    if ((e = cudaMalloc(&s->d_xchg, xchg_bytes)) != cudaSuccess ||
        (e = cudaMemset(s->d_xchg, 0, xchg_bytes)) != cudaSuccess) {
        nmb_sampler_destroy(s);
        return cuda_fail(e, "nmb_sampler_create");
    }
    s->d_pos = s->d_xchg;
    s->d_lnp = s->d_xchg + rows * ndim;
    s->d_flags = reinterpret_cast<int*>(s->d_xchg + rows * ndim + rows);
    if ((e = cudaMalloc(&s->d_q, hrows * ndim * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&s->d_zz, hrows * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&s->d_logu, hrows * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&s->d_nacc, rows * sizeof(long long))) != cudaSuccess ||
        (e = cudaMalloc(&s->d_nan, sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(s->d_nacc, 0, rows * sizeof(long long))) != cudaSuccess ||
        (e = cudaMemset(s->d_nan, 0, sizeof(int))) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking)) != cudaSuccess) {
        nmb_sampler_destroy(s);
        return cuda_fail(e, "nmb_sampler_create");
    }
    if (priors) {
        const size_t np = (size_t)s->C * ndim * 6;
        if ((e = cudaMalloc(&s->d_priors, np * sizeof(double))) != cudaSuccess ||
            (e = cudaMemcpy(s->d_priors, priors, np * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
            nmb_sampler_destroy(s);
            return cuda_fail(e, "nmb_sampler_create: priors");
        }
    }
    *out = s;
    return NMB_OK;
}

extern "C" int nmb_sampler_create(nmb_lc* lc, int64_t W, const double* priors, uint64_t seed, double a, int mode,
                                  nmb_sampler** out) {
    return sampler_create_impl(lc, W, NMB_NPAR, priors, seed, a, mode, out);
}

// Sampler over the GP objective MonoLogProb: walker vectors are [free kernel parameters, 6 transit
// parameters] (see nmb_gp_lnprob); priors [ncand][ndim][6].
extern "C" int nmb_sampler_create_gp(nmb_lc* lc, int64_t W, int kind, const double* kfixed, int freemask,
                                     const double* priors, uint64_t seed, double a, nmb_sampler** out) {
    if (!kfixed) return fail(NMB_ERR_ARG, "nmb_sampler_create_gp: null pointer");
    if (kind != NMB_GP_REAL && kind != NMB_GP_QUASI) return fail(NMB_ERR_ARG, "nmb_sampler_create_gp: unknown kernel kind");
    const int nk = kind == NMB_GP_REAL ? 3 : 5;
    if (freemask < 0 || freemask >= (1 << nk)) return fail(NMB_ERR_ARG, "nmb_sampler_create_gp: bad free mask");
    int nfree = 0;
    for (int i = 0; i < nk; ++i) nfree += (freemask >> i) & 1;
    if (int rc = sampler_create_impl(lc, W, nfree + 6, priors, seed, a, NMB_MODE_WINDOWED, out)) return rc;
    nmb_sampler* s = *out;
    s->gp = true; s->gp_kind = kind; s->gp_freemask = freemask;
    for (int i = 0; i < nk; ++i) s->gp_kfixed[i] = kfixed[i];
    return NMB_OK;
}

// lnprob of `W` rows (plain evaluation) or one half-step of the sampler (sc.enabled) with its objective
static int sampler_eval(nmb_sampler* s, const double* params, int64_t W, double* lnprob, const nmb::SampleCtx& sc) {
    if (!s->gp) return lnprob_impl(s->lc, params, W, s->d_priors, lnprob, nullptr, s->mode, NMB_F64, s->st, sc);
    if (sc.enabled) {
        const int n = (int)(W * s->C);
        CU(launch_pdl(nmb::gp_propose_kernel, dim3((unsigned)((n + 63) / 64)), dim3(64), 0, s->st, sc, (int)W, (int)s->C));
        g_launches++;
    }
    return gp_lnprob_impl(s->lc, params, W, s->gp_kind, s->gp_kfixed, s->gp_freemask, s->d_priors, lnprob, nullptr, s->st, sc);
}

extern "C" int nmb_sampler_set_state(nmb_sampler* s, const double* pos, const double* lnp) {
    if (!s || !pos) return fail(NMB_ERR_ARG, "nmb_sampler_set_state: null pointer");
    const size_t rows = (size_t)s->C * s->W;
    for (size_t i = 0; i < rows * s->ndim; ++i)
        if (!std::isfinite(pos[i])) return fail(NMB_ERR_ARG, std::isnan(pos[i]) ? "At least one parameter value was NaN."
                                                                                : "At least one parameter value was infinite.");
    CU(cudaMemcpyAsync(s->d_pos, pos, rows * s->ndim * sizeof(double), cudaMemcpyHostToDevice, s->st));
    if (lnp) {
        CU(cudaMemcpyAsync(s->d_lnp, lnp, rows * sizeof(double), cudaMemcpyHostToDevice, s->st));
    } else {
        nmb::SampleCtx off{};
        if (int rc = sampler_eval(s, s->d_pos, s->W, s->d_lnp, off)) return rc;
    }
    std::vector<double> h(rows);
    CU(cudaMemcpyAsync(h.data(), s->d_lnp, rows * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (double v : h)
        if (std::isnan(v)) return fail(NMB_ERR_NAN, "The initial lnprob was NaN.");
    s->has_state = true;
    return NMB_OK;
}

// One half-step for walkers [k0, k0+nk) of half `half` (the whole half when nk == W/2).
extern "C" int nmb_sampler_halfstep(nmb_sampler* s, int half, int64_t k0, int64_t nk, int store) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_halfstep: null pointer");
    if (!s->has_state) return fail(NMB_ERR_ARG, "nmb_sampler_halfstep: no state (call set_state first)");
    if (half < 0 || half > 1 || k0 < 0 || nk < 1 || k0 + nk > s->W / 2) return fail(NMB_ERR_ARG, "nmb_sampler_halfstep: bad slice");
    if (store)
        if (int rc = sampler_reserve(s, s->tlen + 1)) return rc;
    const nmb::SampleCtx sc = sampler_ctx(s, half, (int)k0, store);
    return sampler_eval(s, s->d_q, nk, nullptr, sc);
}

// Close a step driven through nmb_sampler_halfstep: advance the RNG step, count a stored column.
extern "C" int nmb_sampler_advance(nmb_sampler* s, int stored, int append_all) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_advance: null pointer");
    if (append_all) {  // (multi-GPU walker split) append the gathered ensemble as one chain column
        if (int rc = sampler_reserve(s, s->tlen + 1)) return rc;
        const nmb::SampleCtx sc = sampler_ctx(s, 0, 0, 1);
        const long long n = (long long)s->C * s->W;
        append_chain_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s->st>>>(sc, n);
        g_launches++;
        CU(cudaGetLastError());
        stored = 1;
    }
    s->step += 1;
    if (stored) s->tlen += 1;
    return NMB_OK;
}

extern "C" int nmb_sampler_run(nmb_sampler* s, int64_t nsteps, int64_t thin, int store) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_run: null pointer");
    if (!s->has_state) return fail(NMB_ERR_ARG, "Cannot have pos0=None if run_mcmc has never been called.");
    if (nsteps < 0 || thin < 1) return fail(NMB_ERR_ARG, "nmb_sampler_run: bad nsteps/thin");
    if (store)
        if (int rc = sampler_reserve(s, s->tlen + nsteps / thin + 1)) return rc;
    const int64_t halfk = s->W / 2;
    for (int64_t i = 0; i < nsteps; ++i) {
        const int st_col = (store && (i % thin == 0)) ? 1 : 0;
        for (int half = 0; half < 2; ++half) {
            const nmb::SampleCtx sc = sampler_ctx(s, half, 0, st_col);
            if (int rc = sampler_eval(s, s->d_q, halfk, nullptr, sc)) return rc;
        }
        s->step += 1;
        if (st_col) s->tlen += 1;
    }
    return sampler_check_nan(s, "lnprob returned NaN.");
}


// ---- one ensemble split by walker over several GPUs (BASELINE config 5) ---------------------------------
// see split_kernels.cuh for the protocol
extern "C" int nmb_sampler_split_export(nmb_sampler* s, void* handle_out) {
    if (!s || !handle_out) return fail(NMB_ERR_ARG, "nmb_sampler_split_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == NMB_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->d_xchg));
    std::memcpy(handle_out, &h, sizeof h);
    return NMB_OK;
}

extern "C" int nmb_sampler_split_attach(nmb_sampler* s, int rank, int world, const void* handles) {
    if (!s || !handles) return fail(NMB_ERR_ARG, "nmb_sampler_split_attach: null pointer");
    if (s->C != 1) return fail(NMB_ERR_ARG, "nmb_sampler_split_attach: the walker split works on a single light curve");
    if (world < 1 || world > nmb::kMaxPeers || rank < 0 || rank >= world)
        return fail(NMB_ERR_ARG, "nmb_sampler_split_attach: rank / world out of range (at most 16 ranks)");
    if ((s->W / 2) % world) return fail(NMB_ERR_ARG, "nmb_sampler_split_attach: half the ensemble must divide evenly over the ranks");
    for (int r = 0; r < world; ++r) {
        if (r == rank || s->peer_base[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + (size_t)r * NMB_IPC_HANDLE_BYTES, sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle(&s->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            s->peer_base[r] = nullptr;
            return cuda_fail(e, "nmb_sampler_split_attach: cudaIpcOpenMemHandle (peer access between the GPUs is required)");
        }
    }
    s->split_rank = rank;
    s->split_world = world;
    return NMB_OK;
}

static nmb::SplitPeers split_peers(const nmb_sampler* s) {
    nmb::SplitPeers pp{};
    pp.world = s->split_world;
    pp.rank = s->split_rank;
    const size_t rows = (size_t)s->C * s->W;
    for (int r = 0; r < s->split_world; ++r) {
        double* base = static_cast<double*>(s->peer_base[r]);
        if (!base) continue;
        pp.pos[r] = base;
        pp.lnp[r] = base + rows * s->ndim;
        pp.flags[r] = reinterpret_cast<int*>(base + rows * s->ndim + rows);
    }
    return pp;
}

// nsteps steps of the split sampler: per half-step this rank's slice [k0, k0 + nk) of the active half is
// updated (two launches), pushed into every peer's replica and the arrival flags are exchanged (one launch).
// No host round trip and no NCCL call inside the loop; synchronises once at the end.
extern "C" int nmb_sampler_run_split(nmb_sampler* s, int64_t nsteps, int store_all) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_run_split: null pointer");
    if (!s->has_state) return fail(NMB_ERR_ARG, "Cannot have pos0=None if run_mcmc has never been called.");
    if (nsteps < 0) return fail(NMB_ERR_ARG, "nmb_sampler_run_split: nsteps < 0");
    const int world = s->split_world, rank = s->split_rank;
    for (int r = 0; r < world; ++r)
        if (r != rank && !s->peer_base[r]) return fail(NMB_ERR_ARG, "nmb_sampler_run_split: peers are not attached");
    const int64_t halfk = s->W / 2, nk = halfk / world, k0 = rank * nk;
    if (store_all)
        if (int rc = sampler_reserve(s, s->tlen + nsteps)) return rc;
    const nmb::SplitPeers pp = split_peers(s);
    static const double tmo_s = std::getenv("NMB_SPLIT_TIMEOUT_S") ? std::atof(std::getenv("NMB_SPLIT_TIMEOUT_S")) : 10.0;
    const unsigned long long timeout_ns = (unsigned long long)(tmo_s * 1e9);
    const int xblocks = (int)std::max<int64_t>(1, std::min<int64_t>(32, (nk * s->ndim + 1023) / 1024));
    if (world > 1) {
        // Handshake before the first half-step: an exchange without rows.  Every rank's replica is in its initial
        // state (set_state's copy and its evaluation of the initial lnprob precede this kernel in the stream)
        // before any peer stores a row into it; without it a rank that is a half-step ahead could have its rows
        // overwritten by a slower rank's own initialisation.
        const int epoch = (int)(++s->split_epoch);
        CU(launch_pdl(nmb::split_exchange_kernel, dim3(1), dim3(256), 0, s->st, pp, (const double*)s->d_pos,
                      (const double*)s->d_lnp, s->ndim, 0LL, 0, epoch, s->d_flags, timeout_ns));
        g_launches++;
    }
    for (int64_t i = 0; i < nsteps; ++i) {
        for (int half = 0; half < 2; ++half) {
            const nmb::SampleCtx sc = sampler_ctx(s, half, (int)k0, 0);
            if (int rc = sampler_eval(s, s->d_q, nk, nullptr, sc)) return rc;
            if (world > 1) {
                const int epoch = (int)(++s->split_epoch);
                CU(launch_pdl(nmb::split_exchange_kernel, dim3(xblocks), dim3(256), 0, s->st, pp, (const double*)s->d_pos,
                              (const double*)s->d_lnp, s->ndim, (long long)(half * halfk + k0), (int)nk, epoch, s->d_flags,
                              timeout_ns));
                g_launches++;
            }
        }
        if (store_all) {
            const nmb::SampleCtx sc = sampler_ctx(s, 0, 0, 1);
            const long long n = (long long)s->C * s->W;
            append_chain_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s->st>>>(sc, n);
            g_launches++;
            CU(cudaGetLastError());
            s->tlen += 1;
        }
        s->step += 1;
    }
    int err = 0;
    CU(cudaMemcpyAsync(&err, s->d_flags + nmb::kMaxPeers + 1, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    if (err) {
        char buf[128];
        snprintf(buf, sizeof buf, "nmb_sampler_run_split: rank %d did not arrive within the time-out", err - 1);
        return fail(NMB_ERR_CUDA, buf);
    }
    return sampler_check_nan(s, "lnprob returned NaN.");
}

// NCCL transport of the same exchange: this rank's slice of a half as [nk][ndim + 1] rows (device), and the
// gathered [world][nk][ndim + 1] rows back into the replica.  Both run in the sampler's stream.
extern "C" int nmb_sampler_pack_half(nmb_sampler* s, int half, int64_t k0, int64_t nk, double* out) {
    if (!s || !out) return fail(NMB_ERR_ARG, "nmb_sampler_pack_half: null pointer");
    if (s->C != 1) return fail(NMB_ERR_ARG, "nmb_sampler_pack_half: single light curve only");
    if (half < 0 || half > 1 || k0 < 0 || nk < 1 || k0 + nk > s->W / 2) return fail(NMB_ERR_ARG, "nmb_sampler_pack_half: bad slice");
    const long long n = (long long)nk * (s->ndim + 1);
    nmb::split_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->st>>>(s->d_pos, s->d_lnp, s->ndim,
                                                                         (long long)(half * (s->W / 2) + k0), (int)nk, out);
    g_launches++;
    CU(cudaGetLastError());
    return NMB_OK;
}
extern "C" int nmb_sampler_unpack_half(nmb_sampler* s, int half, const double* in) {
    if (!s || !in) return fail(NMB_ERR_ARG, "nmb_sampler_unpack_half: null pointer");
    if (s->C != 1) return fail(NMB_ERR_ARG, "nmb_sampler_unpack_half: single light curve only");
    if (half < 0 || half > 1) return fail(NMB_ERR_ARG, "nmb_sampler_unpack_half: bad half");
    const long long halfk = s->W / 2, n = halfk * (s->ndim + 1);
    nmb::split_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->st>>>(s->d_pos, s->d_lnp, s->ndim, half * halfk, halfk, in);
    g_launches++;
    CU(cudaGetLastError());
    return NMB_OK;
}

extern "C" int nmb_sampler_sync(nmb_sampler* s) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_sync: null pointer");
    return sampler_check_nan(s, "lnprob returned NaN.");
}

// The current CUDA device is a property of the host THREAD: a thread that drives a handle staged by another thread
// selects that thread's device first (CandidateBatchSampler's group threads).
extern "C" int nmb_get_device(int* device) {
    if (!device) return fail(NMB_ERR_ARG, "nmb_get_device: null pointer");
    CU(cudaGetDevice(device));
    return NMB_OK;
}
extern "C" int nmb_set_device(int device) {
    CU(cudaSetDevice(device));
    return NMB_OK;
}

extern "C" int nmb_sampler_set_candidate_offset(nmb_sampler* s, int64_t cand0) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_set_candidate_offset: null pointer");
    if (cand0 < 0 || cand0 + s->C > (1LL << 30)) return fail(NMB_ERR_ARG, "nmb_sampler_set_candidate_offset: offset out of range");
    s->cand0 = (int)cand0;
    return NMB_OK;
}

extern "C" int64_t nmb_sampler_len(const nmb_sampler* s) { return s ? s->tlen : -1; }
extern "C" int64_t nmb_sampler_steps(const nmb_sampler* s) { return s ? s->step : -1; }

extern "C" int nmb_sampler_reset(nmb_sampler* s) {
    if (!s) return fail(NMB_ERR_ARG, "nmb_sampler_reset: null pointer");
    s->tlen = 0;
    CU(cudaMemsetAsync(s->d_nacc, 0, (size_t)s->C * s->W * sizeof(long long), s->st));
    CU(cudaStreamSynchronize(s->st));
    return NMB_OK;
}

// what: 0 pos [C][W][6], 1 lnp [C][W], 2 chain [C][W][len][6], 3 lnprobability [C][W][len], 4 naccepted (int64) [C][W]
extern "C" int nmb_sampler_get(nmb_sampler* s, int what, void* out) {
    if (!s || !out) return fail(NMB_ERR_ARG, "nmb_sampler_get: null pointer");
    const size_t rows = (size_t)s->C * s->W;
    switch (what) {
        case 0: CU(cudaMemcpyAsync(out, s->d_pos, rows * s->ndim * 8, cudaMemcpyDeviceToHost, s->st)); break;
        case 1: CU(cudaMemcpyAsync(out, s->d_lnp, rows * 8, cudaMemcpyDeviceToHost, s->st)); break;
        case 2:
            if (s->tlen) CU(cudaMemcpy2DAsync(out, s->tlen * s->ndim * 8, s->d_chain, s->tcap * s->ndim * 8, s->tlen * s->ndim * 8, rows, cudaMemcpyDeviceToHost, s->st));
            break;
        case 3:
            if (s->tlen) CU(cudaMemcpy2DAsync(out, s->tlen * 8, s->d_lnpchain, s->tcap * 8, s->tlen * 8, rows, cudaMemcpyDeviceToHost, s->st));
            break;
        case 4: CU(cudaMemcpyAsync(out, s->d_nacc, rows * 8, cudaMemcpyDeviceToHost, s->st)); break;
        default: return fail(NMB_ERR_ARG, "nmb_sampler_get: unknown item");
    }
    CU(cudaStreamSynchronize(s->st));
    return NMB_OK;
}

// ---- posterior reduction on the device-resident chain (namaste.py:597-608; percentiles as :643-652) ----
// NumPy's percentile ("linear"): virtual index (n-1) q/100, the two neighbouring order statistics,
// a + (b-a) g for g < 1/2 and b - (b-a)(1-g) otherwise.
static double np_lerp(double a, double b, double g) {
    const double d = b - a;
    return g >= 0.5 ? b - d * (1.0 - g) : a + d * g;
}
static void np_virtual_index(long long n, double q, long long& lo, long long& hi, double& g) {
    const double vi = (double)(n - 1) * (q / 100.0);
    lo = (long long)std::floor(vi);
    hi = std::min<long long>(lo + 1, n - 1);
    g = vi - (double)lo;
    if (lo < 0) { lo = 0; hi = 0; g = 0.0; }
}

extern "C" int nmb_sampler_summary(nmb_sampler* s, int64_t nsteps, const double* q, int nq, double* pct, uint8_t* good,
                                   int64_t* ncut_out, int64_t* nsamples_out) {
    if (!s || !q || !pct || !good) return fail(NMB_ERR_ARG, "nmb_sampler_summary: null pointer");
    if (nq < 1 || nq > 16) return fail(NMB_ERR_ARG, "nmb_sampler_summary: between 1 and 16 percentiles");
    for (int i = 0; i < nq; ++i)
        if (!(q[i] >= 0.0 && q[i] <= 100.0)) return fail(NMB_ERR_ARG, "Percentiles must be in the range [0, 100]");
    const long long ncut = std::min<long long>((long long)(nsteps * 0.25), 3000);
    const long long T = s->tlen, L = T - ncut;
    if (L < 1) return fail(NMB_ERR_ARG, "nmb_sampler_summary: the chain is not longer than the burn-in cut");
    if (L > 0x7fffffffLL) return fail(NMB_ERR_ARG, "nmb_sampler_summary: chain too long");
    const size_t rows = (size_t)s->C * s->W;
    const int ndim = s->ndim, ncol = ndim + 1;
    cudaStream_t st = s->st;
    // 1. per walker: 50th and 95th percentile of its post-burn-in lnprob
    long long lo50, hi50, lo95, hi95;
    double g50, g95;
    np_virtual_index(L, 50.0, lo50, hi50, g50);
    np_virtual_index(L, 95.0, lo95, hi95, g95);
    const int hranks[4] = {(int)lo50, (int)hi50, (int)lo95, (int)hi95};
    int* d_ranks = nullptr;
    double* d_wsel = nullptr;
    unsigned char* d_good = nullptr;
    nmb::ColQuery* d_queries = nullptr;
    unsigned int* d_hist = nullptr;
    auto cleanup = [&]() { cudaFree(d_ranks); cudaFree(d_wsel); cudaFree(d_good); cudaFree(d_queries); cudaFree(d_hist); };
#define CUS(call)                                                                     \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) { cleanup(); return cuda_fail(e__, #call); }          \
    } while (0)
    CUS(cudaMalloc(&d_ranks, sizeof hranks));
    CUS(cudaMalloc(&d_wsel, rows * 4 * sizeof(double)));
    CUS(cudaMalloc(&d_good, rows));
    CUS(cudaMemcpyAsync(d_ranks, hranks, sizeof hranks, cudaMemcpyHostToDevice, st));
    nmb::walker_select_kernel<<<(unsigned)rows, nmb::kSelThreads, 0, st>>>(s->d_lnpchain, s->tcap, ncut, (int)L, 4, d_ranks, d_wsel);
    g_launches++;
    std::vector<double> wsel(rows * 4);
    CUS(cudaMemcpyAsync(wsel.data(), d_wsel, rows * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUS(cudaStreamSynchronize(st));
    // 2. failed walkers: 95th percentile not above the median (over walkers) of the medians
    long long nsamp_total = 0;
    std::vector<long long> ngood(s->C, 0);
    for (int64_t c = 0; c < s->C; ++c) {
        std::vector<double> p50(s->W), p95(s->W);
        for (int64_t w = 0; w < s->W; ++w) {
            const double* r = &wsel[(c * s->W + w) * 4];
            p50[w] = np_lerp(r[0], r[1], g50);
            p95[w] = np_lerp(r[2], r[3], g95);
        }
        std::vector<double> sorted = p50;
        std::sort(sorted.begin(), sorted.end());
        const double med = (s->W % 2) ? sorted[s->W / 2] : (sorted[s->W / 2 - 1] + sorted[s->W / 2]) / 2.0;
        for (int64_t w = 0; w < s->W; ++w) {
            good[c * s->W + w] = p95[w] > med ? 1 : 0;
            ngood[c] += good[c * s->W + w];
        }
        nsamp_total += ngood[c] * L;
    }
    if (ncut_out) *ncut_out = ncut;
    if (nsamples_out) *nsamples_out = nsamp_total;
    // 3. percentiles of every column over the kept samples, candidate by candidate
    const int nqueries = ncol * nq * 2;
    CUS(cudaMalloc(&d_queries, nqueries * sizeof(nmb::ColQuery)));
    CUS(cudaMalloc(&d_hist, (size_t)nqueries * 256 * sizeof(unsigned int)));
    CUS(cudaMemcpyAsync(d_good, good, rows, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)nqueries * 256 * sizeof(unsigned int);
    if (smem > 200 * 1024) { cleanup(); return fail(NMB_ERR_ARG, "nmb_sampler_summary: too many percentiles x columns"); }
    if (smem > 48 * 1024) CUS(cudaFuncSetAttribute(nmb::column_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t c = 0; c < s->C; ++c) {
        const long long n = ngood[c] * L;
        double* pc = pct + c * nq * ncol;
        if (n < 1) {
            for (int i = 0; i < nq * ncol; ++i) pc[i] = NAN;
            continue;
        }
        std::vector<nmb::ColQuery> hq(nqueries);
        std::vector<double> gam(nq);
        for (int i = 0; i < nq; ++i) {
            long long lo, hi;
            np_virtual_index(n, q[i], lo, hi, gam[i]);
            for (int d = 0; d < ncol; ++d) {
                hq[(d * nq + i) * 2 + 0] = {0ull, lo, d, 0};
                hq[(d * nq + i) * 2 + 1] = {0ull, hi, d, 0};
            }
        }
        CUS(cudaMemcpyAsync(d_queries, hq.data(), nqueries * sizeof(nmb::ColQuery), cudaMemcpyHostToDevice, st));
        CUS(cudaMemsetAsync(d_hist, 0, (size_t)nqueries * 256 * sizeof(unsigned int), st));
        const long long total = (long long)s->W * L;
        const unsigned blocks = (unsigned)std::min<long long>(148 * 4, (total + nmb::kSelThreads - 1) / nmb::kSelThreads);
        for (int pass = 0; pass < 8; ++pass) {
            nmb::column_hist_kernel<<<blocks, nmb::kSelThreads, smem, st>>>(
                s->d_chain + (size_t)c * s->W * s->tcap * ndim, s->d_lnpchain + (size_t)c * s->W * s->tcap, d_good + c * s->W,
                s->W, s->tcap, ncut, T, ndim, pass, d_queries, nqueries, d_hist);
            nmb::column_pick_kernel<<<(nqueries + 63) / 64, 64, 0, st>>>(d_queries, nqueries, d_hist);
            g_launches += 2;
        }
        CUS(cudaMemcpyAsync(hq.data(), d_queries, nqueries * sizeof(nmb::ColQuery), cudaMemcpyDeviceToHost, st));
        CUS(cudaStreamSynchronize(st));
        for (int d = 0; d < ncol; ++d)
            for (int i = 0; i < nq; ++i) {
                auto val = [](unsigned long long k) {
                    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
                    double v;
                    std::memcpy(&v, &u, sizeof v);
                    return v;
                };
                pc[i * ncol + d] = np_lerp(val(hq[(d * nq + i) * 2].prefix), val(hq[(d * nq + i) * 2 + 1].prefix), gam[i]);
            }
    }
    CUS(cudaGetLastError());
#undef CUS
    cleanup();
    return NMB_OK;
}

// device pointers for zero-copy interop (torch.distributed all-gather of the walker-split mode):
// what: 0 pos, 1 lnp.  The stream is the sampler's own; use nmb_sampler_stream to order against it.
extern "C" void* nmb_sampler_devptr(nmb_sampler* s, int what) {
    if (!s) return nullptr;
    return what == 0 ? (void*)s->d_pos : what == 1 ? (void*)s->d_lnp : nullptr;
}
extern "C" void* nmb_sampler_stream(nmb_sampler* s) { return s ? (void*)s->st : nullptr; }
