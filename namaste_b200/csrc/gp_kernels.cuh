// GP noise-model objective MonoLogProb (namaste/namaste.py:1334-1339): celerite's semiseparable
// Cholesky solve with the transit model as mean, one walker per thread.
//
// celerite is a third-party dependency that the reference does not vendor; the factorisation is the
// published one (Foreman-Mackey et al. 2017, AJ 154, 220, eqs. 44-46 with the pre-conditioned U, V,
// phi of sec. 5.2):
//   S_n = phi_n phi_n^T o (S_{n-1} + D_{n-1} W_{n-1} W_{n-1}^T),   D_n = A_n - U_n S_n U_n^T,
//   W_n = (V_n - U_n S_n) / D_n,   f_n = phi_n o (f_{n-1} + W_{n-1} z_{n-1}),   z_n = r_n - U_n f_n,
//   log L = -1/2 (sum z_n^2 / D_n + sum log D_n + N log 2 pi),
// r = flux - model(t).  Kernels (namaste.py:279-314): "Real" = RealTerm(log_a, log_c) + jitter (J = 1),
// "quasi" = RotationTerm(log_amp, log_timescale, log_period, log_factor; :1355-1379) + jitter
// (one real + one complex term, J = 3).  The recurrence is sequential in n, so the parallelism is over
// walkers; the in-transit part of the mean model comes from the stage-2 kernel, which in this mode
// stores the binned model values instead of reducing residuals.
#pragma once
#include "eval_pipeline.cuh"

namespace nmb {

constexpr int kGpReal = 0, kGpQuasi = 1;
constexpr int kGpMaxK = 5;  // kernel parameters of the larger kernel (quasi: 4 + jitter)

struct GpSpec {
    int kind;            // kGpReal / kGpQuasi
    int nk;              // kernel parameters of this kind (3 / 5)
    int nfree;           // how many of them vary per walker (lead the walker vector, kernel order)
    int ndim;            // nfree + 6
    int freemask;        // bit i set: kernel parameter i is free
    double kfixed[kGpMaxK];
};

// walker vector [ndim] -> the six transit parameters (for the likelihood pipeline)
__global__ void gp_mean_params_kernel(const double* __restrict__ params, long long CW, int ndim, int nfree,
                                      double* __restrict__ mean) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= CW * 6) return;
    const long long w = i / 6;
    const int d = (int)(i - w * 6);
    mean[i] = params[w * ndim + nfree + d];
}

// stretch-move proposals of the GP sampler: one thread per evaluated walker (the transit-only sampler
// proposes inside its stage-1 kernels instead)
__global__ void gp_propose_kernel(SampleCtx sc, int Weval, int ncand) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Weval * ncand) return;
    propose_stretch(sc, i / Weval, i % Weval, Weval);
}

template <int JR, int JC>
__global__ void __launch_bounds__(32)
gp_solve_kernel(LcView lc, const double* __restrict__ var, const double* __restrict__ params, int W, GpSpec gs,
                const double* __restrict__ priors, EvalWs ws, double* __restrict__ lnprob,
                double* __restrict__ loglike, SampleCtx sc) {
    constexpr int J = JR + 2 * JC;
    pdl_wait();
    const long long cw = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cw >= (long long)W * lc.ncand) return;
    const int cand = (int)(cw / W);
    const int n = (int)lc.n[cand];
    const double* __restrict__ gt = lc.t + (long long)cand * lc.npad;
    const double* __restrict__ gf = lc.f + (long long)cand * lc.npad;
    const double* __restrict__ gv = var + (long long)cand * lc.npad;
    const double* __restrict__ gm = ws.modelbuf + cw * lc.npad;
    const double* par = params + cw * gs.ndim;
    const int lo = ws.rlo[cw], hi = ws.rhi[cw];

    double kp[kGpMaxK];
    int nf = 0;
#pragma unroll
    for (int i = 0; i < kGpMaxK; ++i) {
        kp[i] = gs.kfixed[i];
        if (i < gs.nk && ((gs.freemask >> i) & 1)) kp[i] = par[nf++];
    }
    // coefficients (RealTerm / RotationTerm.get_*_coefficients)
    double a[J > 0 ? J : 1], bq[J > 0 ? J : 1], c[J > 0 ? J : 1], dfreq = 0.0, jitter, asum;
    if (JC == 0) {
        a[0] = exp(kp[0]);
        c[0] = exp(kp[1]);
        jitter = exp(2 * kp[2]);
        asum = a[0];
    } else {
        const double f = exp(kp[3]);
        a[0] = exp(kp[0]) * (1.0 + f) / (2.0 + f);
        c[0] = exp(-kp[1]);
        const double ac = exp(kp[0]) / (2.0 + f);
        a[1] = ac; a[2] = ac;       // complex amplitude a_c (b_c = 0)
        bq[1] = 0.0; bq[2] = 0.0;
        c[1] = exp(-kp[1]); c[2] = c[1];
        dfreq = 2 * kPi * exp(-kp[2]);
        jitter = exp(2 * kp[4]);
        asum = a[0] + ac;
    }
    (void)bq;

    double S[J][J], Wv[J], fv[J], U[J], V[J];
#pragma unroll
    for (int i = 0; i < J; ++i) {
        fv[i] = 0.0;
#pragma unroll
        for (int k = 0; k < J; ++k) S[i][k] = 0.0;
    }
    // log det K = sum log D_n is accumulated as a product of mantissas plus a sum of binary exponents
    // (one log at the end instead of one per step); 1/D_n is formed once per step and shared by W_n and
    // the quadratic form.  Both differ from the textbook evaluation by rounding only.
    double D = 0.0, rD = 0.0, z = 0.0, quad = 0.0, tprev = 0.0, mant = 1.0;
    long long esum = 0;
    bool indefinite = false;  // a pivot D_n <= 0 (or NaN): celerite's solver raises LinAlgError there
    for (int i = 0; i < n; ++i) {
        const double ti = __ldg(gt + i);
        const double mean = (i >= lo && i < hi) ? gm[i] : 1.0;
        const double r = __ldg(gf + i) - mean;
        const double A = (__ldg(gv + i) + jitter) + asum;
#pragma unroll
        for (int k = 0; k < JR; ++k) { U[k] = a[k]; V[k] = 1.0; }
        if (JC > 0) {
            double sd, cd;
            sincos(dfreq * ti, &sd, &cd);
            U[JR] = a[JR] * cd;      // a_c cos + b_c sin, b_c = 0
            U[JR + 1] = a[JR] * sd;  // a_c sin - b_c cos
            V[JR] = cd;
            V[JR + 1] = sd;
        }
        if (i == 0) {
            D = A;
            rD = 1.0 / D;
#pragma unroll
            for (int k = 0; k < J; ++k) Wv[k] = V[k] * rD;
            z = r;
        } else {
            const double dt = ti - tprev;
            // the rotation term's real and complex parts share one damping rate, so phi is one number
            const double ph = exp(-c[0] * dt);
#pragma unroll
            for (int j = 0; j < J; ++j) {
#pragma unroll
                for (int k = 0; k < J; ++k) S[j][k] = (ph * ph) * (S[j][k] + D * (Wv[j] * Wv[k]));
                fv[j] = ph * (fv[j] + Wv[j] * z);
            }
            double usu = 0.0, us[J];
#pragma unroll
            for (int k = 0; k < J; ++k) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < J; ++j) s += U[j] * S[j][k];
                us[k] = s;
            }
#pragma unroll
            for (int k = 0; k < J; ++k) usu += us[k] * U[k];
            D = A - usu;
            rD = 1.0 / D;
            double uf = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) {
                Wv[k] = (V[k] - us[k]) * rD;
                uf += U[k] * fv[k];
            }
            z = r - uf;
        }
        quad += (z * z) * rD;
        indefinite |= !(D > 0.0);
        int e2;
        mant *= frexp(D, &e2);
        esum += e2;
        if ((i & 255) == 255) {  // keep the running mantissa product away from underflow
            mant = frexp(mant, &e2);
            esum += e2;
        }
        tprev = ti;
    }
    const double logdet = log(mant) + (double)esum * 0.6931471805599453;
    // not positive definite: -inf, what celerite's log_likelihood(quiet=True) returns (an even number of
    // negative pivots would otherwise give a finite value the sampler could accept)
    const double ll = indefinite ? -INFINITY : -0.5 * ((quad + logdet) + n * 1.8378770664093453);  // log(2 pi)
    const double lp = priors ? ln_prior_dev(par, priors + (long long)cand * gs.ndim * 6, gs.ndim) : 0.0;
    finish_walker(sc, cw, W, ll, lp, lnprob, loglike);
    // leave the range workspace idle for the next evaluation
    ws.rlo[cw] = INT_MAX;
    ws.rhi[cw] = 0;
}

}  // namespace nmb
