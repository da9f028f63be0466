// Device-side affine-invariant ensemble sampler (Goodman & Weare 2010 stretch move) as emcee 2.x
// drives it for Namaste (call sites namaste/namaste.py:590-594).  emcee itself is a third-party
// dependency that is not vendored in the reference; the move is restated from the published
// algorithm (Foreman-Mackey et al. 2013, Alg. 3):
//   per step, for each half-ensemble S0 (complement S1):
//     zz = ((a-1) U + 1)^2 / a,  j ~ U{0..|S1|-1},  q = S1[j] - zz (S1[j] - s)
//     accept iff (ndim-1) ln zz + lnp(q) - lnp(s) > ln U'
// Random numbers come from Philox4x32-10 keyed by the seed with the counter
// (step, half, candidate, walker), so a chain does not depend on how walkers are spread over
// thread blocks or GPUs.  The proposal runs at the top of stage 1 of the likelihood pipeline and
// the accept/reject + chain append run in its finalisation, so a half-step is two launches and
// no host round trip.
#pragma once
#include <stdint.h>

namespace nmb {

struct SampleCtx {
    int enabled;          // 0: plain lnprob evaluation
    int W;                // walkers per candidate (full ensemble)
    int half;             // 0: update first half, 1: update second half
    int k0;               // first walker (index within the half) evaluated by this launch
    int c0;               // index of this batch's first candidate in the random stream (candidate groups of one
                          // catalogue on separate samplers draw the numbers they would draw in one batch)
    int store;            // append to the chain at column tcol
    int ndim;
    long long step;       // global step counter (RNG)
    long long tcol, tcap; // chain column / capacity
    unsigned long long seed;
    double a;
    double* pos;          // [C][W][6]
    double* lnp;          // [C][W]
    double* q;            // [C][Weval][6] proposals (the pipeline's `params`)
    double* zz;           // [C][Weval]
    double* logu;         // [C][Weval]
    double* chain;        // [C][W][tcap][6]
    double* lnpchain;     // [C][W][tcap]
    long long* nacc;      // [C][W]
    int* nan_count;       // lnprob evaluations that returned NaN
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
    const unsigned long long x = (((unsigned long long)hi << 32) | lo) >> 11;
    return (double)x * 1.1102230246251565e-16;  // 2^-53, in [0, 1)
}

// Proposal for evaluated walker `k` (index within this launch) of candidate `cand`: a pure function
// of (seed, step, half, candidate, walker) and the current positions, so several thread blocks may
// recompute it; exactly one of them stores it (propose_store).
constexpr int kMaxDim = 12;  // 6 transit parameters + up to 5 kernel parameters of the GP objective
struct Proposal {
    double q[kMaxDim];
    double zz, u_a;
};
// The part of a proposal that does not depend on the ensemble's state: the stretch factor, the row of the
// complementary walker and the acceptance draw.  A kernel may compute it before its grid dependency resolves.
struct Draw {
    double zz, u_a;
    int i;    // walker being updated (index within the candidate's ensemble)
    int jc;   // complementary walker (index within the candidate's ensemble)
};
__device__ __forceinline__ Draw propose_draw(const SampleCtx& sc, int cand, int k) {
    const int halfk = sc.W / 2;
    const int nS0 = sc.half == 0 ? halfk : sc.W - halfk;
    const int nS1 = sc.W - nS0;
    const int i = (sc.half == 0 ? 0 : halfk) + sc.k0 + k;  // walker being updated
    const int c_base = sc.half == 0 ? halfk : 0;           // complementary half
    uint32_t r[4], r2[4];
    const uint32_t walker = (uint32_t)i, cnd = (uint32_t)(cand + sc.c0);
    philox4x32_10((uint32_t)sc.step, (uint32_t)(sc.step >> 32), (walker << 1) | (uint32_t)sc.half, cnd << 1,
                  (uint32_t)sc.seed, (uint32_t)(sc.seed >> 32), r);
    philox4x32_10((uint32_t)sc.step, (uint32_t)(sc.step >> 32), (walker << 1) | (uint32_t)sc.half, (cnd << 1) | 1u,
                  (uint32_t)sc.seed, (uint32_t)(sc.seed >> 32), r2);
    const double u_z = u01_53(r[0], r[1]), u_j = u01_53(r[2], r[3]);
    Draw dr;
    dr.u_a = u01_53(r2[0], r2[1]);
    const double t = (sc.a - 1.) * u_z + 1;
    dr.zz = t * t / sc.a;
    int j = (int)(u_j * (double)nS1);
    if (j >= nS1) j = nS1 - 1;
    dr.i = i;
    dr.jc = c_base + j;
    return dr;
}
// `s` = the walker's current position (rows of the half being updated do not change during the previous
// half-step, so a caller may have loaded it early); the complementary row is read here.
__device__ __forceinline__ Proposal propose_from(const SampleCtx& sc, int cand, const Draw& dr, const double* s) {
    Proposal pr;
    pr.u_a = dr.u_a;
    pr.zz = dr.zz;
    const double* c = sc.pos + ((long long)cand * sc.W + dr.jc) * sc.ndim;
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d)
        if (d < sc.ndim) pr.q[d] = c[d] - pr.zz * (c[d] - s[d]);
    return pr;
}
__device__ __forceinline__ Proposal propose_compute(const SampleCtx& sc, int cand, int k) {
    const Draw dr = propose_draw(sc, cand, k);
    return propose_from(sc, cand, dr, sc.pos + ((long long)cand * sc.W + dr.i) * sc.ndim);
}
__device__ __forceinline__ void propose_store(const SampleCtx& sc, const Proposal& pr, int cand, int k, int Weval) {
    const long long e = (long long)cand * Weval + k;
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d)
        if (d < sc.ndim) sc.q[e * sc.ndim + d] = pr.q[d];
    sc.zz[e] = pr.zz;
    sc.logu[e] = log(pr.u_a);
}
__device__ __forceinline__ void propose_stretch(const SampleCtx& sc, int cand, int k, int Weval) {
    propose_store(sc, propose_compute(sc, cand, k), cand, k, Weval);
}

// Accept/reject + chain append for evaluated walker `k`; newlnp = loglike + lnprior of its proposal.
__device__ __forceinline__ void accept_stretch(const SampleCtx& sc, int cand, int k, int Weval, double newlnp) {
    const int halfk = sc.W / 2;
    const int i = (sc.half == 0 ? 0 : halfk) + sc.k0 + k;
    const long long e = (long long)cand * Weval + k;
    const long long g = (long long)cand * sc.W + i;
    if (newlnp != newlnp) atomicAdd(sc.nan_count, 1);
    const double lnpdiff = ((sc.ndim - 1.) * log(sc.zz[e]) + newlnp) - sc.lnp[g];
    const bool acc = lnpdiff > sc.logu[e];
    if (acc) {
        for (int d = 0; d < sc.ndim; ++d) sc.pos[g * sc.ndim + d] = sc.q[e * sc.ndim + d];
        sc.lnp[g] = newlnp;
        sc.nacc[g] += 1;
    }
    if (sc.store) {
        double* row = sc.chain + (g * sc.tcap + sc.tcol) * sc.ndim;
        for (int d = 0; d < sc.ndim; ++d) row[d] = sc.pos[g * sc.ndim + d];
        sc.lnpchain[g * sc.tcap + sc.tcol] = sc.lnp[g];
    }
}

}  // namespace nmb
