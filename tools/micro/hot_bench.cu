// Throughput of the in-transit hot path in isolation and with the kernel's surroundings added one
// by one (geometry from loaded times, shared-memory staging + bin reduction), as a function of the
// resident warps per SM sub-partition.  Build: nvcc -O3 -std=c++17 -fmad=false -gencode
// arch=compute_100a,code=sm_100a -o hot_bench hot_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../namaste_b200/csrc/lnprob_kernels.cuh"
using namespace nmb;

// ---- experiment: two interior samples per lane in lock step (ILP 2), MODE 3 -------------------------------
struct D2 { double a, b; };
__device__ __forceinline__ D2 rcp2(D2 x) { return {RelaxedOps::rcp(x.a), RelaxedOps::rcp(x.b)}; }
__device__ __forceinline__ D2 sqrt2(D2 x) { return {RelaxedOps::sqrt(x.a), RelaxedOps::sqrt(x.b)}; }
__device__ __forceinline__ D2 ellpic_x2(D2 p, D2 m1) {
    D2 kc = sqrt2(m1), m0 = {1., 1.}, c = {1., 1.}, rp = rcp2(p), d = rp, e = kc;
    bool more = true;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const D2 g = {e.a * rp.a, e.b * rp.b};
        more = (__double2hiint(fma(m0.a, kHot.one_minus_agm_tol, -kc.a)) > 0) | (__double2hiint(fma(m0.b, kHot.one_minus_agm_tol, -kc.b)) > 0);
        const D2 f = c;
        c = {fma(d.a, rp.a, c.a), fma(d.b, rp.b, c.b)};
        d = {2. * fma(f.a, g.a, d.a), 2. * fma(f.b, g.b, d.b)};
        p = {g.a + p.a, g.b + p.b};
        m0 = {kc.a + m0.a, kc.b + m0.b};
        if (!more) break;
        const D2 se = sqrt2(e);
        kc = {2. * se.a, 2. * se.b};
        e = {kc.a * m0.a, kc.b * m0.b};
        rp = rcp2(p);
    }
    const D2 r = rcp2({m0.a * (m0.a + p.a), m0.b * (m0.b + p.b)});
    return {(kHot.half_pi * fma(c.a, m0.a, d.a)) * r.a, (kHot.half_pi * fma(c.b, m0.b, d.b)) * r.b};
}
__device__ __forceinline__ double interior_tail(double z, double m1, double pi3, double ra, double oma, const QuadConsts& c) {
    const double p = c.p, p2 = c.p2, z2 = z * z;
    const double lg = log_pos_normal(m1);
    const double ee = fma(-(m1 * fma(m1, fma(m1, fma(m1, kHot.eb[3], kHot.eb[2]), kHot.eb[1]), kHot.eb[0])), lg,
                          fma(m1, fma(m1, fma(m1, fma(m1, kHot.ea[3], kHot.ea[2]), kHot.ea[1]), kHot.ea[0]), 1.));
    const double ek = fma(-fma(m1, fma(m1, fma(m1, fma(m1, kHot.kb[4], kHot.kb[3]), kHot.kb[2]), kHot.kb[1]), kHot.kb[0]), lg,
                          fma(m1, fma(m1, fma(m1, fma(m1, kHot.ka[4], kHot.ka[3]), kHot.ka[2]), kHot.ka[1]), kHot.ka[0]));
    const double q = p2 - z2;
    const double t3 = (3 * q) * (ra * pi3);
    const double w7 = fma(7., p2, z2) - 4.;
    const double ck = fma(q, q, fma(-5., z2, 1. + p2));
    const double lam_d = (kHot.two_over_nine_pi * RelaxedOps::rsqrt(oma)) * (fma(ck, ek, (oma * w7) * ee) - t3);
    const double eta_d = (0.5 * p2) * fma(2., z2, p2);
    const double lam_e = 1. - (1. - p2);
    const double num = fma(c.omc2, lam_e, fma(c.c2, lam_d + ((p > z) ? kHot.two_thirds : 0.0), -c.c4 * eta_d));
    return fma(-num, c.rfo, 1.);
}
__device__ __forceinline__ D2 occultquad_interior_x2(D2 z, const QuadConsts& c) {
    const double p = c.p;
    const D2 dm = {z.a - p, z.b - p}, dp = {z.a + p, z.b + p};
    const D2 rdm = rcp2(dm);
    const D2 a = {dm.a * dm.a, dm.b * dm.b}, b = {dp.a * dp.a, dp.b * dp.b};
    const D2 ra = {rdm.a * rdm.a, rdm.b * rdm.b};
    const D2 oma = {1. - a.a, 1. - a.b};
    const D2 ro = rcp2(oma);
    const D2 m1 = {1. - (b.a - a.a) * ro.a, 1. - (b.b - a.b) * ro.b};
    const D2 pi3 = ellpic_x2({dp.a * fabs(rdm.a), dp.b * fabs(rdm.b)}, m1);
    return {interior_tail(z.a, m1.a, pi3.a, ra.a, oma.a, c), interior_tail(z.b, m1.b, pi3.b, ra.b, oma.b, c)};
}

// MODE 0: z arithmetic -> occultquad_hot -> register sum
// MODE 1: z from t[i] + off[k] (global loads) and the walker constants, as in stage 2
// MODE 2: MODE 1 + fluxes parked in shared memory, bins reduced by one lane per timestamp
template <int MODE, int S, bool RELAXED>
__global__ void __launch_bounds__(128) bench(double* out, const double* __restrict__ t, const double* __restrict__ off,
                                             const double* __restrict__ f, const double* __restrict__ w, int n,
                                             int passes, double p, double tcen, double vel, double b, int lo, int hi) {
    __shared__ double scratch[4][2 * 96 + 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double par[6] = {tcen, b, vel, p, 0.34, 0.30};
    const WalkerConsts wc = make_walker_consts(par, S);
    constexpr int TSW = 96 / S, SP = S | 1;
    double acc = 0.0;
    const int gwarp = blockIdx.x * 4 + warp;
    for (int ps = 0; ps < passes; ++ps) {
        // lo < hi: walk the slots of one walker's in-transit range [lo, hi) like stage 2 does
        const int nsl = (hi - lo + TSW - 1) / TSW;
        const int sub = (hi > lo) ? lo + ((gwarp + ps) % nsl) * TSW : ((gwarp * 37 + ps) * TSW) % (n - TSW);
        const int nts = (hi > lo) ? min(TSW, hi - sub) : TSW;
        const int nsamp = nts * S;
        for (int j = lane; j < nsamp; j += 32) {
            const int h = j / S, k = j - h * S;
            double F;
            if (MODE == 3) {  // two samples per lane and trip: half the trips, same number of samples
                if ((j / 32) & 1) continue;
                const double z1 = 0.10 + 0.0008 * ((sub + j * 13) % 977), z2 = 0.10 + 0.0008 * ((sub + (j + 32) * 13) % 977);
                const D2 r = occultquad_interior_x2({z1, z2}, wc.q);
                F = r.a + r.b;
            } else if (MODE == 0) {
                const double z = 0.10 + 0.0008 * ((sub + j * 13) % 977);
                if (!(RELAXED ? occultquad_relaxed(z, wc.q, F) : occultquad_hot(z, wc.q, F))) F = 1.0;
            } else {
                const double ft = __ldg(t + sub + h) + __ldg(off + k);
                const double x = ft - wc.tcen;
                const double vx = wc.vel * x;
                const double q = wc.b2 + vx * vx;
                const double z = FastOps::sqrt(q);
                if (z > wc.q.onep) F = wc.q.fout;
                else if (!(RELAXED ? occultquad_relaxed(z, wc.q, F) : occultquad_hot(z, wc.q, F))) F = 1.0;
            }
            if (MODE == 2) scratch[warp][h * SP + k] = F; else acc += F;
        }
        if (MODE == 2) {
            __syncwarp();
            for (int h = lane; h < nts; h += 32) {
                const double m = bin_mean<S>(scratch[warp] + h * SP);
                const double r = __ldg(f + sub + h) - m;
                acc += __ldg(w + sub + h) * (r * r);
            }
            __syncwarp();
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int S, bool RELAXED = false>
void run(int blocks_per_sm, int passes, const double* t, const double* off, const double* f, const double* w, int n,
         double tcen, double vel, int lo = 0, int hi = 0) {
    double* d;
    const int blocks = 148 * blocks_per_sm;
    cudaMalloc(&d, blocks * 128 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE, S, RELAXED><<<blocks, 128>>>(d, t, off, f, w, n, passes, 0.06, tcen, vel, hi > lo ? 0.5 : 0.3, lo, hi);
    cudaEventRecord(e0);
    bench<MODE, S, RELAXED><<<blocks, 128>>>(d, t, off, f, w, n, passes, 0.06, tcen, vel, hi > lo ? 0.5 : 0.3, lo, hi);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double samples = (double)blocks * 4 * passes * (96 / S) * S;
    if (hi > lo) samples *= (double)(hi - lo) / (((hi - lo + 96 / S - 1) / (96 / S)) * (96 / S));
    printf("%s mode %d S=%d warps/SMSP=%d: %.3f ms  %.2f Gsamples/s  %.0f cycles per 32 samples per SMSP\n",
           RELAXED ? "relaxed" : "strict ", MODE, S, blocks_per_sm, ms, samples / ms / 1e6, ms * 1e-3 * 1.965e9 * 592 / (samples / 32));
    cudaFree(d);
}

int main() {
    const int n = 4096, S = 11;
    double ht[n], hf[n], hw[n], hoff[32];
    const double cad = 0.0204318;
    for (int i = 0; i < n; ++i) { ht[i] = 2061.0 + cad * i; hf[i] = 1.0; hw[i] = -2.0 / (5.4e-5 * 5.4e-5); }
    for (int k = 0; k < S; ++k) hoff[k] = cad * (k * (1.0 / S));
    double *t, *f, *w, *off;
    cudaMalloc(&t, sizeof ht); cudaMalloc(&f, sizeof hf); cudaMalloc(&w, sizeof hw); cudaMalloc(&off, sizeof hoff);
    cudaMemcpy(t, ht, sizeof ht, cudaMemcpyHostToDevice); cudaMemcpy(f, hf, sizeof hf, cudaMemcpyHostToDevice);
    cudaMemcpy(w, hw, sizeof hw, cudaMemcpyHostToDevice); cudaMemcpy(off, hoff, sizeof hoff, cudaMemcpyHostToDevice);
    // a transit so slow that every timestamp of the light curve is interior (z spans 0.3 .. 0.9)
    const double tcen = ht[n / 2], vel = 0.85 / (cad * n / 2);
    const int passes = 400;
    for (int wps : {1, 2, 3, 4}) run<0, S>(wps, passes, t, off, f, w, n, tcen, vel);
    for (int wps : {1, 2, 3, 4}) run<1, S>(wps, passes, t, off, f, w, n, tcen, vel);
    for (int wps : {1, 2, 3, 4}) run<2, S>(wps, passes, t, off, f, w, n, tcen, vel);
    for (int wps : {1, 2, 3, 4, 5, 6}) run<0, S, true>(wps, passes, t, off, f, w, n, tcen, vel);
#ifdef NMB_BENCH_X2
    for (int wps : {1, 2, 3, 4, 5}) run<3, S, true>(wps, passes, t, off, f, w, n, tcen, vel);
#endif
    for (int wps : {1, 2, 3, 4, 5, 6}) run<2, S, true>(wps, passes, t, off, f, w, n, tcen, vel);
    // a realistic walker: b = 0.5, vel = 3 -> ~31 timestamps in range, ingress/egress and margins included
    {
        const double tc = ht[n / 2] + 0.37 * cad, v = 3.0;
        const double half = sqrt(1.06 * 1.06 - 0.25) / v;
        int lo = 0, hi = 0;
        for (int i = 0; i < n; ++i) { if (ht[i] + 0.91 * cad < tc - half) lo = i + 1; if (ht[i] < tc + half) hi = i + 1; }
        lo -= 1; hi += 2;  // the window kernel's grid-cell margins
        printf("realistic walker: range [%d, %d) = %d timestamps\n", lo, hi, hi - lo);
        for (int wps : {1, 2, 3, 4}) run<2, S>(wps, passes, t, off, f, w, n, tc, v, lo, hi);
        for (int wps : {1, 2, 3, 4, 5, 6}) run<2, S, true>(wps, passes, t, off, f, w, n, tc, v, lo, hi);
    }
    return 0;
}
