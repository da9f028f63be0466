// Does a half-rate FP64 instruction hide other instructions under it?  K integer min ops (VIMNMX) and
// L shared-memory broadcast loads per DFMA, 8 independent DFMA chains per thread, 4..16 warps per SM
// sub-partition.  If DFMA only occupied the FP64 pipe, time would stay flat until K + L ~ 1.2.
#include <cstdio>
template <int K, int L>
__global__ void mix(double* out, int iters, double a, double b) {
    __shared__ double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = threadIdx.x * 1e-3;
    __syncthreads();
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    int m = 0x7fffffff;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = fma(x[i], a, b);
#pragma unroll
            for (int k = 0; k < K; ++k) m = min(m, __double2hiint(x[i]) + k + it);
#pragma unroll
            for (int l = 0; l < L; ++l) acc += 0 * sm[(it + i + l) & 63];
        }
    }
    double s = acc;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678 || m == 77) out[0] = s;
}
template <int K, int L>
void run(int wps) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * wps, threads = 128;
    mix<K, L><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    mix<K, L><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_smsp = (double)wps * iters * 8;   // warp-DFMAs per sub-partition
    printf("K=%d int ops, L=%d LDS per DFMA, %2d warps/SMSP: %.3f ms, %.2f cycles per DFMA per SMSP\n", K, L, wps, ms,
           ms * 1e-3 * 1.965e9 / dfma_per_smsp);
    cudaFree(d);
}
int main() {
    for (int w : {4, 8}) { run<0, 0>(w); run<1, 0>(w); run<2, 0>(w); run<3, 0>(w); run<0, 1>(w); run<1, 1>(w); }
    return 0;
}
