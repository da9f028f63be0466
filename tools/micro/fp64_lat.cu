// FP64 pipe microbenchmark: dependent-issue latency and throughput vs warps/SM and ILP.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, int iters, double a, double b, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 1234.5) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int warps_per_sm, int iters) {
    double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    int threads = 32 * (warps_per_sm < 32 ? warps_per_sm : 32);
    int blocks_per_sm = warps_per_sm <= 32 ? 1 : warps_per_sm / 32;
    int blocks = 148 * blocks_per_sm;
    chain<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9, c);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chain<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9, c);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    double per = (double)cy / iters / ILP;
    double tf = 2.0 * ILP * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    printf("warps/SM=%2d ILP=%d  cycles per DFMA per warp = %.2f (per dependent step %.2f)  %.2f TFLOP/s\n", warps_per_sm, ILP, per, per * ILP, tf);
}
int main() {
    int it = 20000;
    run<1>(4, it); run<2>(4, it); run<4>(4, it); run<8>(4, it);
    run<1>(8, it); run<1>(16, it); run<2>(16, it); run<1>(32, it); run<2>(32, it); run<1>(64, it); run<4>(16, it);
    return 0;
}
