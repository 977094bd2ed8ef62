// Microbenchmark: DFMA throughput per SM as a function of the independent chains per thread and the warps per SM
// (how much ILP two warps per scheduler need before the fp64 pipe is the bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_bench tools/dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void k(double* out, int iters, double a, double b) {
    double acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
void run(double* d, int warps) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<N><<<148, warps * 32>>>(d, 10, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<N><<<148, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = 148.0 * warps * 32 * (double)iters * 4 * N;
    printf("chains %2d warps/SM %2d: %6.2f TFLOP/s  (%.2f clk per warp-DFMA per scheduler at 1.965 GHz)\n", N, warps, 2 * dfma / ms / 1e9,
           ms * 1e-3 * 1.965e9 / ((double)iters * 4 * N * (warps / 4.0 < 1 ? 1 : warps / 4.0)));
}

int main() {
    double* d;
    cudaMalloc(&d, 148 * 1024 * 8);
    for (int w : {4, 8, 16}) {
        run<1>(d, w); run<2>(d, w); run<4>(d, w); run<8>(d, w); run<12>(d, w); run<16>(d, w); run<24>(d, w); run<36>(d, w);
    }
    return 0;
}
