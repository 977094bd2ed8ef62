// Microbenchmark (round 2): can a DFMA loop be fed its matrix operand from the kernel-parameter constant bank with a
// UNIFORM DYNAMIC index (LDCU.64 UR, c[0x0][UR+imm] -> DFMA R, R, UR, R) at a useful rate while the records stream
// through ~30 KB of parameters?  lanes = poles, C poles per lane, x from shared memory (3C LDS.64 per record).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldcu_stream ldcu_stream.cu ; run: ./ldcu_stream
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NREC = 400;
struct Recs { double h[NREC][9]; short col[NREC]; };
template <int C>
__global__ void __launch_bounds__(256) kern(const __grid_constant__ Recs R, const double* __restrict__ X, double* __restrict__ Y, int ncol, int reps, int nsplit) {
    extern __shared__ double xs[];                 // [ncol*3][32*C] per CTA (shared by its warps: read-only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < ncol * 3 * 32 * C; i += blockDim.x) xs[i] = X[i];
    __syncthreads();
    double acc[3][C];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[m][c] = 0.0;
    const int per = NREC / nsplit, r0 = __reduce_max_sync(0xffffffffu, (warp % nsplit) * per), r1 = r0 + per;   // REDUX -> uniform register: the record index is provably warp-uniform
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int i = r0; i < r1; ++i) {
            const double* xa = xs + R.col[i] * (3 * 32 * C) + lane;
            double x[3][C];
#pragma unroll
            for (int mi = 0; mi < 3; ++mi)
#pragma unroll
                for (int c = 0; c < C; ++c) x[mi][c] = xa[(mi * C + c) * 32];
#pragma unroll
            for (int mi = 0; mi < 3; ++mi)
#pragma unroll
                for (int mo = 0; mo < 3; ++mo)
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[mo][c] = fma(R.h[i][mo * 3 + mi], x[mi][c], acc[mo][c]);
        }
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int c = 0; c < C; ++c) s += acc[m][c];
    Y[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int C>
void run(int nwarps, int ctas_per_sm, int ncol, int nsplit) {
    Recs* h = new Recs;
    for (int i = 0; i < NREC; ++i) { for (int e = 0; e < 9; ++e) h->h[i][e] = 1e-3 * (i + e); h->col[i] = (short)((i * 7) % ncol); }
    double *X, *Y;
    const size_t smem = (size_t)ncol * 3 * 32 * C * 8;
    cudaMalloc(&X, smem); cudaMemset(X, 0, smem);
    const int grid = 148 * ctas_per_sm;
    cudaMalloc(&Y, (size_t)grid * nwarps * 32 * 8);
    cudaFuncSetAttribute(kern<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 200;
    kern<C><<<grid, nwarps * 32, smem>>>(*h, X, Y, ncol, 2, nsplit);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<C><<<grid, nwarps * 32, smem>>>(*h, X, Y, ncol, reps, nsplit);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)grid * nwarps * 32 * (NREC / nsplit) * 9.0 * C * reps;      // lane-DFMAs
    printf("C=%d warps/CTA=%d CTAs/SM=%d nsplit=%d ncol=%d smem=%zu KB: %.3f ms, %.2f TFLOP/s fp64 (%s)\n", C, nwarps, ctas_per_sm, nsplit, ncol, smem / 1024, ms,
           2.0 * dfma / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(X); cudaFree(Y); delete h;
}
int main() {
    // nsplit = number of distinct record streams among the warps of a CTA (1: every warp walks all records)
    for (int nsplit : {1, 2, 4}) {
        run<1>(4, 1, 32, nsplit); run<1>(4, 2, 32, nsplit); run<1>(4, 4, 32, nsplit); run<1>(8, 2, 32, nsplit); run<1>(8, 4, 32, nsplit);
        run<2>(2, 2, 32, nsplit < 2 ? nsplit : 2); run<2>(4, 1, 32, nsplit); run<2>(4, 2, 32, nsplit); run<2>(4, 4, 32, nsplit); run<2>(8, 2, 32, nsplit);
        run<4>(4, 1, 32, nsplit); run<4>(4, 2, 32, nsplit); run<4>(8, 1, 32, nsplit); run<4>(8, 2, 16, nsplit);
    }
    return 0;
}
