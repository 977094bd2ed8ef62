// Microbenchmark: shared-memory wavefronts of 64-bit / 128-bit loads for the lane-address patterns of the row-tile
// kernel (lanes = poles).  Prints cycles per warp-load at 8 warps per SM (LSU-throughput regime).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_bench tools/lds_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void lds128(double& x, double& y, unsigned a) { asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a)); }

template <int MODE>   // 0: LDS.64, 1: LDS.128
__global__ void __launch_bounds__(256, 1) bench(const int* __restrict__ offs, int base, int iters, long long* out, double* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 200 * 1024 / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = i;
    __syncthreads();
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(sm) + base + offs[threadIdx.x & 31] * 8;
    double acc = 0, acc2 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const unsigned a = s0 + (it & 15) * 5840;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) acc += lds64(a + u * 11680);
            else { double x, y; lds128(x, y, a + u * 11680); acc += x; acc2 += y; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc + acc2 == 1.2345) sink[0] = acc;
}

int main() {
    struct Pat { const char* name; int mode; std::vector<int> o; };
    std::vector<Pat> pats;
    auto mk = [&](const char* nm, int mode, auto f) { Pat p{nm, mode, {}}; for (int j = 0; j < 32; ++j) p.o.push_back(f(j)); pats.push_back(p); };
    mk("stride1", 0, [](int j) { return j; });
    mk("stride2", 0, [](int j) { return 2 * j; });
    mk("stride3", 0, [](int j) { return 3 * j; });
    mk("stride9", 0, [](int j) { return 9 * j; });
    mk("stride27", 0, [](int j) { return 27 * j; });
    mk("stride5", 0, [](int j) { return 5 * j; });
    mk("plainA3", 0, [](int j) { return (j % 3) + 9 * (j / 3); });
    mk("plainA9", 0, [](int j) { return (j % 9) + 27 * (j / 9); });
    mk("plainA27@20", 0, [](int j) { return ((20 + j) % 27) + 81 * ((20 + j) / 27); });
    mk("plainA27@0", 0, [](int j) { return (j % 27) + 81 * (j / 27); });
    mk("bcast64", 0, [](int) { return 0; });
    mk("half-bcast64", 0, [](int j) { return j < 16 ? j : 0; });
    mk("bcast128", 1, [](int) { return 0; });
    mk("contig128", 1, [](int j) { return 2 * j; });
    mk("stride6-128", 1, [](int j) { return 6 * j; });
    mk("pair+stride3: (2 poles adjacent) 128", 1, [](int j) { return 2 * (j % 4) + 27 * (j / 4); });
    int* d_off; long long* d_out; double* d_sink;
    cudaMalloc(&d_off, 128); cudaMalloc(&d_out, 8 * 1024); cudaMalloc(&d_sink, 8);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int iters = 2000;
    for (int base : {0, 64, 8}) {
        for (auto& p : pats) {
            cudaMemcpy(d_off, p.o.data(), 128, cudaMemcpyHostToDevice);
            for (int rep = 0; rep < 2; ++rep) {
                if (p.mode == 0) bench<0><<<148, 256, 220 * 1024>>>(d_off, base, iters, d_out, d_sink);
                else bench<1><<<148, 256, 220 * 1024>>>(d_off, base, iters, d_out, d_sink);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", p.name, cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += (double)h[i];
            avg /= 148;
            printf("base %2d  %-40s cycles per warp-load %.2f\n", base, p.name, avg / ((double)iters * 8 * 8));
        }
    }
    return 0;
}
