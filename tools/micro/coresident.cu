// Can two different kernels share an SM?  (development microbenchmark)
// A: 320 threads, big dynamic smem, spins ~50 us.  B: 128 threads, small smem, spins ~20 us.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ long long gt() { long long r; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r)); return r; }
template <int REGS>
__global__ void __launch_bounds__(320, 1) kA(long long* out, int ns, int usebar) {
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    if (usebar && threadIdx.x == 0) {
        unsigned bar = (unsigned)__cvta_generic_to_shared(sm);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    }
    __syncthreads();
    double acc[REGS];
    for (int i = 0; i < REGS; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    while (gt() - t0 < ns) {
        for (int i = 0; i < REGS; ++i) acc[i] = acc[i] * 1.0000001 + 1e-9;
    }
    double s = 0; for (int i = 0; i < REGS; ++i) s += acc[i];
    sm[threadIdx.x] = (unsigned char)s;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)s; }
}
template <int REGS>
__global__ void __launch_bounds__(128) kBr(long long* out, int ns) {
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    double acc[REGS];
    for (int i = 0; i < REGS; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    while (gt() - t0 < ns) {
        for (int i = 0; i < REGS; ++i) acc[i] = acc[i] * 1.0000001 + 1e-9;
    }
    double s = 0; for (int i = 0; i < REGS; ++i) s += acc[i];
    sm[threadIdx.x] = (unsigned char)s;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)s; }
}
template <int MODE>   // 0: mbarrier only; 1: TMA bulk g2s (shared::cluster); 2: + fence.mbarrier_init cluster; 3: bulk s2g store; 4: cp.reduce bulk
__global__ void __launch_bounds__(320, 1) kAtma(long long* out, int ns, const double* src, double* dst) {
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    unsigned bar = (unsigned)__cvta_generic_to_shared(sm);
    unsigned buf = bar + 128;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        if (MODE >= 2) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && MODE >= 1) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf), "l"(src + blockIdx.x * 512), "r"(4096), "r"(bar) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar) : "memory");
        if (MODE >= 3) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (MODE == 3) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + blockIdx.x * 512), "r"(buf), "r"(4096) : "memory");
            else asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst + blockIdx.x * 512), "r"(buf), "r"(4096) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    __syncthreads();
    double a = threadIdx.x;
    while (gt() - t0 < ns) a = a * 1.0000001 + 1e-9;
    sm[4096 + 256 + threadIdx.x] = (unsigned char)a;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)a; }
}
__global__ void __launch_bounds__(128) kBcp(long long* out, int ns, const double* src) {   // B with cp.async (LDGSTS)
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    unsigned dstp = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 8;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dstp), "l"(src + blockIdx.x * 128 + threadIdx.x) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    double a = threadIdx.x;
    while (gt() - t0 < ns) a = a * 1.0000001 + 1e-9;
    sm[2048 + threadIdx.x] = (unsigned char)a;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)a; }
}
struct Big { double v[1536]; };   // 12 KB
__global__ void __launch_bounds__(128) kBbig(long long* out, int ns, const __grid_constant__ Big big) {
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    double a = threadIdx.x;
    while (gt() - t0 < ns) a = a * big.v[threadIdx.x & 1023] + 1e-9;
    sm[threadIdx.x] = (unsigned char)a;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)a; }
}
struct Big6 { double v[768]; };   // 6 KB
__global__ void __launch_bounds__(320, 1) kAbig(long long* out, int ns, int usebar, const __grid_constant__ Big6 big) {
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    double a = threadIdx.x;
    while (gt() - t0 < ns) a = a * big.v[threadIdx.x & 511] + 1e-9;
    sm[threadIdx.x] = (unsigned char)a;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)a; }
}
__global__ void __launch_bounds__(128) kB(long long* out, int ns) {
    extern __shared__ __align__(128) unsigned char sm[];
    long long t0 = gt();
    double a = threadIdx.x;
    while (gt() - t0 < ns) a = a * 1.0000001 + 1e-9;
    sm[threadIdx.x] = (unsigned char)a;
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid(); out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = gt(); out[blockIdx.x*4+3] = (long long)a; }
}
template <class LA, class LB>
void run2(const char* name, LA launchA, LB launchB) {
    long long *dA, *dB; cudaMalloc(&dA, 148 * 32); cudaMalloc(&dB, 148 * 32);
    cudaStream_t sa, sb;
    cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
    for (int rep = 0; rep < 2; ++rep) { launchB(dB, sb); launchA(dA, sa); cudaDeviceSynchronize(); }
    long long hA[148 * 4], hB[148 * 4];
    cudaMemcpy(hA, dA, sizeof hA, cudaMemcpyDeviceToHost); cudaMemcpy(hB, dB, sizeof hB, cudaMemcpyDeviceToHost);
    long long t0 = hB[1]; for (int i = 0; i < 148; ++i) { if (hA[i*4+1] < t0) t0 = hA[i*4+1]; if (hB[i*4+1] < t0) t0 = hB[i*4+1]; }
    int ov = 0; long long maxA = 0, minA = 1LL << 62, endA = 0;
    for (int i = 0; i < 148; ++i) {
        for (int j = 0; j < 148; ++j)
            if (hB[j*4] == hA[i*4] && hB[j*4+1] < hA[i*4+1] && hB[j*4+2] > hA[i*4+1]) { ++ov; break; }
        if (hA[i*4+1] - t0 > maxA) maxA = hA[i*4+1] - t0;
        if (hA[i*4+1] - t0 < minA) minA = hA[i*4+1] - t0;
        if (hA[i*4+2] - t0 > endA) endA = hA[i*4+2] - t0;
    }
    printf("%-44s : A starts %5.1f..%5.1f us, all done %5.1f us, A CTAs co-resident with B: %d/148 (err %s)\n",
           name, minA / 1e3, maxA / 1e3, endA / 1e3, ov, cudaGetErrorString(cudaGetLastError()));
    cudaFree(dA); cudaFree(dB); cudaStreamDestroy(sa); cudaStreamDestroy(sb);
}
template <class KA>
void run(const char* name, KA ka, int smemA, int smemB, bool carve, int usebar, bool prio) {
    long long *dA, *dB; cudaMalloc(&dA, 148 * 32); cudaMalloc(&dB, 148 * 32);
    cudaStream_t sa, sb;
    int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, lo);
    cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, prio ? hi : lo);
    cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, smemA);
    cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB);
    if (carve) {
        cudaFuncSetAttribute(ka, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(kB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    for (int rep = 0; rep < 2; ++rep) {
        kB<<<148, 128, smemB, sb>>>(dB, 20000);
        ka<<<148, 320, smemA, sa>>>(dA, 50000, usebar);
        cudaDeviceSynchronize();
    }
    long long hA[148 * 4], hB[148 * 4];
    cudaMemcpy(hA, dA, sizeof hA, cudaMemcpyDeviceToHost); cudaMemcpy(hB, dB, sizeof hB, cudaMemcpyDeviceToHost);
    long long t0 = hB[1]; for (int i = 0; i < 148; ++i) { if (hA[i*4+1] < t0) t0 = hA[i*4+1]; if (hB[i*4+1] < t0) t0 = hB[i*4+1]; }
    int ov = 0; long long maxA = 0, minA = 1LL << 62, endA = 0;
    for (int i = 0; i < 148; ++i) {
        for (int j = 0; j < 148; ++j)
            if (hB[j*4] == hA[i*4] && hB[j*4+1] < hA[i*4+1] && hB[j*4+2] > hA[i*4+1]) { ++ov; break; }
        if (hA[i*4+1] - t0 > maxA) maxA = hA[i*4+1] - t0;
        if (hA[i*4+1] - t0 < minA) minA = hA[i*4+1] - t0;
        if (hA[i*4+2] - t0 > endA) endA = hA[i*4+2] - t0;
    }
    printf("%-44s smemA=%3dK smemB=%3dK carve=%d bar=%d prio=%d : A starts %5.1f..%5.1f us, all done %5.1f us, A CTAs co-resident with B: %d/148 (err %s)\n",
           name, smemA >> 10, smemB >> 10, (int)carve, usebar, (int)prio, minA / 1e3, maxA / 1e3, endA / 1e3, ov, cudaGetErrorString(cudaGetLastError()));
    cudaFree(dA); cudaFree(dB); cudaStreamDestroy(sa); cudaStreamDestroy(sb);
}
template <int MODE> void tma_case(const char* name, const double* src, double* dst, bool bcp) {
    cudaFuncSetAttribute(kAtma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 142 << 10);
    run2(name, [&](long long* d, cudaStream_t s) { kAtma<MODE><<<148, 320, 141964, s>>>(d, 50000, src, dst); },
         [&](long long* d, cudaStream_t s) { if (bcp) kBcp<<<148, 128, 50176, s>>>(d, 20000, src); else kB<<<148, 128, 50176, s>>>(d, 20000); });
}
int main() {
    cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 << 10);
    cudaFuncSetAttribute(kBcp, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 << 10);
    double *src, *dst; cudaMalloc(&src, 148 * 4096 * 2); cudaMalloc(&dst, 148 * 4096 * 2); cudaMemset(src, 0, 148 * 4096 * 2); cudaMemset(dst, 0, 148 * 4096 * 2);
    tma_case<0>("A mbarrier only", src, dst, false);
    tma_case<1>("A TMA bulk load", src, dst, false);
    tma_case<2>("A TMA bulk load + cluster fence", src, dst, false);
    tma_case<3>("A TMA load + bulk store", src, dst, false);
    tma_case<4>("A TMA load + bulk reduce", src, dst, false);
    tma_case<0>("A mbarrier only, B cp.async", src, dst, true);
    tma_case<4>("A TMA all, B cp.async", src, dst, true);
    return 0;
}
