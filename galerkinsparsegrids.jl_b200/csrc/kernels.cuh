// kernels.cuh -- sm_100a fp64 kernels of libgsgb200: matrix-free tensor-product sweeps
// y = alpha * D_d x + beta * y over (multi-level, multi-cell, multi-mode) blocks, the RK stage
// updates, batched reconstruct_DG and the CSR cross-check SpMV.
//
// Data model.  For sweep axis d the state decomposes into POLE GROUPS:
// all multi-level blocks that agree in the other D-1 levels.  A group with other-level sum s has
// poles of NQ = 2^p one-dimensional cells (p = n - s; N' = K * NQ entries) and every pole is
// multiplied by the principal sub-block M[0:N', 0:N'] of the 1-D matrix
// (src/multidim_derivative.jl:30-55).  Inside a group an ITEM r fixes the
// other dims' cells; for every 1-D cell q the item owns one contiguous multi-cell of KD = K^D
// doubles at  base[level(q)] + KDp * (lo + S * (c(q) + C(level(q)) * hi)),  r = lo + S * hi.
// Within a multi-cell the entry (a, m_d, b) sits at  a + A*m_d + K*A*b  with A = K^(d-1).
// DEVICE LAYOUT: identical to the reference's vector layout (src/dg_vmethods.jl:48-73) except
// that every multi-cell is padded from KD to KDp = KD rounded up to even, so that every cell
// starts on a 16-byte boundary (TMA bulk copies, 128-bit accesses).
//
// Kernels in this file, by pole length N' = K * 2^p of the class they serve:
//   sweep_stream_kernel<K, PAIR>   N' <= 32   persistent TMA ring (loader / storer / compute warps); PAIR = one
//                                             pass for both directions of a pair over 2-D sub-planes (n' <= 2)
//   sweep_rowtile_kernel<K, C>     N' >= 48   (items with >= 64 poles) persistent; whole multi-cells by TMA, row groups,
//                                             bank-permuted lanes = poles, staged bulk store / reduce-add output
//   sweep_consth_kernel<K, P>      N' = 48    (smaller items) matrix unrolled into constant-bank operands
//   sweep_long2_kernel<K, C, NB>   N' >= 96   (smaller items) register-tiled: lanes = poles, C poles per lane, record stream
//   sweep_short_tma_kernel (small k^D: many multi-cells per tile), sweep_generic_kernel (any k <= 10, any length)
//   rk_stage_kernel, rk_final_kernel, rk4_taylor_kernel(_cells)   RK4 updates
//   sweep_flat_kernel<K>           small index sets: one launch per right-hand side, all directions, no atomics
//   reconstruct2_kernel (reconstruct_kernel)   batched reconstruct_DG;   spmv_csr_kernel   cross-check SpMV
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

namespace gsgk {

constexpr int MAXL = 16;

struct GroupDev {
    long long base[MAXL + 1];  // block offset for level_d = 0..p
    int p;                     // pole has 2^p cells
    int S;                     // prod of cells of dims < d
    int nitems;                // S * prod of cells of dims > d
    int pad;
};

struct TileDev {
    int group;
    int r0;     // first item
    int nr;     // number of consecutive items
    int ebase;  // in-cell offset of the tile's pole sub-range (0 for whole-item tiles)
};

struct Bcsr {                 // K x K block CSR of the 1-D matrix, block columns ascending
    const int* rowptr;
    const int* col;
    const double* val;        // blocks of KK2 doubles, row-major (m_out, m_in)
    int KK2;                  // K*K rounded up to even (16-byte aligned blocks)
};

__device__ __forceinline__ void q_decode(int q, int& ld, int& cd, int& Cd) {
    ld = q == 0 ? 0 : 32 - __clz(q);
    cd = q == 0 ? 0 : q - (1 << (ld - 1));
    Cd = ld <= 1 ? 1 : 1 << (ld - 1);
}

// KDp = padded multi-cell stride of the device layout (K^D rounded up to even)
__device__ __forceinline__ long long cell_addr(const long long* base, int S, int q, int r, int KDp) {
    int ld, cd, Cd;
    q_decode(q, ld, cd, Cd);
    const int lo = r % S, hi = r / S;
    return base[ld] + (long long)KDp * (lo + (long long)S * (cd + (long long)Cd * hi));
}

// ------------------------------------------------------------------------------------------
// Short poles, Blackwell-native path: ONE persistent, warp-specialised kernel for all short
// classes of a sweep.  Warp 0 is the TMA producer: it streams tiles (CT multi-cells each, every
// cell one contiguous 16-byte-aligned run) into an NS-deep shared-memory ring with
// cp.async.bulk + mbarrier complete_tx, and streams finished tiles back with bulk stores -- or,
// for y += alpha*D_d x, with cp.reduce.async.bulk.add.f64 so y is never loaded into the SM.
// The 8 compute warps wait on the tile's `full` barrier, own one pole per thread (registers),
// multiply by the dense N' x N' block (broadcast shared-memory loads), write alpha*result in
// place, fence to the async proxy and arrive on the tile's `done` barrier.
// ------------------------------------------------------------------------------------------
struct TileS {                // self-contained: the producer never touches the group table
    long long base[4];        // block offsets for level_d = 0..3 (register-resident classes: P <= 3)
    int S;                    // prod of cells of dims < d
    int r0;
    short nr;
    short P;
    int pad;
};

struct ShortParams {
    int KD, KDp, A;
    int stage_doubles;  // doubles per ring stage
    int nstage;
};

__device__ __forceinline__ unsigned smid_u32() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ long long gtimer() { long long r; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r)); return r; }

namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_red_add_f64(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace tma

// Dense principal sub-blocks of the 1-D matrix for the register-resident classes, passed as a
// __grid_constant__ kernel parameter: the fully unrolled DFMAs read H as immediate constant-bank
// operands (c[0][imm]), so the matrix costs neither shared-memory bandwidth nor registers.
template <int K>
struct ShortDims {
    __host__ __device__ static constexpr int pmax() {
        int p = 0;
        while (p < 3 && (K << (p + 1)) <= 32) ++p;
        return p;
    }
    __host__ __device__ static constexpr int hoff(int P) {          // offset of class P's block (each padded to even)
        int o = 0;
        for (int p = 0; p < P; ++p) o += (((K << p) * (K << p)) + 1) & ~1;
        return o;
    }
    __host__ __device__ static constexpr int htotal() { return hoff(pmax() + 1); }
};

template <int K>
struct HDense {
    double v[ShortDims<K>::htotal()];
};

// One pole per thread: x[N'] in registers, rows accumulated RC at a time (RC independent DFMA
// chains), results alpha*acc written back in place.
template <int K, int P>
__device__ __forceinline__ void short_tile_compute(double* xs, const HDense<K>& hd, int nr, int KDp, int A, int PI,
                                                   const int* pole_off, double alpha, int ctid, int ncth) {
    constexpr int NQ = 1 << P, NP = K * NQ;
    constexpr int RC = NP < 12 ? NP : 12;        // rows per accumulation chunk
    constexpr int HO = ShortDims<K>::hoff(P);
    const size_t qstride = (size_t)nr * KDp;
    for (int r = 0; r < nr; ++r) {
        for (int j = ctid; j < PI; j += ncth) {
            double* pole = xs + (size_t)r * KDp + pole_off[j];
            double x[NP];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int m = 0; m < K; ++m) x[q * K + m] = pole[q * qstride + A * m];
#pragma unroll
            for (int i0 = 0; i0 < NP; i0 += RC) {
                double acc[RC];
#pragma unroll
                for (int i = 0; i < RC; ++i) acc[i] = 0.0;
#pragma unroll
                for (int jx = 0; jx < NP; ++jx)
#pragma unroll
                    for (int i = 0; i < RC; ++i)
                        if (i0 + i < NP) acc[i] = fma(hd.v[HO + (i0 + i) * NP + jx], x[jx], acc[i]);
#pragma unroll
                for (int i = 0; i < RC; ++i)
                    if (i0 + i < NP) {
                        const int q = (i0 + i) / K, m = (i0 + i) % K;
                        pole[q * qstride + A * m] = alpha * acc[i];
                    }
            }
        }
    }
}

constexpr int SHORT_TMA_COMPUTE_WARPS = 8;

// producer side (whole warp 0): lane q moves the multi-cells of 1-D cell q, one bulk copy per run
// of memory-contiguous items (items are contiguous while lo = r % S does not wrap).
__device__ __forceinline__ void short_tma_load(const double* __restrict__ X, const TileS& t, double* dst, unsigned bar,
                                               int KDp, int lane) {
    const int NQ = 1 << t.P, S = t.S, nr = t.nr;
    if (lane == 0) tma::mbar_expect_tx(bar, (unsigned)(NQ * nr * KDp * 8));
    __syncwarp();
    if (lane < NQ) {
        const int q = lane;
        int r = 0;
        while (r < nr) {
            const int lo = (t.r0 + r) % S;
            const int run = min(nr - r, S - lo);
            const double* src = X + cell_addr(t.base, S, q, t.r0 + r, KDp);
            tma::bulk_g2s(tma::smem_u32(dst + (size_t)(q * nr + r) * KDp), src, (unsigned)(run * KDp * 8), bar);
            r += run;
        }
    }
}

__device__ __forceinline__ void short_tma_store(double* __restrict__ Y, const TileS& t, const double* srcs, int KDp,
                                                int lane, int accumulate) {
    const int NQ = 1 << t.P, S = t.S, nr = t.nr;
    if (lane < NQ) {
        const int q = lane;
        int r = 0;
        while (r < nr) {
            const int lo = (t.r0 + r) % S;
            const int run = min(nr - r, S - lo);
            double* dstg = Y + cell_addr(t.base, S, q, t.r0 + r, KDp);
            const unsigned sa = tma::smem_u32(srcs + (size_t)(q * nr + r) * KDp);
            if (accumulate) tma::bulk_red_add_f64(dstg, sa, (unsigned)(run * KDp * 8));
            else tma::bulk_s2g(dstg, sa, (unsigned)(run * KDp * 8));
            r += run;
        }
    }
    tma::bulk_commit();
}

template <int K>
__global__ void __launch_bounds__(32 * (SHORT_TMA_COMPUTE_WARPS + 1), 1)
sweep_short_tma_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, int accumulate,
                       const GroupDev* __restrict__ groups, const TileS* __restrict__ tiles, int ntiles,
                       const __grid_constant__ HDense<K> hd, const ShortParams prm,
                       int* __restrict__ counter,       // dynamic tile scheduler (zeroed before the launch)
                       long long* __restrict__ dbg) {   // dbg: optional per-phase clock stamps (CTA 0)
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smraw);     // full[NS], done[NS]
    TileS* sdesc = reinterpret_cast<TileS*>(smraw + 128);                        // per-stage tile descriptor
    const int PI = prm.KD / K;
    int* pole_off = reinterpret_cast<int*>(smraw + 128 + 8 * sizeof(TileS));
    size_t off = 128 + 8 * sizeof(TileS) + (size_t)PI * 4;
    off = (off + 127) & ~(size_t)127;
    double* ring = reinterpret_cast<double*>(smraw + off);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int NS = prm.nstage;
    const int KDp = prm.KDp;

    for (int j = tid; j < PI; j += blockDim.x) {
        const int b = j / prm.A, a = j - b * prm.A;
        pole_off[j] = a + K * prm.A * b;
    }
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            tma::mbar_init(tma::smem_u32(&bars[s]), 1);
            tma::mbar_init(tma::smem_u32(&bars[NS + s]), SHORT_TMA_COMPUTE_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================= TMA producer / storer (warp 0, lanes = 1-D cells) =================
        // Tiles are claimed with an atomic counter (the CTAs of this persistent kernel may start at
        // different times because long-pole CTAs of the same sweep share the SMs).  A claimed tile's
        // descriptor is parked in shared memory for the compute warps; P = -1 marks the end.
        auto claim_and_load = [&](int s) -> bool {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(counter, 1);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            const unsigned bar = tma::smem_u32(&bars[s]);
            if (idx >= ntiles) {
                if (lane == 0) {
                    sdesc[s].P = -1;
                    tma::mbar_arrive(bar);          // release the compute warps without data
                }
                __syncwarp();
                return false;
            }
            const TileS t = tiles[idx];
            if (lane == 0) sdesc[s] = t;
            __syncwarp();
            short_tma_load(X, t, ring + (size_t)s * prm.stage_doubles, bar, KDp, lane);
            return true;
        };
        int inflight = 0;
        bool more = true;
        for (int s = 0; s < NS && more; ++s) {
            more = claim_and_load(s);
            if (more) ++inflight;
        }
        for (int it = 0; inflight > 0; ++it) {
            const int s = it % NS;
            const unsigned ph = (unsigned)((it / NS) & 1);
            const long long c0 = clock64();
            tma::mbar_wait(tma::smem_u32(&bars[NS + s]), ph);          // tile computed
            const long long c1 = clock64();
            const TileS tcur = sdesc[s];
            short_tma_store(Y, tcur, ring + (size_t)s * prm.stage_doubles, KDp, lane, accumulate);
            --inflight;
            const long long c2 = clock64();
            long long c3 = c2;
            if (more) {
                tma::bulk_wait_read0();        // the stage may be overwritten once its stores have read it
                __syncwarp();
                c3 = clock64();
                more = claim_and_load(s);
                if (more) ++inflight;
            }
            if (dbg && blockIdx.x == 0 && it < 64 && lane == 0) {
                dbg[it * 8 + 0] = c0; dbg[it * 8 + 1] = c1; dbg[it * 8 + 2] = c2; dbg[it * 8 + 3] = c3;
                dbg[it * 8 + 4] = clock64();
            }
        }
        tma::bulk_wait0();
    } else {
        // ================= compute warps =================
        const int ctid = tid - 32, ncth = 32 * SHORT_TMA_COMPUTE_WARPS;
        for (int it = 0;; ++it) {
            const int s = it % NS;
            const unsigned ph = (unsigned)((it / NS) & 1);
            const long long w0 = clock64();
            tma::mbar_wait(tma::smem_u32(&bars[s]), ph);
            const long long w1 = clock64();
            const int P = sdesc[s].P, nr = sdesc[s].nr;
            if (P < 0) break;
            double* xs = ring + (size_t)s * prm.stage_doubles;
            switch (P) {
                case 0: short_tile_compute<K, 0>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
                case 1: if constexpr (ShortDims<K>::pmax() >= 1) short_tile_compute<K, 1>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
                case 2: if constexpr (ShortDims<K>::pmax() >= 2) short_tile_compute<K, 2>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
                case 3: if constexpr (ShortDims<K>::pmax() >= 3) short_tile_compute<K, 3>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
            }
            tma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(tma::smem_u32(&bars[NS + s]));
            if (dbg && blockIdx.x == 0 && it < 64 && ctid == 0) {
                dbg[it * 8 + 5] = w0; dbg[it * 8 + 6] = w1; dbg[it * 8 + 7] = clock64();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Streaming kernel, second generation (the shipped path): same data flow as above, but the
// producer is split into a LOADER warp and a STORER warp so that loads, stores and the wait for a
// store's shared-memory read-out no longer serialise in one instruction stream (measured with the
// clock stamps: the single producer spent ~5 500 cycles per tile, the compute warps ~2 000).
//  * tile descriptors carry ready-made runs {first multi-cell, count, position in the stage}:
//    no address arithmetic (divisions) in the producer, one bulk copy per run;
//  * the loader claims the NEXT tile (atomic counter) and fetches its descriptor while it waits
//    for a free stage, so the claim round trip is off the critical path;
//  * the storer frees a stage one iteration late (cp.async.bulk.wait_group.read 1), so a store's
//    read-out overlaps the next store's issue.
// Barriers per stage: full (loader -> compute, tx bytes), done (compute -> storer), empty
// (storer -> loader).
// ------------------------------------------------------------------------------------------
constexpr int TILE2_MAXRUN = 16;

struct TileS2 {                // 144 bytes = 36 words, fetched one word per lane
    int nruns;
    int ncell;                 // multi-cells in the tile (= nr << P)
    short P, nr;
    int pad;
    struct Run {
        int cell0;             // first multi-cell (global index; offset = cell0 * KDp doubles)
        short scell;           // position in the stage (multi-cell units)
        short n;               // consecutive multi-cells
    } run[TILE2_MAXRUN];
};
static_assert(sizeof(TileS2) == 144, "TileS2 layout");
constexpr int TILE2_WORDS = sizeof(TileS2) / 4;

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// compute one tile: one pole per thread and item; U items are processed together for ILP
template <int K, int P, int U>
__device__ __forceinline__ void short_tile_compute2(double* xs, const HDense<K>& hd, int nr, int KDp, int A, int PI,
                                                    const int* pole_off, double alpha, int ctid, int ncth) {
    constexpr int NQ = 1 << P, NP = K * NQ;
    constexpr int RC = NP < 12 ? NP : 12;        // rows per accumulation chunk
    constexpr int HO = ShortDims<K>::hoff(P);
    const size_t qstride = (size_t)nr * KDp;
    for (int r0 = 0; r0 < nr; r0 += U) {
        for (int j = ctid; j < PI; j += ncth) {
            double* pole[U];
            double x[U][NP];
#pragma unroll
            for (int uu = 0; uu < U; ++uu) {
                const int r = r0 + uu < nr ? r0 + uu : nr - 1;     // tail: recompute the last item (idempotent)
                pole[uu] = xs + (size_t)r * KDp + pole_off[j];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int m = 0; m < K; ++m) x[uu][q * K + m] = pole[uu][q * qstride + A * m];
            }
#pragma unroll
            for (int i0 = 0; i0 < NP; i0 += RC) {
                double acc[U][RC];
#pragma unroll
                for (int uu = 0; uu < U; ++uu)
#pragma unroll
                    for (int i = 0; i < RC; ++i) acc[uu][i] = 0.0;
#pragma unroll
                for (int jx = 0; jx < NP; ++jx)
#pragma unroll
                    for (int i = 0; i < RC; ++i)
                        if (i0 + i < NP) {
#pragma unroll
                            for (int uu = 0; uu < U; ++uu)
                                acc[uu][i] = fma(hd.v[HO + (i0 + i) * NP + jx], x[uu][jx], acc[uu][i]);
                        }
#pragma unroll
                for (int uu = 0; uu < U; ++uu) {
                    if (uu > 0 && r0 + uu >= nr) continue;          // tail duplicate: nothing to write
#pragma unroll
                    for (int i = 0; i < RC; ++i)
                        if (i0 + i < NP) {
                            const int q = (i0 + i) / K, m = (i0 + i) % K;
                            pole[uu][q * qstride + A * m] = alpha * acc[uu][i];
                        }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Direction-pair fusion for the streaming classes.  For the pair (a, b = a+1) and fixed cells of the other
// D-2 dimensions, the blocks with level_a + level_b <= n' (n' = n - sum of the other levels) form a 2-D
// sparse sub-plane that is closed under the poles of BOTH directions; for n' <= 2 it has 1 / 3 / 8 multi-cells,
// i.e. it fits one ring stage.  A PAIR tile holds whole sub-planes: y = alpha_a D_a x + alpha_b D_b x is formed
// from ONE load of x and ONE store / reduce-add of y (the unfused sweeps move x twice and y three times).
// One thread owns the k x k x (cells of the sub-plane) "mini-grid" of one combination of the other modes: all of
// it sits in registers, so results are written back in place without any barrier.
// Slot order of a sub-plane (shared with the host): for q_b = 0.., for q_a = 0.. with lvl(q_a) + lvl(q_b) <= n'.
// ------------------------------------------------------------------------------------------
namespace pairp {
__host__ __device__ constexpr int lvl(int q) { return q == 0 ? 0 : (q == 1 ? 1 : (q < 4 ? 2 : 3)); }
__host__ __device__ constexpr int ncell(int np) { return np == 0 ? 1 : (np == 1 ? 3 : 8); }
// first slot of row q_b (branch-free in the optimiser's eyes: the unrolled loops must fold these to constants)
__host__ __device__ constexpr int rowstart(int np, int qb) {
    return np == 2 ? (qb == 0 ? 0 : (qb == 1 ? 4 : qb + 4)) : (np == 1 ? (qb == 0 ? 0 : 2) : 0);
}
__host__ __device__ constexpr int slot(int np, int qa, int qb) {      // -1 if (qa, qb) is not in the sub-plane
    return (qa < (1 << np) && qb < (1 << np) && lvl(qa) + lvl(qb) <= np) ? rowstart(np, qb) + qa : -1;
}
}  // namespace pairp

template <int K, int NPR, int QA, int QB>
__device__ __forceinline__ void pair_cell_out(const HDense<K>& hd, const double (&x)[pairp::ncell(NPR)][K][K],
                                              double* base, const int (&cofs)[pairp::ncell(NPR)], int Aa, int Ab,
                                              double alpha_a, double alpha_b) {
    constexpr int S = pairp::slot(NPR, QA, QB);
    if constexpr (S >= 0) {
        constexpr int PA = NPR - pairp::lvl(QB);          // pole class along a in row q_b
        constexpr int PB = NPR - pairp::lvl(QA);          // pole class along b in column q_a
        constexpr int NA = K << PA, NB = K << PB;
        constexpr int HOA = ShortDims<K>::hoff(PA), HOB = ShortDims<K>::hoff(PB);
#pragma unroll
        for (int ma = 0; ma < K; ++ma)
#pragma unroll
            for (int mb = 0; mb < K; ++mb) {
                double sa = 0.0, sb = 0.0;
#pragma unroll
                for (int qa2 = 0; qa2 < (1 << PA); ++qa2)
#pragma unroll
                    for (int m2 = 0; m2 < K; ++m2)
                        sa = fma(hd.v[HOA + (QA * K + ma) * NA + qa2 * K + m2], x[pairp::slot(NPR, qa2, QB)][m2][mb], sa);
#pragma unroll
                for (int qb2 = 0; qb2 < (1 << PB); ++qb2)
#pragma unroll
                    for (int m2 = 0; m2 < K; ++m2)
                        sb = fma(hd.v[HOB + (QB * K + mb) * NB + qb2 * K + m2], x[pairp::slot(NPR, QA, qb2)][ma][m2], sb);
                base[cofs[S] + Aa * ma + Ab * mb] = fma(alpha_a, sa, alpha_b * sb);
            }
    }
}

template <int K, int NPR, int QB, int... QAs>
__device__ __forceinline__ void pair_row_out(const HDense<K>& hd, const double (&x)[pairp::ncell(NPR)][K][K], double* base,
                                             const int (&cofs)[pairp::ncell(NPR)], int Aa, int Ab, double alpha_a,
                                             double alpha_b, std::integer_sequence<int, QAs...>) {
    (pair_cell_out<K, NPR, QAs, QB>(hd, x, base, cofs, Aa, Ab, alpha_a, alpha_b), ...);
}

template <int K, int NPR, int... QBs>
__device__ __forceinline__ void pair_all_out(const HDense<K>& hd, const double (&x)[pairp::ncell(NPR)][K][K], double* base,
                                             const int (&cofs)[pairp::ncell(NPR)], int Aa, int Ab, double alpha_a,
                                             double alpha_b, std::integer_sequence<int, QBs...>) {
    (pair_row_out<K, NPR, QBs>(hd, x, base, cofs, Aa, Ab, alpha_a, alpha_b, std::make_integer_sequence<int, (1 << NPR)>{}), ...);
}

// n' = 2 sub-planes: split a mini-grid's output rows over two warps (needs a named barrier among the compute
// warps between the reads and the in-place writes).  Compile with -DGSG_PAIR_SPLIT_ROWS=0 for the barrier-free
// variant (one thread per mini-grid, all 72 outputs): ptxas then spills 728 bytes even at 255 registers.
#ifndef GSG_PAIR_SPLIT_ROWS
#define GSG_PAIR_SPLIT_ROWS 1
#endif
constexpr bool PAIR_SPLIT_ROWS = GSG_PAIR_SPLIT_ROWS != 0;

// one PAIR tile: nr sub-planes (items); cell (slot s, item r) sits at xs + (s*nr + r)*KDp
template <int K, int NPR>
__device__ __forceinline__ void pair_tile_compute(double* xs, const HDense<K>& hd, int nr, int KDp, int Aa, int NO,
                                                  const int* pair_off, double alpha_a, double alpha_b, int ctid,
                                                  int ncth) {
    constexpr int NC = pairp::ncell(NPR);
    const int Ab = K * Aa;
    if constexpr (NPR == 2 && PAIR_SPLIT_ROWS) {
        // 8-cell sub-planes: one item per tile and only NO = k^(D-2) mini-grids, each 1188 DFMAs at k = 3 -- two
        // warps share a mini-grid group (rows q_b = 0 / q_b >= 1 of the outputs) so that six of the eight warps work
        const int wid = ctid >> 5, lane = ctid & 31, nw = ncth >> 5;     // nw is even: the halves (2m, 2m+1) of a
        const int ngroups = (NO + 31) / 32;                               // group always run in the same round
        const int total = nr * ngroups * 2;
        for (int u0 = 0; u0 < total; u0 += nw) {
            const int u = u0 + wid;
            const int part = u & 1, og = (u >> 1) % ngroups, r = (u >> 1) / ngroups;
            const int o = og * 32 + lane;
            const bool valid = u < total && o < NO;
            double* base = xs + (size_t)(valid ? r : 0) * KDp + pair_off[valid ? o : 0];
            int cofs[NC];
#pragma unroll
            for (int sidx = 0; sidx < NC; ++sidx) cofs[sidx] = sidx * nr * KDp;
            double x[NC][K][K];
            if (valid) {
#pragma unroll
                for (int sidx = 0; sidx < NC; ++sidx)
#pragma unroll
                    for (int ma = 0; ma < K; ++ma)
#pragma unroll
                        for (int mb = 0; mb < K; ++mb) x[sidx][ma][mb] = base[cofs[sidx] + Aa * ma + Ab * mb];
            }
            // the two halves of a mini-grid are written by different warps: every compute thread passes this
            // barrier (unconditionally) after reading its x and before any result of the round lands
            asm volatile("bar.sync 1, %0;" ::"r"(ncth) : "memory");
            if (valid) {
                if (part == 0)
                    pair_all_out<K, NPR>(hd, x, base, cofs, Aa, Ab, alpha_a, alpha_b, std::integer_sequence<int, 0>{});
                else
                    pair_all_out<K, NPR>(hd, x, base, cofs, Aa, Ab, alpha_a, alpha_b, std::integer_sequence<int, 1, 2, 3>{});
            }
        }
        return;
    }
    for (int g = ctid; g < nr * NO; g += ncth) {
        const int r = g / NO, o = g - r * NO;
        double* base = xs + (size_t)r * KDp + pair_off[o];
        int cofs[NC];
#pragma unroll
        for (int sidx = 0; sidx < NC; ++sidx) cofs[sidx] = sidx * nr * KDp;
        double x[NC][K][K];
#pragma unroll
        for (int sidx = 0; sidx < NC; ++sidx)
#pragma unroll
            for (int ma = 0; ma < K; ++ma)
#pragma unroll
                for (int mb = 0; mb < K; ++mb) x[sidx][ma][mb] = base[cofs[sidx] + Aa * ma + Ab * mb];
        pair_all_out<K, NPR>(hd, x, base, cofs, Aa, Ab, alpha_a, alpha_b, std::make_integer_sequence<int, (1 << NPR)>{});
    }
}

// compute warps: 8 for the single-direction tiles; 6 for PAIR tiles (256 threads leave 255 registers per thread:
// the 72-value mini-grid of an 8-cell sub-plane then stays in registers; with 320 threads the cap is 168 and
// ptxas spilled 560 bytes)
template <bool PAIR>
struct StreamCfg {
    static constexpr int CW = PAIR ? 6 : 8;
    static constexpr int THREADS = 32 * (CW + 2);
};
constexpr int STREAM_COMPUTE_WARPS = 8;
constexpr int STREAM_THREADS = 32 * (STREAM_COMPUTE_WARPS + 2);

template <int K, bool PAIR>
__global__ void __launch_bounds__(StreamCfg<PAIR>::THREADS, 1)
sweep_stream_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, double alpha_b, int accumulate,
                    const TileS2* __restrict__ tiles, int ntiles,
                    const __grid_constant__ HDense<K> hd, const ShortParams prm,
                    int* __restrict__ counter,       // dynamic tile scheduler (zeroed before the launch)
                    long long* __restrict__ dbg) {   // dbg: optional per-phase clock stamps (CTA 0)
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smraw);     // full[NS], done[NS], empty[NS]
    TileS2* sdesc = reinterpret_cast<TileS2*>(smraw + 128);                      // per-stage tile descriptor
    const int PI = prm.KD / K;            // poles per item; PAIR: the table below holds PI / K mini-grid offsets
    int* pole_off = reinterpret_cast<int*>(smraw + 128 + 4 * sizeof(TileS2));
    size_t off = 128 + 4 * sizeof(TileS2) + (size_t)PI * 4;
    off = (off + 127) & ~(size_t)127;
    double* ring = reinterpret_cast<double*>(smraw + off);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int NS = prm.nstage;            // <= 4
    const long long g_t0 = dbg ? gtimer() : 0;
    const int KDp = prm.KDp;
    const unsigned cell_bytes = (unsigned)KDp * 8u;

    if constexpr (PAIR) {                 // other-mode combination o -> in-cell offset (modes a, b = a+1 taken out)
        for (int j = tid; j < PI / K; j += blockDim.x) {
            const int b = j / prm.A, a = j - b * prm.A;
            pole_off[j] = a + K * K * prm.A * b;
        }
    } else {
        for (int j = tid; j < PI; j += blockDim.x) {
            const int b = j / prm.A, a = j - b * prm.A;
            pole_off[j] = a + K * prm.A * b;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            tma::mbar_init(tma::smem_u32(&bars[s]), 32);      // every loader lane writes descriptor words and arrives
            // "computed" and "free": every lane arrives for itself (under the PTX model lane 0's arrive after a
            // __syncwarp would do, but compute-sanitizer's racecheck credits only a thread's own arrive)
            tma::mbar_init(tma::smem_u32(&bars[4 + s]), 32 * StreamCfg<PAIR>::CW);
            tma::mbar_init(tma::smem_u32(&bars[8 + s]), 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================= loader =================
        auto claim = [&]() -> int {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(counter, 1);
            return __shfl_sync(0xffffffffu, idx, 0);
        };
        auto fetch = [&](int idx, int& w0, int& w1) {       // descriptor words lane and lane + 32
            w0 = 0; w1 = 0;
            if (idx < ntiles) {
                const int* src = reinterpret_cast<const int*>(tiles + idx);
                w0 = __ldg(src + lane);
                if (lane + 32 < TILE2_WORDS) w1 = __ldg(src + lane + 32);
            }
        };
        int idx = claim();
        int w0, w1;
        fetch(idx, w0, w1);
        for (int it = 0;; ++it) {
            const int s = it % NS;
            const long long c0 = clock64();
            if (it >= NS) tma::mbar_wait(tma::smem_u32(&bars[8 + s]), (unsigned)(((it / NS) - 1) & 1));
            const long long c1 = clock64();
            int* dw = reinterpret_cast<int*>(&sdesc[s]);
            const unsigned bar = tma::smem_u32(&bars[s]);
            if (idx >= ntiles) {
                if (lane == 0) sdesc[s].P = -1;
                tma::mbar_arrive(bar);              // release the compute warps without data (all 32 lanes arrive)
                break;
            }
            dw[lane] = w0;
            if (lane + 32 < TILE2_WORDS) dw[lane + 32] = w1;
            __syncwarp();
            const TileS2& t = sdesc[s];
            // each lane's arrive releases the descriptor words it wrote; lane 0's also posts the byte count
            if (lane == 0) tma::mbar_expect_tx(bar, (unsigned)t.ncell * cell_bytes);
            else tma::mbar_arrive(bar);
            __syncwarp();
            if (lane < t.nruns) {
                const TileS2::Run rn = t.run[lane];
                tma::bulk_g2s(tma::smem_u32(ring + (size_t)s * prm.stage_doubles + (size_t)rn.scell * KDp),
                              X + (size_t)rn.cell0 * KDp, (unsigned)rn.n * cell_bytes, bar);
            }
            const long long c2 = clock64();
            idx = claim();                          // next tile: claimed and fetched while this one is in flight
            fetch(idx, w0, w1);
            if (dbg && blockIdx.x == 0 && it < 64 && lane == 0) {
                dbg[it * 8 + 0] = c0; dbg[it * 8 + 1] = c1; dbg[it * 8 + 2] = c2;
            }
        }
    } else if (warp == 1) {
        // ================= storer =================
        int it = 0;
        for (;; ++it) {
            const int s = it % NS;
            const long long c0 = clock64();
            tma::mbar_wait(tma::smem_u32(&bars[4 + s]), (unsigned)((it / NS) & 1));          // tile computed
            const long long c1 = clock64();
            const TileS2& t = sdesc[s];
            if (t.P < 0) break;
            if (lane < t.nruns) {
                const TileS2::Run rn = t.run[lane];
                double* dstg = Y + (size_t)rn.cell0 * KDp;
                const unsigned sa = tma::smem_u32(ring + (size_t)s * prm.stage_doubles + (size_t)rn.scell * KDp);
                if (accumulate) tma::bulk_red_add_f64(dstg, sa, (unsigned)rn.n * cell_bytes);
                else tma::bulk_s2g(dstg, sa, (unsigned)rn.n * cell_bytes);
            }
            tma::bulk_commit();
            if (it >= 1) {
                bulk_wait_read1();                  // the previous tile's stage has been read out
                __syncwarp();
                tma::mbar_arrive(tma::smem_u32(&bars[8 + (it - 1) % NS]));
            }
            if (dbg && blockIdx.x == 0 && it < 64 && lane == 0) {
                dbg[it * 8 + 3] = c0; dbg[it * 8 + 4] = c1; dbg[it * 8 + 5] = clock64();
            }
        }
        tma::bulk_wait0();                          // every store has completed before the CTA exits
        if (dbg && lane == 0 && blockIdx.x < 512) {
            long long* o = dbg + 4096 + blockIdx.x * 4;
            o[0] = smid_u32(); o[1] = g_t0; o[2] = gtimer(); o[3] = it;
        }
    } else {
        // ================= compute warps =================
        const int ctid = tid - 64, ncth = 32 * StreamCfg<PAIR>::CW;
        for (int it = 0;; ++it) {
            const int s = it % NS;
            const long long w0 = clock64();
            tma::mbar_wait(tma::smem_u32(&bars[s]), (unsigned)((it / NS) & 1));
            const long long w1 = clock64();
            const int P = sdesc[s].P, nr = sdesc[s].nr;
            if (P < 0) {
                tma::mbar_arrive(tma::smem_u32(&bars[4 + s]));      // pass the end marker on
                break;
            }
            double* xs = ring + (size_t)s * prm.stage_doubles;
            if constexpr (PAIR) {
                switch (P) {              // P = n' of the sub-planes in this tile
                    case 0: pair_tile_compute<K, 0>(xs, hd, nr, KDp, prm.A, PI / K, pole_off, alpha, alpha_b, ctid, ncth); break;
                    case 1: if constexpr (ShortDims<K>::pmax() >= 1) pair_tile_compute<K, 1>(xs, hd, nr, KDp, prm.A, PI / K, pole_off, alpha, alpha_b, ctid, ncth); break;
                    case 2: if constexpr (ShortDims<K>::pmax() >= 2 && K <= 3) pair_tile_compute<K, 2>(xs, hd, nr, KDp, prm.A, PI / K, pole_off, alpha, alpha_b, ctid, ncth); break;
                }
            } else
            switch (P) {
                case 0: short_tile_compute2<K, 0, 4>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
                case 1: if constexpr (ShortDims<K>::pmax() >= 1) short_tile_compute2<K, 1, 2>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
                case 2: if constexpr (ShortDims<K>::pmax() >= 2) short_tile_compute2<K, 2, 2>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
                case 3: if constexpr (ShortDims<K>::pmax() >= 3) short_tile_compute2<K, 3, 1>(xs, hd, nr, KDp, prm.A, PI, pole_off, alpha, ctid, ncth); break;
            }
            tma::fence_proxy_async();
            tma::mbar_arrive(tma::smem_u32(&bars[4 + s]));
            if (dbg && blockIdx.x == 0 && it < 64 && ctid == 0) {
                dbg[it * 8 + 6] = w1 - w0; dbg[it * 8 + 7] = clock64() - w1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Generic poles (any p, any K): a tile holds PT = nr * NPOLE poles; x is staged TRANSPOSED in
// shared memory as xs[row = q*K+m][pole] so that lanes = poles read conflict-free; each thread
// computes one block-row (K outputs) of one pole from the K x K block-CSR matrix (uniform,
// L1/L2-resident loads), writes ys, and the tile is streamed back coalesced.
//   TL = K * NPOLE contiguous-or-strided entries per (cell, item):
//   t -> a = t % Amin, m = (t / Amin) % K, bl = t / (K*Amin);  pole = a + Amin*bl,
//   in-cell offset e = ebase + a + A*m + K*A*bl.
// K == 0 instantiates the run-time-k fallback (k up to 10).
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
sweep_generic_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, double beta,
                     const GroupDev* __restrict__ groups, const TileDev* __restrict__ tiles,
                     Bcsr M, int krt, int p, int KDp, int A, int NPOLE, int Amin) {
    constexpr int KC = K ? K : 10;
    const int kk = K ? K : krt;
    const int NQ = 1 << p, NP = kk * NQ;
    extern __shared__ __align__(16) double smem[];
    __shared__ long long sbase[MAXL + 1];
    __shared__ int sS;

    const TileDev t = tiles[blockIdx.x];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nr = t.nr, PT = nr * NPOLE, TL = kk * NPOLE;
    double* xs = smem;
    double* ys = smem + (size_t)NP * PT;
    int* tab_s = reinterpret_cast<int*>(ys + (size_t)NP * PT);   // TL entries: m*PT + pole
    int* tab_g = tab_s + TL;                                      // TL entries: in-cell offset

    if (tid <= p) sbase[tid] = groups[t.group].base[tid];
    if (tid == 0) sS = groups[t.group].S;
    for (int tt = tid; tt < TL; tt += nth) {
        const int a = tt % Amin, rest = tt / Amin;
        const int m = rest % kk, bl = rest / kk;
        tab_s[tt] = m * PT + a + Amin * bl;
        tab_g[tt] = t.ebase + a + A * m + kk * A * bl;
    }
    __syncthreads();
    const int S = sS;
    const int warp = tid >> 5, lane = tid & 31, nwarp = nth >> 5;
    const int ncell = NQ * nr;

    for (int c = warp; c < ncell; c += nwarp) {
        const int q = c / nr, r = c - q * nr;
        const double* src = X + cell_addr(sbase, S, q, t.r0 + r, KDp);
        double* dst = xs + (size_t)q * kk * PT + r * NPOLE;
        for (int tt = lane; tt < TL; tt += 32) dst[tab_s[tt]] = src[tab_g[tt]];
    }
    __syncthreads();

    const int nunit = NQ * PT;
    for (int u = tid; u < nunit; u += nth) {
        const int q = u / PT, pt = u - q * PT;
        double acc[KC];
#pragma unroll
        for (int m = 0; m < KC; ++m) acc[m] = 0.0;
        const int b1 = M.rowptr[q + 1];
        for (int blk = M.rowptr[q]; blk < b1; ++blk) {
            const int qc = __ldg(M.col + blk);
            if (qc >= NQ) break;
            const double* xv = xs + (size_t)qc * kk * PT + pt;
            const double* hv = M.val + (size_t)blk * M.KK2;
            double xr[KC];
#pragma unroll
            for (int mi = 0; mi < KC; ++mi)
                if (mi < kk) xr[mi] = xv[mi * PT];
#pragma unroll
            for (int mo = 0; mo < KC; ++mo)
                if (mo < kk) {
#pragma unroll
                    for (int mi = 0; mi < KC; ++mi)
                        if (mi < kk) acc[mo] = fma(__ldg(hv + mo * kk + mi), xr[mi], acc[mo]);
                }
        }
        double* yv = ys + (size_t)q * kk * PT + pt;
#pragma unroll
        for (int mo = 0; mo < KC; ++mo)
            if (mo < kk) yv[mo * PT] = acc[mo];
    }
    __syncthreads();

    for (int c = warp; c < ncell; c += nwarp) {
        const int q = c / nr, r = c - q * nr;
        double* dstg = Y + cell_addr(sbase, S, q, t.r0 + r, KDp);
        const double* srcs = ys + (size_t)q * kk * PT + r * NPOLE;
        if (beta == 0.0) {
            for (int tt = lane; tt < TL; tt += 32) dstg[tab_g[tt]] = alpha * srcs[tab_s[tt]];
        } else {
            for (int tt = lane; tt < TL; tt += 32) atomicAdd(&dstg[tab_g[tt]], alpha * srcs[tab_s[tt]]);      // beta == 1
        }
    }
}

// ------------------------------------------------------------------------------------------
// Long poles: record stream helpers
// ------------------------------------------------------------------------------------------
constexpr int LONG_CH = 8;      // records per ring chunk
constexpr int LONG_NBUF = 4;    // ring depth

template <int K>
struct LongRec {
    static constexpr int KK = K * K;
    static constexpr int BYTES = (KK * 8 + 8 + 15) & ~15;     // values + {col, flags}, 16-byte multiple
};

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------
// Long poles, register-tiled (the shipped path for N' > 32): lanes = poles, C poles per lane.
// A CTA tile holds PT = 32*C consecutive poles of one pole group (flattened pole index
// item * PI + j, so a tile may straddle items) and one ROW PART of the principal sub-block; the
// whole x tile sits in shared memory as xs[row][PT] (conflict-free for lanes = poles).  Every
// warp owns a contiguous range of whole block-rows and streams its K x K block records through
// a private cp.async ring (compact record stream per class).  Per record a lane loads
// the K*K uniform H values once (broadcast LDS) and K*C of its own x values and issues K*K*C
// DFMAs: shared-memory wavefronts per DFMA fall from 2.9 (C = 1) to 1.2 (C = 4) -- shared-memory
// delivery, not the fp64 pipe, is what bounds this kernel.
// Global traffic needs no tables or scratch: with pole = c*32 + lane every warp-level access of a
// fixed (cell, mode, c) touches 32 consecutive poles, i.e. one contiguous (A > 1) or stride-K
// (A = 1, the other modes fill the gaps) run.  For y += alpha*D x the old y of the NEXT block-row
// is prefetched into registers before that row's records are processed, so the read-modify-write
// latency is hidden behind the row's arithmetic.
//   element (q, mo, pole): base[level(q)] + KS*(c(q) + C(q)*hi) + KDp*lo + a + K*A*b + A*mo,
//   pole -> item r = pole / PI (lo = r % S, hi = r / S), j = pole % PI (a = j % A, b = j / A).
// ------------------------------------------------------------------------------------------
struct TileL2 {
    int ctab;      // offset of the group's cell table (one CellOfs per 1-D cell q)
    int S;         // prod of cells of dims < d
    int item0;     // first item of the tile and ...
    int j0;        // ... first pole inside it (flattened pole = item * PI + j)
    int lo0, hi0;  // item0 = lo0 + S * hi0
    int npoles;    // <= 32*C
    int part;      // row part of the matrix this CTA computes
};

struct CellOfs {   // multi-cell of 1-D cell q for item (lo, hi): bq + KDp*lo + kc*hi
    long long bq;  // base[level(q)] + KDp*S*c(q)
    long long kc;  // KDp*S*C(level(q))
};

__device__ __forceinline__ void cp_async8_zfill(unsigned dst, const void* src, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

template <int K, int C>
struct LongOperands {          // one block record's operands, register-resident
    double h[K * K];
    double x[K][C];
    int flags;
};

// shared-memory loads with explicit 32-bit addresses: keeps the record loop's address arithmetic
// in registers (the compiler otherwise rematerialises the smem bases from SR_TID every record)
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double d;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(d) : "r"(addr));
    return d;
}
__device__ __forceinline__ void lds_v2f64(double& a, double& b, unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ int2 lds_v2s32(unsigned addr) {
    int2 r;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}

// NB = depth of the per-warp record ring: 4 (8 warps) or 2 (16 warps next to a 196 KB x tile)
template <int K, int C, int NB>
__global__ void __launch_bounds__(NB == 2 ? 512 : 256)
sweep_long2_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, int accumulate,
                   const CellOfs* __restrict__ celltab, const int* __restrict__ offtab,
                   const TileL2* __restrict__ tiles, const unsigned char* __restrict__ recs,
                   const int* __restrict__ partBlk, const int* __restrict__ partRow, int p,
                   int qc0, int nqc,                   // column pass: block columns [qc0, qc0 + nqc) of the matrix
                   int KDp, int A, int PI,
                   long long* __restrict__ dbg) {     // dbg: optional per-warp clock stamps (first 8 CTAs)
    constexpr int KK = K * K, REC = LongRec<K>::BYTES;
    constexpr int CHB = LONG_CH * REC;                        // bytes per ring chunk
    constexpr int RINGREC = LONG_CH * NB;              // records the ring holds
    constexpr int PT = 32 * C;
    const long long t_start = clock64();
    const long long g_t0 = dbg ? gtimer() : 0;
    const int NP = K * nqc;                                   // x rows staged by this pass
    extern __shared__ __align__(128) unsigned char smraw[];

    const TileL2 t = tiles[blockIdx.x];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int warp = tid >> 5, nwarp = nth >> 5;
    int lane = tid & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));
    double* xs = reinterpret_cast<double*>(smraw);                                  // NP * PT
    unsigned char* ring = smraw + (size_t)NP * PT * 8 + (size_t)warp * (NB * CHB);
    unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    unsigned xs_s = (unsigned)__cvta_generic_to_shared(xs) + lane * 8;
    // opaque copies: stops the compiler from rematerialising these bases (S2R + IMADs) per record
    asm volatile("mov.u32 %0, %0;" : "+r"(ring_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(xs_s));
    const CellOfs* ctab = celltab + t.ctab;

    // this warp's records [b0, b1) and first block-row q
    const int gpart = t.part * nwarp + warp;
    const int b0 = partBlk[gpart], b1 = partBlk[gpart + 1];
    int q = partRow[gpart];
    const int c_first = b0 / LONG_CH, c_last = b1 > b0 ? (b1 - 1) / LONG_CH : c_first - 1;

    auto issue_chunk = [&](int c) {
        if (c <= c_last) {
            const unsigned char* src = recs + (size_t)c * CHB;
            const unsigned dst = ring_s + (c % NB) * CHB;
            for (int g = lane; g < CHB / 16; g += 32) cp_async16(dst + g * 16, src + g * 16);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int j = 0; j < NB - 1; ++j) issue_chunk(c_first + j);

    // per-lane pole constants (no integer divisions: the tile carries its first item / pole)
    long long u[C], v[C];
    bool ok[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int pl = c * 32 + lane;
        ok[c] = pl < t.npoles;
        int jj = t.j0 + (ok[c] ? pl : 0), lo = t.lo0, hi = t.hi0;
        while (jj >= PI) {
            jj -= PI;
            if (++lo == t.S) { lo = 0; ++hi; }
        }
        u[c] = (long long)KDp * lo + offtab[jj];
        v[c] = hi;
    }

    // ---- stage the x tile in (asynchronous 8-byte copies, coalesced per (cell, mode, c))
    for (int qq = warp; qq < nqc; qq += nwarp) {
        const CellOfs co = ctab[qc0 + qq];
        const unsigned dst = xs_s + qq * (K * PT * 8);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double* src = X + co.bq + u[c] + co.kc * v[c];
#pragma unroll
            for (int mo = 0; mo < K; ++mo) cp_async8_zfill(dst + (mo * PT + c * 32) * 8, src + A * mo, ok[c]);
        }
    }
    cp_async_commit();
    const long long t_issued = clock64();
    cp_async_wait<0>();
    __syncthreads();
    const long long t_staged = clock64();
    if (b1 <= b0) return;

    // ---- stream this warp's block records, software-pipelined one record ahead (two register
    //      sets used alternately)
    double acc[K][C];
#pragma unroll
    for (int m = 0; m < K; ++m)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[m][c] = 0.0;

    long long rowofs[C];
    auto row_begin = [&](int qr) {      // element offsets of block-row qr
        const CellOfs co = ctab[qr];
#pragma unroll
        for (int c = 0; c < C; ++c) rowofs[c] = co.bq + u[c] + co.kc * v[c];
    };
    auto fetch = [&](int i, LongOperands<K, C>& o) {
        if ((i & (LONG_CH - 1)) == 0 || i == b0) {
            // entering chunk ci: every lane has finished reading chunk ci-1 (its last record is in
            // registers), so that buffer is refilled with chunk ci+NB-1; then chunk ci must have landed
            __syncwarp();
            issue_chunk(i / LONG_CH + NB - 1);
            cp_async_wait<NB - 1>();
            __syncwarp();
        }
        const unsigned ra = ring_s + (i & (RINGREC - 1)) * REC;
        if constexpr ((KK & 1) == 1) {
#pragma unroll
            for (int e = 0; e + 1 < KK; e += 2) lds_v2f64(o.h[e], o.h[e + 1], ra + e * 8);
            o.h[KK - 1] = lds_f64(ra + (KK - 1) * 8);
        } else {
#pragma unroll
            for (int e = 0; e < KK; e += 2) lds_v2f64(o.h[e], o.h[e + 1], ra + e * 8);
        }
        const int2 meta = lds_v2s32(ra + KK * 8);
        o.flags = meta.y;
        const unsigned xa = xs_s + meta.x * (K * PT * 8);
#pragma unroll
        for (int mi = 0; mi < K; ++mi)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) o.x[mi][cc] = lds_f64(xa + (mi * PT + cc * 32) * 8);
    };
    auto compute = [&](const LongOperands<K, C>& o) {
#pragma unroll
        for (int mi = 0; mi < K; ++mi)
#pragma unroll
            for (int mo = 0; mo < K; ++mo)
#pragma unroll
                for (int cc = 0; cc < C; ++cc) acc[mo][cc] = fma(o.h[mo * K + mi], o.x[mi][cc], acc[mo][cc]);
    };
    auto row_end = [&](bool more) {
#pragma unroll
        for (int cc = 0; cc < C; ++cc)
#pragma unroll
            for (int mo = 0; mo < K; ++mo) {
                if (ok[cc]) {
                    // y += ... as a reduction at the L2 (RED.ADD.F64): no read-modify-write through the SM, and sweeps of
                    // different directions may accumulate into y concurrently
                    if (accumulate) atomicAdd(&Y[rowofs[cc] + A * mo], alpha * acc[mo][cc]);
                    else Y[rowofs[cc] + A * mo] = alpha * acc[mo][cc];
                }
                acc[mo][cc] = 0.0;
            }
        ++q;
        if (more) row_begin(q);
    };

    row_begin(q);
    LongOperands<K, C> ra_, rb_;
    fetch(b0, ra_);
    for (int i = b0;;) {
        if (i + 1 < b1) fetch(i + 1, rb_);
        compute(ra_);
        if (ra_.flags & 1) row_end(i + 1 < b1);
        if (++i >= b1) break;
        if (i + 1 < b1) fetch(i + 1, ra_);
        compute(rb_);
        if (rb_.flags & 1) row_end(i + 1 < b1);
        if (++i >= b1) break;
    }
    cp_async_wait<0>();
    if (dbg && blockIdx.x < 8 && lane == 0 && warp < 8) {
        long long* o = dbg + (blockIdx.x * 8 + warp) * 8;
        o[0] = t_start; o[1] = t_issued; o[2] = t_staged; o[3] = clock64(); o[4] = b1 - b0; o[5] = q - partRow[gpart];
    }
    if (dbg && tid == 0) {          // CTA placement log: {smid, start, end, class tag}
        const int slot = atomicAdd(reinterpret_cast<int*>(dbg + 1023), 1);
        if (slot < 500) {
            long long* o = dbg + 6144 + slot * 4;
            o[0] = smid_u32(); o[1] = g_t0; o[2] = gtimer(); o[3] = p * 100 + qc0 / (nqc > 0 ? nqc : 1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Medium poles (N' = 48, 96 at k = 3), matrix in the constant bank.  The block pattern of the 1-D
// operator is structural -- block (q, r) is stored iff the closed supports of the two hierarchical
// cells intersect or touch periodically (checked against the host's H at
// plan creation) -- so the whole principal sub-block is unrolled at compile time and every DFMA
// takes its H operand as an immediate constant-bank address of a __grid_constant__ parameter:
// the matrix costs no shared-memory bandwidth at all (the register-tiled kernel above is bound by
// exactly that).  One warp = 32 consecutive poles (lanes = poles); warps are independent (no CTA
// barrier): each stages its own x tile xs[row][32] with asynchronous copies, then walks the rows
// with 3 conflict-free LDS.64 + 9 DFMA per block; old y is prefetched two block-rows ahead.
// ------------------------------------------------------------------------------------------
namespace pat {
constexpr int FULL = 1 << 16;
__host__ __device__ constexpr int cell_lo(int q) {
    if (q == 0) return 0;
    int l = 0;
    for (int t = q; t; t >>= 1) ++l;
    return (q - (1 << (l - 1))) * (FULL >> (l - 1));
}
__host__ __device__ constexpr int cell_hi(int q) {
    if (q == 0) return FULL;
    int l = 0;
    for (int t = q; t; t >>= 1) ++l;
    return (q - (1 << (l - 1)) + 1) * (FULL >> (l - 1));
}
__host__ __device__ constexpr bool touch(int q, int r) {
    const int a1 = cell_lo(q), b1 = cell_hi(q), a2 = cell_lo(r), b2 = cell_hi(r);
    return (a1 <= b2 && a2 <= b1) || (b1 == FULL && a2 == 0) || (b2 == FULL && a1 == 0);
}
// number of pattern blocks before (q, r) in row-major order, rows/cols < nq
__host__ __device__ constexpr int block_index(int nq, int q, int r) {
    int n = 0;
    for (int i = 0; i < q; ++i)
        for (int j = 0; j < nq; ++j) n += touch(i, j) ? 1 : 0;
    for (int j = 0; j < r; ++j) n += touch(q, j) ? 1 : 0;
    return n;
}
__host__ __device__ constexpr int nblocks(int nq) { return block_index(nq, nq, 0); }
}  // namespace pat

template <int K, int P>
struct HBlocks {
    double v[pat::nblocks(1 << P) * K * K];
};

template <int K, int P, int Q, int R>
__device__ __forceinline__ void consth_block(const HBlocks<K, P>& hb, unsigned xs_s, double (&acc)[2][K]) {
    if constexpr (pat::touch(Q, R)) {
        constexpr int BI = pat::block_index(1 << P, Q, R);
        constexpr int OFF = BI * K * K;
        constexpr int SET = BI & 1;
        double x[K];
#pragma unroll
        for (int mi = 0; mi < K; ++mi) x[mi] = lds_f64(xs_s + (R * K + mi) * 256);
#pragma unroll
        for (int mi = 0; mi < K; ++mi)
#pragma unroll
            for (int mo = 0; mo < K; ++mo) acc[SET][mo] = fma(hb.v[OFF + mo * K + mi], x[mi], acc[SET][mo]);
    }
}

template <int K, int P, int Q, int... Rs>
__device__ __forceinline__ void consth_row(const HBlocks<K, P>& hb, unsigned xs_s, double (&acc)[2][K],
                                           std::integer_sequence<int, Rs...>) {
    (consth_block<K, P, Q, Rs>(hb, xs_s, acc), ...);
}

template <int K, int P>
struct ConstHCtx {
    const double* X;
    double* Y;
    const CellOfs* ctab_s;     // the group's cell table, in shared memory
    long long u, v;
    double alpha;
    int A;
    bool ok, accumulate;
    unsigned xs_s;
};

template <int K, int P, int Q>
__device__ __forceinline__ void consth_rows(const HBlocks<K, P>& hb, const ConstHCtx<K, P>& cx,
                                            long long (&rowofs)[3]) {
    constexpr int NQ = 1 << P;
    if constexpr (Q < NQ) {
        if constexpr (Q + 2 < NQ) {          // row offsets two block-rows ahead (slots rotate modulo 3)
            const CellOfs co = cx.ctab_s[Q + 2];
            rowofs[(Q + 2) % 3] = co.bq + cx.u + co.kc * cx.v;
        }
        double acc[2][K];
#pragma unroll
        for (int mo = 0; mo < K; ++mo) acc[0][mo] = acc[1][mo] = 0.0;
        consth_row<K, P, Q>(hb, cx.xs_s, acc, std::make_integer_sequence<int, NQ>{});
        if (cx.ok) {
#pragma unroll
            for (int mo = 0; mo < K; ++mo) {
                const double r = acc[0][mo] + acc[1][mo];
                if (cx.accumulate) atomicAdd(&cx.Y[rowofs[Q % 3] + cx.A * mo], cx.alpha * r);      // RED.ADD.F64
                else cx.Y[rowofs[Q % 3] + cx.A * mo] = cx.alpha * r;
            }
        }
        consth_rows<K, P, Q + 1>(hb, cx, rowofs);
    }
}

constexpr int CONSTH_WARPS = 4;

template <int K, int P>
__global__ void __launch_bounds__(32 * CONSTH_WARPS)
sweep_consth_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, int accumulate,
                    const CellOfs* __restrict__ celltab, const int* __restrict__ offtab,
                    const TileL2* __restrict__ tiles, int ntiles, int KDp, int A, int PI,
                    const __grid_constant__ HBlocks<K, P> hb, long long* __restrict__ dbg) {
    constexpr int NQ = 1 << P, NP = K * NQ;
    const long long g_t0 = dbg ? gtimer() : 0;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int warp = threadIdx.x >> 5;
    int lane = threadIdx.x & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));
    // warps are independent (no CTA-wide barrier below), one tile each.  (A persistent variant -- grid
    // stride loop, <= 128 registers -- does share SMs with the streaming kernel, but costs 40 more
    // registers or spills and was slower end to end)
    const int ti = blockIdx.x * CONSTH_WARPS + warp;
    if (ti >= ntiles) return;
    const TileL2 t = tiles[ti];
    unsigned char* wbase = smraw + (size_t)warp * (NP * 32 * 8 + NQ * sizeof(CellOfs));
    CellOfs* ctab_s = reinterpret_cast<CellOfs*>(wbase + NP * 32 * 8);
    unsigned xs_s = (unsigned)__cvta_generic_to_shared(wbase) + lane * 8;
    asm volatile("mov.u32 %0, %0;" : "+r"(xs_s));

    // per-lane pole constants
    const bool ok = lane < t.npoles;
    long long u, v;
    {
        int jj = t.j0 + (ok ? lane : 0), lo = t.lo0, hi = t.hi0;
        while (jj >= PI) {
            jj -= PI;
            if (++lo == t.S) { lo = 0; ++hi; }
        }
        u = (long long)KDp * lo + offtab[jj];
        v = hi;
    }
    // stage the cell table and the x tile
    for (int qq = lane; qq < NQ; qq += 32) ctab_s[qq] = celltab[t.ctab + qq];
    __syncwarp();
#pragma unroll 4
    for (int qq = 0; qq < NQ; ++qq) {
        const CellOfs co = ctab_s[qq];
        const double* src = X + co.bq + u + co.kc * v;
#pragma unroll
        for (int mo = 0; mo < K; ++mo) cp_async8_zfill(xs_s + (qq * K + mo) * 256, src + A * mo, ok);
    }
    cp_async_commit();

    ConstHCtx<K, P> cx;
    cx.X = X; cx.Y = Y; cx.ctab_s = ctab_s; cx.u = u; cx.v = v; cx.alpha = alpha; cx.A = A;
    cx.ok = ok; cx.accumulate = accumulate != 0; cx.xs_s = xs_s;
    long long rowofs[3];
#pragma unroll
    for (int r = 0; r < 2 && r < NQ; ++r) {          // rows 0 and 1 here, the rest two rows ahead
        const CellOfs co = ctab_s[r];
        rowofs[r] = co.bq + u + co.kc * v;
    }
    cp_async_wait<0>();
    __syncwarp();
    consth_rows<K, P, 0>(hb, cx, rowofs);
    if (dbg && threadIdx.x == 0 && blockIdx.x < 512) {
        long long* o = dbg + 1024 + blockIdx.x * 4;
        o[0] = smid_u32(); o[1] = g_t0; o[2] = gtimer();
    }
}

// ------------------------------------------------------------------------------------------
// Long poles, ROW-TILE kernel (N' >= 48 when an item has >= 64 poles; rowtile.inl builds the tile programs).
// Unit of data movement = the whole multi-cell (all k^(D-1) poles of an item for one 1-D cell: ONE contiguous,
// 16-byte aligned run): the x tile is staged with a handful of TMA bulk copies (UBLKCP), lanes = poles (C per
// lane, in a bank-permuted order: the 16 lanes of a half-warp read 16 different 8-byte banks), and a
// CTA = (item, tile of the class's program): the tile's x cells and block records arrive on one mbarrier.
// Warp (row group, pole warp) walks its ROW GROUPS: up to RT_R output rows that share most of their columns are
// accumulated together, so an x value read from shared memory feeds up to RT_R * K DFMAs and a broadcast H value
// feeds C: the loop is bound by the fp64 pipe, not by shared-memory wavefronts.  A record = {x cell, row mask}
// followed by one K x K block per row of the mask (uniform branches); the blocks of rows (0, 1) and (2, 3) are
// multiplied interleaved, which doubles the independent DFMA chains per thread (K * C -> 2 K C: with two warps per
// scheduler the dependent-issue latency of DFMA is what bounds the loop).  Finished rows go through a per-row-group
// staging cell in shared memory and leave as ONE bulk store (complete rows, beta = 0) or bulk reduce-add
// (UBLKRED.ADD.F64: accumulating sweeps and the partial sums of rows above a subtree tile -- those rows are zeroed
// beforehand when beta = 0), so y is written in whole 16-byte aligned multi-cells whatever the lane order.
// ------------------------------------------------------------------------------------------
constexpr int RT_MAXX = 40;        // x cells per tile
constexpr int RT_MAXRG = 4;        // row groups (warps along the rows) per CTA
constexpr int RT_R = 4;            // output rows accumulated together

struct RTTile {
    int nx;                // x cells of the tile
    int rec_ofs;           // byte offset of its record blob in the program's blob (multiple of 16)
    int rec_bytes;         // bytes of the blob (multiple of 16)
    int grp0;              // first entry in the program's group array
    int rg_end[RT_MAXRG];  // groups [rg_end[g-1], rg_end[g]) (relative to grp0) belong to row group g
    int xq[RT_MAXX];       // 1-D cell of x slot i
};

struct RTGroup {
    int q[RT_R];           // output 1-D cells (-1: unused)
    int rofs;              // byte offset of the group's first record inside the tile's blob
    int nrec;              // records (= distinct x cells of the group)
    int partial;           // bit r: row r is a partial sum (always reduced into y)
    int pad;
};

struct RTWork {
    int ctab;              // offset of the group's cell table
    int lo, hi;            // item = lo + S * hi
    int tile;
    int nx, rec_ofs, rec_bytes, pad;   // copies of the tile's fields (one dependent load less before the first bulk copy)
};

constexpr int RT_NBAR = 4;         // the x cells of a tile arrive on RT_NBAR barriers (records with the first)

__global__ void zero_cells_kernel(double* __restrict__ Y, const int* __restrict__ cells, int ncells, int KDp) {
    for (int c = blockIdx.x; c < ncells; c += gridDim.x) {
        double2* dst = reinterpret_cast<double2*>(Y + (size_t)cells[c] * KDp);
        for (int e = threadIdx.x; e < KDp / 2; e += blockDim.x) dst[e] = make_double2(0.0, 0.0);
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// shared memory: [RT_NBAR full barriers + 1 "x free" barrier (64 B)] [2 * RG staging cells] [nx x cells] [block entries]
// PERSISTENT: CTA b takes the work items sched[b], sched[b + gridDim.x], ... until -1 (the host deals the items to
// the CTAs, most expensive first onto the least loaded CTA).  Warp 0 doubles as
// the producer: once every warp has left the main loops of a tile (the x cells and entries are dead -- the row
// epilogues only touch registers and the staging cells), it issues the bulk copies of the CTA's next tile, so that
// tile's load latency overlaps the epilogues of this one and no CTA launch sits between two tiles.
template <int K, int C, int NT>
__global__ void __launch_bounds__(NT, 1)
sweep_rowtile_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, int accumulate,
                     const CellOfs* __restrict__ celltab, const int* __restrict__ offtab,
                     const RTWork* __restrict__ work, const int* __restrict__ sched, const long long* __restrict__ xsrc,
                     const RTTile* __restrict__ tiles, const RTGroup* __restrict__ groups,
                     const unsigned char* __restrict__ recs, int KDp, int A, int PW, int RG,
                     long long* __restrict__ dbg) {      // dbg: optional per-tile time stamps (tools/stamps_rowtile.py)
    constexpr int KK = K * K;
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smraw);     // [0, RT_NBAR): full; [RT_NBAR]: x free
    const unsigned cell_bytes = (unsigned)KDp * 8u;
    double* stage = reinterpret_cast<double*>(smraw + 64);                       // 2 * RG staging cells
    double* xs = reinterpret_cast<double*>(smraw + 64 + (size_t)2 * RG * cell_bytes);
    const unsigned xs_s = tma::smem_u32(xs);

    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));
    const int nwarps = PW * RG;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < RT_NBAR; ++b) tma::mbar_init(tma::smem_u32(bar + b), 1);
        tma::mbar_init(tma::smem_u32(bar + RT_NBAR), (unsigned)nwarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * RG) stage[(size_t)tid * KDp + KDp - 1] = 0.0;        // the padding slot of a multi-cell stays zero
    __syncthreads();

    // producer (whole warp 0): the bulk copies of work item wi; cell i arrives on barrier (i * RT_NBAR) / nx, the
    // entries on barrier 0.  The only loads in front of the copies: the work item and its cells' source offsets.
    auto issue_tile = [&](int wi) {
        const int4 v1 = *(reinterpret_cast<const int4*>(work + wi) + 1);          // nx, rec_ofs, rec_bytes
        const int nx = v1.x;
        if (lane < RT_NBAR) {
            const int c0 = (lane * nx + RT_NBAR - 1) / RT_NBAR, c1 = ((lane + 1) * nx + RT_NBAR - 1) / RT_NBAR;
            tma::mbar_expect_tx(tma::smem_u32(bar + lane), (unsigned)(c1 - c0) * cell_bytes + (lane == 0 ? (unsigned)v1.z : 0u));
        }
        __syncwarp();
        if (lane == 0) tma::bulk_g2s(xs_s + (unsigned)nx * cell_bytes, recs + v1.y, (unsigned)v1.z, tma::smem_u32(bar));
        for (int i = lane; i < nx; i += 32)
            tma::bulk_g2s(xs_s + (unsigned)i * cell_bytes, X + xsrc[(size_t)wi * RT_MAXX + i], cell_bytes,
                          tma::smem_u32(bar + (i * RT_NBAR) / nx));
    };
    int wi = sched[blockIdx.x];
    if (warp == 0 && wi >= 0) issue_tile(wi);

    // per-lane pole constants: pole slot c of this lane = entry (pw * C + c) * 32 + lane of the bank-permuted table;
    // a negative entry ~o marks a padding lane (it reads pole offset o like a neighbour and stores nothing)
    const int pw = warp % PW, rg = warp / PW;
    unsigned pob[C];
    bool ok[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int v = offtab[(pw * C + c) * 32 + lane];
        ok[c] = v >= 0;
        pob[c] = (unsigned)(v >= 0 ? v : ~v) * 8u;
    }
    const unsigned A8 = (unsigned)A * 8u;
    const unsigned st_s = tma::smem_u32(stage) + (unsigned)(2 * rg) * cell_bytes;
    const bool issuer = pw == 0 && lane == 0;
    const int nbar = 32 * PW;
    unsigned sbuf = 0;                      // staging cell of the next row (alternates)

    for (int it = 0; wi >= 0; ++it) {
        const int wcur = wi;
        wi = sched[(size_t)(it + 1) * gridDim.x + blockIdx.x];        // -1 after the CTA's last item
        const int wnext = wi;
        const unsigned par = (unsigned)it & 1u;
        const long long g_t0 = dbg ? gtimer() : 0;
        long long t_main = 0, t_epi = 0;
        const int4 w0 = *reinterpret_cast<const int4*>(work + wcur);              // ctab, lo, hi, tile
        const int4 w1 = *(reinterpret_cast<const int4*>(work + wcur) + 1);        // nx, rec_ofs, rec_bytes
        const unsigned rs_s = xs_s + (unsigned)w1.x * cell_bytes;
        const RTTile* __restrict__ T = tiles + w0.w;
        const CellOfs* __restrict__ ctab = celltab + w0.x;
        const long long item_ofs = (long long)KDp * w0.y;
        const int g_begin = T->grp0 + (rg == 0 ? 0 : T->rg_end[rg - 1]), g_end = T->grp0 + T->rg_end[rg];
        // a warp leaves the tile's x cells and entries behind; warp 0 then fetches the CTA's next tile
        auto done_reading = [&]() {
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(tma::smem_u32(bar + RT_NBAR));
            if (warp == 0 && wnext >= 0) {
                tma::mbar_wait(tma::smem_u32(bar + RT_NBAR), par);
                issue_tile(wnext);
            }
        };
        int4 gq = make_int4(-1, -1, -1, -1), gm = make_int4(0, 0, 0, 0);
        if (g_begin < g_end) {
            gq = *reinterpret_cast<const int4*>(groups + g_begin);
            gm = *(reinterpret_cast<const int4*>(groups + g_begin) + 1);          // rofs, nrec, partial, pad
        }
        // EVERY warp waits for the tile's first barrier before it may signal "done reading": a warp without rows in
        // this tile must not run ahead and arrive twice in one phase of the "x free" barrier
        tma::mbar_wait(tma::smem_u32(bar), par);
        if (g_begin >= g_end) {             // no rows for this row group in this tile
            done_reading();
            continue;
        }
        const long long g_t1 = dbg ? gtimer() : 0;
        unsigned arrived = 1u;              // barriers this warp has seen complete in this tile

        for (int gi = g_begin; gi < g_end; ++gi) {
            const int4 cq = gq, cm = gm;
            if (gi + 1 < g_end) {           // next group's descriptor while this one is computed
                gq = *reinterpret_cast<const int4*>(groups + gi + 1);
                gm = *(reinterpret_cast<const int4*>(groups + gi + 1) + 1);
            }
            const int qs[RT_R] = {cq.x, cq.y, cq.z, cq.w};
            long long yofs[RT_R];           // output cells, fetched now and used in the epilogue
#pragma unroll
            for (int r = 0; r < RT_R; ++r) {
                yofs[r] = 0;
                if (issuer && qs[r] >= 0) {
                    const CellOfs co = ctab[qs[r]];
                    yofs[r] = co.bq + item_ofs + co.kc * w0.z;
                }
            }
            double acc[RT_R][K][C];
#pragma unroll
            for (int r = 0; r < RT_R; ++r)
#pragma unroll
                for (int m = 0; m < K; ++m)
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[r][m][c] = 0.0;
            const long long c_0 = dbg ? clock64() : 0;
            unsigned ptr = rs_s + (unsigned)cm.x;
            int2 hd = lds_v2s32(ptr);                // x cell offset (bytes); row mask | barrier of the cell << 8
            for (int i = 0; i < cm.y; ++i) {
                const int2 hc = hd;
                const unsigned bi = ((unsigned)hc.y >> 8) & (RT_NBAR - 1);
                if (!((arrived >> bi) & 1u)) {       // uniform over the warp
                    tma::mbar_wait(tma::smem_u32(bar + bi), par);
                    arrived |= 1u << bi;
                }
                const unsigned xa = xs_s + (unsigned)hc.x;
                double xv[K][C];
#pragma unroll
                for (int mi = 0; mi < K; ++mi)
#pragma unroll
                    for (int c = 0; c < C; ++c) xv[mi][c] = lds_f64(xa + pob[c] + A8 * mi);
                unsigned bp = ptr + 8;
                ptr += 8 + KK * 8 * __popc(hc.y & ((1 << RT_R) - 1));
                if (i + 1 < cm.y) hd = lds_v2s32(ptr);         // the next record's header while this one is multiplied
#pragma unroll
                for (int half = 0; half < RT_R / 2; ++half) {
                    const int m2 = (hc.y >> (2 * half)) & 3; // uniform over the CTA's lanes
                    if (m2 == 3) {                   // both rows of the pair: 2 K C independent DFMA chains
                        double ha[KK], hb[KK];
#pragma unroll
                        for (int e = 0; e < KK; ++e) ha[e] = lds_f64(bp + e * 8);
#pragma unroll
                        for (int e = 0; e < KK; ++e) hb[e] = lds_f64(bp + (KK + e) * 8);
                        bp += 2 * KK * 8;
#pragma unroll
                        for (int mi = 0; mi < K; ++mi) {
#pragma unroll
                            for (int mo = 0; mo < K; ++mo)
#pragma unroll
                                for (int c = 0; c < C; ++c)
                                    acc[2 * half][mo][c] = fma(ha[mo * K + mi], xv[mi][c], acc[2 * half][mo][c]);
#pragma unroll
                            for (int mo = 0; mo < K; ++mo)
#pragma unroll
                                for (int c = 0; c < C; ++c)
                                    acc[2 * half + 1][mo][c] = fma(hb[mo * K + mi], xv[mi][c], acc[2 * half + 1][mo][c]);
                        }
                    } else if (m2 != 0) {
                        double ha[KK];
#pragma unroll
                        for (int e = 0; e < KK; ++e) ha[e] = lds_f64(bp + e * 8);
                        bp += KK * 8;
                        if (m2 == 1) {
#pragma unroll
                            for (int mi = 0; mi < K; ++mi)
#pragma unroll
                                for (int mo = 0; mo < K; ++mo)
#pragma unroll
                                    for (int c = 0; c < C; ++c)
                                        acc[2 * half][mo][c] = fma(ha[mo * K + mi], xv[mi][c], acc[2 * half][mo][c]);
                        } else {
#pragma unroll
                            for (int mi = 0; mi < K; ++mi)
#pragma unroll
                                for (int mo = 0; mo < K; ++mo)
#pragma unroll
                                    for (int c = 0; c < C; ++c)
                                        acc[2 * half + 1][mo][c] = fma(ha[mo * K + mi], xv[mi][c], acc[2 * half + 1][mo][c]);
                        }
                    }
                }
            }
            if (gi + 1 == g_end) done_reading();
            const long long c_1 = dbg ? clock64() : 0;
#pragma unroll
            for (int r = 0; r < RT_R; ++r) {
                if (qs[r] < 0) continue;
                // two staging cells alternate; ONE barrier per row tells the row group's warps both "this row is staged"
                // and "the other cell is free again" (the issuer first waits for the previous row's bulk read)
                const unsigned sb = st_s + sbuf * cell_bytes;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int mo = 0; mo < K; ++mo)
                        if (ok[c]) sts_f64(sb + pob[c] + A8 * mo, alpha * acc[r][mo][c]);
                tma::fence_proxy_async();
                if (issuer) tma::bulk_wait_read0();
                named_bar_sync(1 + rg, nbar);
                if (issuer) {
                    double* yrow = Y + yofs[r];
                    if (accumulate != 0 || ((cm.z >> r) & 1)) tma::bulk_red_add_f64(yrow, sb, cell_bytes);
                    else tma::bulk_s2g(yrow, sb, cell_bytes);
                    tma::bulk_commit();
                }
                sbuf ^= 1u;
            }
            if (dbg) { t_main += c_1 - c_0; t_epi += clock64() - c_1; }
        }
        if (dbg && lane == 0 && wcur < 1000) {       // per tile: start, data ready, end (ns); smid; per-warp clocks summed
            long long* o = dbg + (size_t)wcur * 8;
            if (warp == 0) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                o[0] = g_t0; o[1] = g_t1; o[4] = smid; o[7] = w0.w;
            }
            atomicMax(reinterpret_cast<unsigned long long*>(o + 2), (unsigned long long)gtimer());
            atomicAdd(reinterpret_cast<unsigned long long*>(o + 5), (unsigned long long)t_main);
            atomicAdd(reinterpret_cast<unsigned long long*>(o + 6), (unsigned long long)t_epi);
            atomicMax(reinterpret_cast<unsigned long long*>(o + 3), (unsigned long long)t_main);
        }
    }
    if (issuer) tma::bulk_wait_read0();                  // shared memory must outlive the last bulk read
}

// development aid: a spinner with a chosen resource footprint, to probe which kernels share an SM
__global__ void debug_spin_kernel(long long* __restrict__ dbg, int slot, int ns) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const long long t0 = gtimer();
    double a = threadIdx.x;
    while (gtimer() - t0 < ns) a = a * 1.0000001 + 1e-9;
    smraw[threadIdx.x] = (unsigned char)a;
    if (threadIdx.x == 0 && blockIdx.x < 512) {
        long long* o = dbg + slot + blockIdx.x * 4;
        o[0] = smid_u32(); o[1] = t0; o[2] = gtimer(); o[3] = (long long)a;
    }
}

// ------------------------------------------------------------------------------------------
// RK stage updates (classical RK4)
//   w   = u + cw * k
//   acc = (first ? u : acc) + ca * k
// and the final  u = acc + ca * k.
// ------------------------------------------------------------------------------------------
__global__ void rk_stage_kernel(long long N, const double* __restrict__ u, const double* __restrict__ k,
                                double* __restrict__ acc, double* __restrict__ w, double cw, double ca,
                                int first) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const double ui = u[i], ki = k[i];
        const double ai = first ? ui : acc[i];
        w[i] = fma(cw, ki, ui);
        acc[i] = fma(ca, ki, ai);
    }
}

__global__ void rk_final_kernel(long long N, double* __restrict__ u, const double* __restrict__ k,
                                const double* __restrict__ acc, double ca) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
        u[i] = fma(ca, k[i], acc[i]);
}

// u += c1 v1 + c2 v2 + c3 v3 + c4 v4   (Taylor form of RK4 for linear right-hand sides; summed
// smallest term first)
__global__ void rk4_taylor_kernel(long long N, double* __restrict__ u, const double* __restrict__ v1,
                                  const double* __restrict__ v2, const double* __restrict__ v3,
                                  const double* __restrict__ v4, double c1, double c2, double c3, double c4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        double s = c4 * v4[i];
        s = fma(c3, v3[i], s);
        s = fma(c2, v2[i], s);
        s = fma(c1, v1[i], s);
        u[i] += s;
    }
}

__global__ void rk4_taylor_cells_kernel(const int* __restrict__ cells, long long ncells, int KDp, double* __restrict__ u,
                                        const double* __restrict__ v1, const double* __restrict__ v2,
                                        const double* __restrict__ v3, const double* __restrict__ v4, double c1,
                                        double c2, double c3, double c4) {
    for (long long ci = blockIdx.x; ci < ncells; ci += gridDim.x) {
        const long long base = (long long)cells[ci] * KDp;
        for (int e = threadIdx.x; e < KDp; e += blockDim.x) {
            const long long i = base + e;
            double s = c4 * v4[i];
            s = fma(c3, v3[i], s);
            s = fma(c2, v2[i], s);
            s = fma(c1, v1[i], s);
            u[i] += s;
        }
    }
}

__global__ void scale_kernel(long long N, double* __restrict__ y, double beta) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) y[i] *= beta;
}

__global__ void sumsq_kernel(long long N, const double* __restrict__ x, double* __restrict__ out) {
    double s = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
        s = fma(x[i], x[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) ws[warp] = s;
    __syncthreads();
    if (warp == 0) {
        s = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) atomicAdd(out, s);
    }
}

// ------------------------------------------------------------------------------------------
// tensor_construct on the device (src/tensor_construct.jl:19-63, vector overload): the coefficient of
// prod_d f_d(x_d) on the sparse index set is the product of the 1-D coefficients, `val = one(T); for d in 1:D
// val *= coeff[d][l_d][c_d][m_d]` (same multiplication order: bit-identical to the host version).
// One CTA per multi-cell (grid-stride); v1d = D concatenated 1-D vectors of length K * 2^n in vector layout.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tensor_construct_kernel(const int* __restrict__ cell_block, long long ncells, const unsigned char* __restrict__ blk_level,
                        const long long* __restrict__ blk_poffset, const double* __restrict__ v1d, int D, int k, int n1d,
                        int KD, int KDp, double* __restrict__ out) {
    __shared__ int base1d[16];          // k * (first cell of level l_d + c_d): index of mode 0 in the 1-D vector
    for (long long ci = blockIdx.x; ci < ncells; ci += gridDim.x) {
        const int b = cell_block[ci];
        const long long poff = blk_poffset[b];
        __syncthreads();
        if (threadIdx.x == 0) {
            long long rem = ci - poff / KDp;      // cell index inside the block, first dimension fastest
            for (int d = 0; d < D; ++d) {
                const int l = blk_level[(size_t)b * D + d];
                const int cells = l <= 1 ? 1 : 1 << (l - 1);
                const int c = (int)(rem % cells);
                rem /= cells;
                base1d[d] = k * ((l == 0 ? 0 : 1 << (l - 1)) + c);
            }
        }
        __syncthreads();
        double* dst = out + ci * KDp;
        for (int e = threadIdx.x; e < KD; e += blockDim.x) {
            int r = e;
            double val = 1.0;
            for (int d = 0; d < D; ++d) {
                const int m = r % k;
                r /= k;
                val = __dmul_rn(val, v1d[(size_t)d * n1d + base1d[d] + m]);
            }
            dst[e] = val;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Batched reconstruct_DG (src/dg_methods.jl:150-165): one warp per point.
//   1. the warp fills a per-point table of the D*(n+1)*K one-dimensional basis values
//      v(k, l, cell(x_i,l), m, x_i) and the cell indices (src/dg_methods.jl:27-36,70-79);
//   2. lanes stride the K^D coefficients of the point's cell in every multi-level block,
//      multiply by the product of the D table entries (accumulated from 1.0 over i = 1..D as
//      `V` does, src/dg_methods.jl:47-54) and accumulate; one warp reduction per point.
// Basis tables: leg[(K_MAX+1) x 2(K_MAX+1)], dg[k x 2k] exactly as the reference stores them.
// ------------------------------------------------------------------------------------------
struct ReconTables {
    const unsigned char* blk_level;   // nblocks * D  (0-based levels)
    const long long* blk_offset;      // nblocks
    const double* leg;                // (KMAX+1) * 2(KMAX+1)
    const double* dg;                 // k * 2k
    int nblocks, D, k, n, KD, KDp, legw;   // legw = 2*(KMAX+1); KDp = padded cell stride
};

__device__ __forceinline__ double poly_eval(const double* __restrict__ v, int half, double x) {
    // array2poly, src/1d_dg_functions.jl:15-28
    if (fabs(x) > 1.0) return 0.0;
    const bool neg = signbit(x);
    double s = 0.0;
    for (int i = half - 1; i >= 0; --i) {
        const double t = neg ? -v[i + half] : v[i + half];
        s = __dadd_rn(__dmul_rn(s, x), v[i] + t);
    }
    return s;
}

__global__ void __launch_bounds__(256)
reconstruct_kernel(ReconTables T, const double* __restrict__ coeffs, const double* __restrict__ pts,
                   long long npts, double* __restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int D = T.D, k = T.k, n1 = T.n + 1;
    const int ntab = D * n1 * k;
    double* bt = smem + (size_t)warp * (ntab + D * n1);           // basis values [d][l][m]
    int* ci = reinterpret_cast<int*>(bt + ntab);                  // cell index [d][l] (0-based)
    const double sqrt2 = sqrt(2.0);

    for (long long pt = (long long)blockIdx.x * nwarp + warp; pt < npts; pt += (long long)gridDim.x * nwarp) {
        __syncwarp();
        {   // a point outside [0, 1]^D (or NaN) is a BoundsError in the reference: flag it, never dereference
            bool bad = false;
            for (int d = lane; d < D; d += 32) {
                const double xd = pts[pt * D + d];
                bad = bad || !(xd >= 0.0 && xd <= 1.0);
            }
            if (__any_sync(0xffffffffu, bad)) {
                if (lane == 0) out[pt] = __longlong_as_double(0x7ff8000000000000LL);
                continue;
            }
        }
        for (int idx = lane; idx < ntab; idx += 32) {
            const int m = idx % k, dl = idx / k, l = dl % n1, d = dl / n1;
            const double x = pts[pt * D + d];
            long long cell;   // 1-based, src/dg_methods.jl:70-79
            if (l <= 1) cell = 1;
            else if (x >= 1.0) cell = 1LL << (l - 1);
            else cell = 1 + (long long)floor((double)(1LL << (l - 1)) * x);
            double val;
            if (l == 0) {
                val = poly_eval(T.leg + m * T.legw, T.legw / 2, 2.0 * x - 1.0) * sqrt2;
            } else {
                const double sc = (double)(1LL << l);
                val = poly_eval(T.dg + m * 2 * k, k, sc * x - (double)(2 * cell - 1)) * sqrt(sc);
            }
            bt[idx] = val;
            if (m == 0) ci[dl] = (int)(cell - 1);
        }
        __syncwarp();
        double acc = 0.0;
        for (int b = 0; b < T.nblocks; ++b) {
            const unsigned char* lv = T.blk_level + (size_t)b * D;
            long long lin = 0, stride = 1;
            for (int d = 0; d < D; ++d) {
                const int l = lv[d];
                lin += (long long)ci[d * n1 + l] * stride;
                stride *= (l <= 1) ? 1 : (1LL << (l - 1));
            }
            const double* cf = coeffs + T.blk_offset[b] + lin * T.KDp;
            for (int e = lane; e < T.KD; e += 32) {
                int rem = e;
                double prod = 1.0;
                for (int d = 0; d < D; ++d) {
                    const int m = rem % k;
                    rem /= k;
                    prod *= bt[(d * n1 + lv[d]) * k + m];
                }
                acc = fma(cf[e], prod, acc);
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[pt] = acc;
    }
}

// Second-generation reconstruct kernel (the shipped path): still one warp per point, but the per-mode
// product  prod_i v_i[m_i]  is factored into a LOW half (dims < nlow) and a HIGH half, each tabulated once per
// (point, multi-level) by the warp (KL + KH products instead of D multiplications, divisions and table
// look-ups for every one of the k^D coefficients), the element -> (low, high) index split is tabulated once
// per CTA, and the cell offsets of all multi-levels are computed by the lanes in parallel.
// ~70 instead of ~480 instructions per multi-level and lane at D=4, k=4: 1.5 M -> see DESIGN.md 4.7.
__global__ void __launch_bounds__(256)
reconstruct2_kernel(ReconTables T, const double* __restrict__ coeffs, const double* __restrict__ pts,
                    long long npts, double* __restrict__ out, int KL, int KH, int nlow) {
    extern __shared__ __align__(16) unsigned char rsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int D = T.D, k = T.k, n1 = T.n + 1;
    const int ntab = D * n1 * k;
    const int KLH = KL + KH;
    // per CTA: lohi[KD] (element -> low / high product index), pdig[KL+KH] (product index -> packed mode
    // digits), lvl[nblocks*D] (block levels); per warp: boff[nblocks] (8 B), bt[ntab], plh[KL+KH], ci[D*n1]
    unsigned* lohi = reinterpret_cast<unsigned*>(rsm);
    const size_t lohi_bytes = ((size_t)T.KD * 4 + 15) & ~(size_t)15;
    unsigned* pdig = reinterpret_cast<unsigned*>(rsm + lohi_bytes);
    const size_t pdig_bytes = ((size_t)KLH * 4 + 15) & ~(size_t)15;
    unsigned char* lvl = rsm + lohi_bytes + pdig_bytes;
    const size_t lvl_bytes = ((size_t)T.nblocks * D + 15) & ~(size_t)15;
    const size_t per_warp = ((size_t)T.nblocks * 8 + (size_t)(ntab + KLH) * 8 + (size_t)D * n1 * 4 + 15) & ~(size_t)15;
    unsigned char* wb = rsm + lohi_bytes + pdig_bytes + lvl_bytes + (size_t)warp * per_warp;
    long long* boff = reinterpret_cast<long long*>(wb);
    double* bt = reinterpret_cast<double*>(wb + (size_t)T.nblocks * 8);
    double* plh = bt + ntab;
    int* ci = reinterpret_cast<int*>(plh + KLH);
    const double sqrt2 = sqrt(2.0);
    const int nhigh = D - nlow;

    for (int e = threadIdx.x; e < T.KD; e += blockDim.x) lohi[e] = (unsigned)(e % KL) | ((unsigned)(e / KL) << 16);
    for (int j = threadIdx.x; j < KLH; j += blockDim.x) {       // up to 6 digits of 4 bits (k <= 10, D <= 12)
        int rem = j < KL ? j : j - KL;
        const int nd = j < KL ? nlow : nhigh;
        unsigned pk = 0;
        for (int d = 0; d < nd; ++d) { pk |= (unsigned)(rem % k) << (4 * d); rem /= k; }
        pdig[j] = pk;
    }
    for (int i = threadIdx.x; i < T.nblocks * D; i += blockDim.x) lvl[i] = T.blk_level[i];
    __syncthreads();

    for (long long pt = (long long)blockIdx.x * nwarp + warp; pt < npts; pt += (long long)gridDim.x * nwarp) {
        __syncwarp();
        {   // a point outside [0, 1]^D (or NaN) is a BoundsError in the reference: flag it, never dereference
            bool bad = false;
            for (int d = lane; d < D; d += 32) {
                const double xd = pts[pt * D + d];
                bad = bad || !(xd >= 0.0 && xd <= 1.0);
            }
            if (__any_sync(0xffffffffu, bad)) {
                if (lane == 0) out[pt] = __longlong_as_double(0x7ff8000000000000LL);
                continue;
            }
        }
        for (int idx = lane; idx < ntab; idx += 32) {
            const int m = idx % k, dl = idx / k, l = dl % n1, d = dl / n1;
            const double x = pts[pt * D + d];
            long long cell;   // 1-based, src/dg_methods.jl:70-79
            if (l <= 1) cell = 1;
            else if (x >= 1.0) cell = 1LL << (l - 1);
            else cell = 1 + (long long)floor((double)(1LL << (l - 1)) * x);
            double val;
            if (l == 0) {
                val = poly_eval(T.leg + m * T.legw, T.legw / 2, 2.0 * x - 1.0) * sqrt2;
            } else {
                const double sc = (double)(1LL << l);
                val = poly_eval(T.dg + m * 2 * k, k, sc * x - (double)(2 * cell - 1)) * sqrt(sc);
            }
            bt[idx] = val;
            if (m == 0) ci[dl] = (int)(cell - 1);
        }
        __syncwarp();
        for (int b = lane; b < T.nblocks; b += 32) {          // cell offsets of all multi-levels
            const unsigned char* lv = lvl + (size_t)b * D;
            long long lin = 0, stride = 1;
            for (int d = 0; d < D; ++d) {
                const int l = lv[d];
                lin += (long long)ci[d * n1 + l] * stride;
                stride *= (l <= 1) ? 1 : (1LL << (l - 1));
            }
            boff[b] = T.blk_offset[b] + lin * T.KDp;
        }
        __syncwarp();
        double acc = 0.0;
        for (int b = 0; b < T.nblocks; ++b) {
            const unsigned char* lv = lvl + (size_t)b * D;
            for (int j = lane; j < KLH; j += 32) {
                unsigned pk = pdig[j];
                const int d0 = j < KL ? 0 : nlow, d1 = j < KL ? nlow : D;
                double prod = 1.0;
                for (int d = d0; d < d1; ++d) {
                    prod *= bt[(d * n1 + lv[d]) * k + (pk & 15u)];
                    pk >>= 4;
                }
                plh[j] = prod;
            }
            __syncwarp();
            const double* cf = coeffs + boff[b];
            for (int e = lane; e < T.KD; e += 32) {
                const unsigned u = lohi[e];
                acc = fma(cf[e], plh[u & 0xffffu] * plh[KL + (u >> 16)], acc);
            }
            __syncwarp();
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[pt] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// Third-generation reconstruct kernel (round 2; the shipped path for k <= 5, D <= 4): one THREAD per C points.
// The warp-per-point kernels above re-read the k^D coefficients of every multi-level for every point through
// L1/L2 (~1 MB per point at D=4, k=4, n=8) and spend their instructions on index arithmetic.  Here
//   * the points are sorted by a Morton key of their leading bits first (host side, cub radix sort), so the C
//     points of a thread and the 32 threads of a warp fall into the same cell of almost every multi-level: a
//     coefficient is loaded ONCE per thread (one L1 sector per warp) and used for C points;
//   * the sum over the k^D modes is evaluated as the SEPARABLE contraction
//         sum_{m_D} v_D[m_D] ( ... sum_{m_2} v_2[m_2] ( sum_{m_1} v_1[m_1] c[m_1, m_2, .., m_D] ) ... )
//     in memory order: k^D + k^(D-1) + ... + k FMAs and D accumulators instead of D multiplications per mode;
//   * the D*k one-dimensional basis values of a point live in registers and are re-evaluated only for the
//     dimensions whose level changed between consecutive multi-levels (the layout order changes dimension 1
//     fastest); the basis coefficient tables are kernel parameters (constant-bank operands).
// Basis values follow `v` / `array2poly` op for op (un-contracted Horner, src/1d_dg_functions.jl:15-28).
// ------------------------------------------------------------------------------------------
template <int K>
struct ReconBasis {                // leg[m][i], dg[m][i] for i < K (lower half) and the sgn-part dg[m][K + i]
    double leg[K][K];
    double dg[K][2 * K];
};

template <int K, int C, int DIM>
struct ReconContract {
    // res[c] = contraction of the sub-block at cf (dimensions 0..DIM) with the basis values of point c
    template <int NV>
    static __device__ __forceinline__ void run(const double* const (&cf)[NV], const double (&v)[C][4][K], double (&res)[C]) {
        constexpr int STRIDE = DIM == 0 ? 1 : (DIM == 1 ? K : (DIM == 2 ? K * K : K * K * K));
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = 0.0;
        if constexpr (DIM == 0 && NV == 1 && (K % 2) == 0) {
            // k even: every row of k coefficients starts 16-byte aligned (cells are padded to an even length)
#pragma unroll
            for (int m = 0; m < K; m += 2) {
                const double2 cv = __ldg(reinterpret_cast<const double2*>(cf[0] + m));
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    res[c] = fma(cv.x, v[c][0][m], res[c]);
                    res[c] = fma(cv.y, v[c][0][m + 1], res[c]);
                }
            }
        } else if constexpr (DIM == 0) {
#pragma unroll
            for (int m = 0; m < K; ++m) {
                if constexpr (NV == 1) {
                    const double cv = __ldg(cf[0] + m);
#pragma unroll
                    for (int c = 0; c < C; ++c) res[c] = fma(cv, v[c][0][m], res[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) res[c] = fma(__ldg(cf[c] + m), v[c][0][m], res[c]);
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const double* sub[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) sub[i] = cf[i] + m * STRIDE;
                double r[C];
                ReconContract<K, C, DIM - 1>::template run<NV>(sub, v, r);
#pragma unroll
                for (int c = 0; c < C; ++c) res[c] = fma(r[c], v[c][DIM][m], res[c]);
            }
        }
    }
};

template <int K>
__device__ __forceinline__ double recon_poly(const double* __restrict__ lo, const double* __restrict__ hi, double x) {
    // array2poly restricted to the K coefficients that can be non-zero (the skipped leading terms are exact zeros)
    if (fabs(x) > 1.0) return 0.0;
    const bool neg = signbit(x);
    double s = 0.0;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        const double t = hi ? (neg ? -hi[i] : hi[i]) : 0.0;
        s = __dadd_rn(__dmul_rn(s, x), lo[i] + t);
    }
    return s;
}

#ifndef GSG_RECON3_MINB
#define GSG_RECON3_MINB 3
#endif
template <int K, int D, int C>
__global__ void __launch_bounds__(128, GSG_RECON3_MINB)
reconstruct3_kernel(ReconTables T, const __grid_constant__ ReconBasis<K> B, const double* __restrict__ coeffs,
                    const double* __restrict__ pts, const unsigned* __restrict__ perm, long long npts,
                    double* __restrict__ out) {
    static_assert(D >= 1 && D <= 4, "instantiated for D <= 4");
    const long long t0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * C;
    if (t0 >= npts) return;
    const double sqrt2 = sqrt(2.0);
    long long idx[C];
    double x[C][D];
    bool bad[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const long long i = t0 + c < npts ? t0 + c : npts - 1;          // tail: recompute the last point
        idx[c] = perm[i];
        bad[c] = false;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            x[c][d] = pts[idx[c] * D + d];
            bad[c] = bad[c] || !(x[c][d] >= 0.0 && x[c][d] <= 1.0);      // BoundsError in the reference: flag, never index
            if (bad[c]) x[c][d] = 0.0;
        }
    }
    double v[C][4][K];
    int cell[C][D];
    int lprev[D];
#pragma unroll
    for (int d = 0; d < D; ++d) lprev[d] = -1;
    double acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.0;

    for (int b = 0; b < T.nblocks; ++b) {
        const unsigned char* lv = T.blk_level + (size_t)b * D;
        int l[D];
#pragma unroll
        for (int d = 0; d < D; ++d) l[d] = lv[d];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (l[d] == lprev[d]) continue;
            lprev[d] = l[d];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double xd = x[c][d];
                long long cl;                                            // 1-based, src/dg_methods.jl:70-79
                if (l[d] <= 1) cl = 1;
                else if (xd >= 1.0) cl = 1LL << (l[d] - 1);
                else cl = 1 + (long long)floor((double)(1LL << (l[d] - 1)) * xd);
                cell[c][d] = (int)(cl - 1);
                if (l[d] == 0) {
                    const double y = 2.0 * xd - 1.0;
#pragma unroll
                    for (int m = 0; m < K; ++m) v[c][d][m] = recon_poly<K>(B.leg[m], nullptr, y) * sqrt2;
                } else {
                    const double sc = (double)(1LL << l[d]);
                    const double y = sc * xd - (double)(2 * cl - 1);
                    const double sq = sqrt(sc);
#pragma unroll
                    for (int m = 0; m < K; ++m) v[c][d][m] = recon_poly<K>(B.dg[m], B.dg[m] + K, y) * sq;
                }
            }
        }
        const long long boff = T.blk_offset[b];
        const double* cf[C];
        bool same = true;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            long long lin = 0, stride = 1;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                lin += (long long)cell[c][d] * stride;
                stride *= (l[d] <= 1) ? 1 : (1LL << (l[d] - 1));
            }
            cf[c] = coeffs + boff + lin * T.KDp;
            same = same && cf[c] == cf[0];
        }
        double r[C];
        if (same) {
            const double* const one[1] = {cf[0]};
            ReconContract<K, C, D - 1>::template run<1>(one, v, r);
        } else {
            const double* many[C];
#pragma unroll
            for (int c = 0; c < C; ++c) many[c] = cf[c];
            ReconContract<K, C, D - 1>::template run<C>(many, v, r);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += r[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (t0 + c < npts) out[idx[c]] = bad[c] ? __longlong_as_double(0x7ff8000000000000LL) : acc[c];
}

// Morton key of the leading `bits` bits of every coordinate (dimension 0 in the least significant position)
__global__ void recon_keys_kernel(const double* __restrict__ pts, long long npts, int D, int bits, unsigned* __restrict__ keys,
                                  unsigned* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    unsigned key = 0;
    const double scale = (double)(1u << bits);
    for (int d = 0; d < D; ++d) {
        const double xd = pts[i * D + d];
        unsigned q = (xd >= 0.0 && xd < 1.0) ? (unsigned)(xd * scale) : (xd >= 1.0 ? (1u << bits) - 1u : 0u);
        for (int bb = 0; bb < bits; ++bb) key |= ((q >> bb) & 1u) << (bb * D + d);
    }
    keys[i] = key;
    vals[i] = (unsigned)i;
}

// ------------------------------------------------------------------------------------------
// Flat sweep for SMALL index sets (launch-latency regime: BASELINE configs 2 and 3, N = 1e4 .. 3e5): ONE launch
// computes  y = beta * y + sum_{d in list} c_d * M_d x  for every direction of the list, with no global atomics and
// no tiling by pole class.  A CTA owns `cpc` consecutive multi-cells.  Its warps share out the work items
// (direction, cell, chunk of <= 32 poles, record slice): the lanes of a warp are the poles of the chunk (x is
// gathered straight from global memory -- the whole state of these configurations sits in L2 -- and consecutive poles
// read consecutive addresses), times 32 / PIp record sub-lanes when a cell has fewer than 32 poles; a lane walks its
// share of the block row of the item's 1-D cell (block CSR of the 1-D matrix, cut at the end of the principal
// sub-block of the pole's class) in a counted, unrolled loop, the sub-lanes are reduced with shuffles, and the K
// outputs of a pole are added into the warp's PRIVATE copy of the CTA's output cells in shared memory (no races:
// one item at a time per warp, distinct poles per lane).  One barrier, then the copies are summed into y.
//   FlatCD table (ints, FLATCD stride per (multi-cell, direction)): {row begin, row end (derivative blocks),
//   S, class p << 16 | 1-D cell q, then for level l = 0..p the multi-cell index of the item's first cell of level l}
//   -- everything a record needs besides its block column, so the dependent-load chain is table -> column -> x.
//   FlatMat: block records; the pre-squared Laplacian blocks S_p = (H[0:N',0:N'])^2 of the short classes
//   (src/multidim_derivative.jl:71-79) have their own records per class: row q of class p is
//   [rowptr[cls_row0[p] + q], rowend[cls_row0[p] + q]).
//   Only items whose class lies in [pmin, pmax] are processed.
// ------------------------------------------------------------------------------------------
struct FlatCell {             // host-side description of a (multi-cell, direction): pole group, item, 1-D cell
    int group, r, q;
};

constexpr int FLAT_MAXD = 12;
constexpr int FLATCD = 4 + ((MAXL + 1 + 3) & ~3);      // ints per (multi-cell, direction)

struct FlatDirs {
    double c[FLAT_MAXD];
    int A[FLAT_MAXD];
    int d[FLAT_MAXD];
    int ndir;
};

struct FlatMat {
    const int* rowptr;        // squared blocks only (sq != 0)
    const int* rowend;
    const int* col;
    const double* val;
    int KK2;
    int sq;
    int cls_row0[MAXL + 1];
};

constexpr int FLAT_THREADS = 256;
constexpr int FLAT_WARPS = FLAT_THREADS / 32;

constexpr int FLAT_U = 4;          // records per sub-lane and item: their gathers are in flight together

template <int K>
__global__ void __launch_bounds__(FLAT_THREADS)
sweep_flat_kernel(const double* __restrict__ X, double* __restrict__ Y, double beta, const FlatDirs fd,
                  const int* __restrict__ celltab, int D, int ncells, int cpc, const FlatMat M, int KD, int KDp,
                  int PI, int PIp, int pmin, int pmax) {
    extern __shared__ __align__(16) double acc_s[];          // FLAT_WARPS private copies of cpc * KDp outputs, then po_s
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.x * cpc;
    const int nc = min(cpc, ncells - c0);
    const int slab = cpc * KDp;
    double* mine = acc_s + warp * slab;
    int* po_s = reinterpret_cast<int*>(acc_s + FLAT_WARPS * slab);      // [ndir][PI]: in-cell offset of pole j along d
    for (int i = lane; i < slab; i += 32) mine[i] = 0.0;
    for (int i = tid; i < fd.ndir * PI; i += FLAT_THREADS) {
        const int di = i / PI, j = i - di * PI, A = fd.A[di];
        const int b = j / A;
        po_s[i] = (j - b * A) + K * A * b;
    }
    // longest block row among this CTA's (cell, direction) pairs (nc * ndir <= 32): coarse cells (rows of up to 2^n
    // records) are cut into more record slices, so that an item is ONE batch of FLAT_U records per sub-lane and the
    // CTA's warps share a long row.  The slice count is a power of two (item decoding by shifts).
    int len = 0;
    if (lane < nc * fd.ndir) {
        const int* cdp = celltab + ((size_t)(c0 + lane % nc) * D + fd.d[lane / nc]) * FLATCD;
        len = __ldg(cdp + 1) - __ldg(cdp);
    }
    len = __reduce_max_sync(0xffffffffu, len);
    const int jl = lane & (PIp - 1), sub = lane / PIp, G = 32 / PIp;     // pole inside the chunk, record sub-lane
    const int nch = (PI + PIp - 1) / PIp;
    int rs_log = 0;
    while (rs_log < 6 && ((G * FLAT_U) << rs_log) < len) ++rs_log;
    const int RS = 1 << rs_log;
    const int per_dir = nc * nch;
    const int nitems = fd.ndir * per_dir * RS;
    __syncthreads();
    for (int it = warp; it < nitems; it += FLAT_WARPS) {      // warp-uniform control flow throughout
        const int rsl = it & (RS - 1);
        int t = it >> rs_log;
        const int di = t / per_dir;
        t -= di * per_dir;
        const int cl = nc == 1 ? 0 : t / nch;
        const int ch = t - cl * nch;
        const int A = fd.A[di], dd = fd.d[di];
        // the item's table record, one coalesced load: lanes 0..3 = {row begin, row end, S, p << 16 | q}, lane 4 + l =
        // multi-cell index of the item's first cell of level l (fetched with shuffles: no dependent table load later)
        const int tv = lane < FLATCD ? __ldg(celltab + ((size_t)(c0 + cl) * D + dd) * FLATCD + lane) : 0;
        const int pq = __shfl_sync(0xffffffffu, tv, 3);
        const int p = pq >> 16;
        if (p < pmin || p > pmax) continue;
        int rbeg = __shfl_sync(0xffffffffu, tv, 0), rend = __shfl_sync(0xffffffffu, tv, 1);
        const int S = __shfl_sync(0xffffffffu, tv, 2);
        if (M.sq) {
            const int rb = M.cls_row0[p] + (pq & 0xffff);
            rbeg = __ldg(M.rowptr + rb);
            rend = __ldg(M.rowend + rb);
        }
        const int step = G << rs_log;
        int rec = rbeg + rsl * G + sub;
        if (__all_sync(0xffffffffu, rec >= rend)) continue;          // empty slice
        const int j = ch * PIp + jl;
        const bool valid = j < PI;
        const int po = po_s[di * PI + (valid ? j : 0)];
        double r[K];
#pragma unroll
        for (int m = 0; m < K; ++m) r[m] = 0.0;
        // FLAT_U records at a time: columns and block values first (they depend on the record only), then the x gathers
        for (; __any_sync(0xffffffffu, rec < rend); rec += FLAT_U * step) {
            double xr[FLAT_U][K], hv[FLAT_U][K * K];
            int qc[FLAT_U];
            bool on[FLAT_U];
#pragma unroll
            for (int u = 0; u < FLAT_U; ++u) {
                const int r = rec + u * step;
                on[u] = r < rend;
                const int rr = on[u] ? r : rbeg;
                qc[u] = __ldg(M.col + rr);
                const double* hp = M.val + (size_t)rr * M.KK2;
#pragma unroll
                for (int e = 0; e < K * K; ++e) hv[u][e] = __ldg(hp + e);
            }
#pragma unroll
            for (int u = 0; u < FLAT_U; ++u) {
                int ld, cdv, Cd;
                q_decode(qc[u], ld, cdv, Cd);
                const long long cell = (long long)__shfl_sync(0xffffffffu, tv, 4 + ld) + (long long)S * cdv;
                const double* xv = X + cell * KDp + po;
#pragma unroll
                for (int mi = 0; mi < K; ++mi) xr[u][mi] = on[u] ? xv[A * mi] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < FLAT_U; ++u)
#pragma unroll
                for (int mo = 0; mo < K; ++mo)
#pragma unroll
                    for (int mi = 0; mi < K; ++mi) r[mo] = fma(hv[u][mo * K + mi], xr[u][mi], r[mo]);
        }
        for (int o = PIp; o < 32; o <<= 1) {
#pragma unroll
            for (int m = 0; m < K; ++m) r[m] += __shfl_xor_sync(0xffffffffu, r[m], o);
        }
        if (valid && sub == 0) {
            const double cd = fd.c[di];
#pragma unroll
            for (int m = 0; m < K; ++m) mine[cl * KDp + po + A * m] += cd * r[m];
        }
        __syncwarp();                    // the next item of this warp may touch the same outputs from other lanes
    }
    __syncthreads();
    double* yo = Y + (size_t)c0 * KDp;
    for (int i = tid; i < nc * KDp; i += FLAT_THREADS) {
        if (nc > 1 ? (i % KDp < KD) : (i < KD)) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < FLAT_WARPS; ++w) sum += acc_s[w * slab + i];
            yo[i] = beta == 0.0 ? sum : fma(beta, yo[i], sum);
        }
    }
}

// ------------------------------------------------------------------------------------------
// CSR SpMV cross-check: LANES lanes per row, int32 columns.
// ------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(256)
spmv_csr_kernel(long long nrows, const long long* __restrict__ rowptr, const int* __restrict__ col,
                const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gid / LANES;
    const int sub = (int)(gid % LANES);
    double s = 0.0;
    if (row < nrows) {
        const long long e = rowptr[row + 1];
        for (long long p = rowptr[row] + sub; p < e; p += LANES) s = fma(val[p], __ldg(x + col[p]), s);
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < nrows && sub == 0) y[row] = s;
}

}  // namespace gsgk
