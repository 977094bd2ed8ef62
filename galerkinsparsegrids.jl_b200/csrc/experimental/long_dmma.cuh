// long_dmma.cuh -- EXPERIMENTAL, NOT COMPILED INTO libgsgb200.so (round 1).
// Long-pole sweep kernel on the fp64 tensor pipe (mma.sync.m8n8k4.f64) with host-side
// fragment-ordered 24x12 tile records.  Validated on B200: parity 1.2e-16 relative against the
// oracle on every test case, compute-sanitizer clean (commit "Long-pole kernel on the fp64 tensor
// pipe").  Measured (profiles/r1_long_kernel_variants.md): not faster than the 3x3-record kernel in
// kernels.cuh because both are bound by per-tile setup / flush and 9-12 % warp occupancy, not by the
// FMA loop; p = 8 degenerates to one warp per CTA (x tile 197 KB).  Kept as the starting point for
// round 2 (smaller per-warp scratch, 16-pole tiles for p >= 7, persistent CTAs).  The matching host
// record builder is in git history (gsg_b200.cu of the same commit).
#if 0
// ------------------------------------------------------------------------------------------
// Long poles (p >= 4 at k = 3): batched dense-tile mat-vecs on the fp64 tensor pipe (DMMA).
// A CTA tile holds PT <= 32 poles (a sub-range a0..a0+na x b0..b0+nb of one item's poles, or nr
// whole items when an item has few poles).  The x tile sits in shared memory pole-major,
// xsT[pole][row] with a row stride = 4 (mod 32) doubles so that DMMA B-fragment loads are
// conflict-free.  The principal sub-block of class p is cut into TR x TCc tiles (24 x 12 at
// k = 3); tiles holding a stored entry are kept as a stream of records, pre-arranged on the host
// in mma.m8n8k4 A-fragment order, plus {column tile, end-of-row-tile flag} (25-45 % of the tiles
// at p = 8, all of them at p = 4).  Feeding H to scalar DFMAs from shared memory is bound by
// the 128 B/clk LDS delivery rate (8 B per lane per FMA ~ 25 % of fp64 peak, measured); the
// DMMA path loads every operand once per lane for 256 FMAs (~0.75 B/FMA).  Every warp owns a
// contiguous range of row tiles (host partition balanced in record count) and streams its records
// through a private cp.async ring; a finished row tile is written through a per-warp scratch so
// the global write is coalesced.
//   in-item order t -> a = t % na, m = (t / na) % K, bl = t / (K*na):
//   pole = a + na*bl, in-cell offset = ebase + a + A*m + K*A*bl.
// ------------------------------------------------------------------------------------------
struct TileLong {
    int group;
    int r0;
    int ebase;
    short nr, na, nb, part;   // part: which row part of the matrix this CTA computes
};

template <int K>
struct LongTile {
    // rows: multiple of 8 (DMMA M) and of K (whole 1-D cells); cols: multiple of 4 (DMMA K) and K
    static constexpr int TR = (K == 3) ? 24 : (K == 5 ? 40 : 16);
    static constexpr int TCc = TR / 2;
    static constexpr int MB = TR / 8, KB = TCc / 4;
    static constexpr int BYTES = TR * TCc * 8 + 16;         // fragments + {col tile, flags, pad}
};

__host__ __device__ constexpr int long_xs_stride(int NP) { return ((NP + 27) / 32) * 32 + 4; }   // = 4 (mod 32), >= NP

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int K, int NBUF>
__global__ void __launch_bounds__(256, 1)
sweep_long_kernel(const double* __restrict__ X, double* __restrict__ Y, double alpha, double beta,
                  const GroupDev* __restrict__ groups, const TileLong* __restrict__ tiles,
                  const unsigned char* __restrict__ recs, const int* __restrict__ partRec,
                  const int* __restrict__ partRow, int p, int KDp, int A) {
    using LT = LongTile<K>;
    constexpr int TR = LT::TR, TCc = LT::TCc, MB = LT::MB, KB = LT::KB, REC = LT::BYTES;
    constexpr int CPT = TR / K;                               // 1-D cells per row tile
    constexpr int WARP_BYTES = NBUF * REC + TR * 32 * 8;      // ring + scratch[TR][32] per warp
    const int NQ = 1 << p, NP = K * NQ;
    const int XS = long_xs_stride(NP);
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ long long sbase[MAXL + 1];
    __shared__ int sS;
    __shared__ short tab_p[K * 32];     // pole-in-item of in-item element t
    __shared__ short tab_m[K * 32];     // mode m of in-item element t
    __shared__ int tab_g[K * 32];       // in-cell offset of in-item element t

    const TileLong t = tiles[blockIdx.x];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarp = nth >> 5;
    const int na = t.na, nb = t.nb, nr = t.nr;
    const int PIt = na * nb;            // poles per item in this tile
    const int TL = K * PIt;
    double* xsT = reinterpret_cast<double*>(smraw);                                  // 32 * XS
    long long* caddr = reinterpret_cast<long long*>(smraw + (size_t)32 * XS * 8);    // NQ * nr cell offsets
    const int caddr_bytes = (NQ * nr * 8 + 15) & ~15;
    unsigned char* wbase = smraw + (size_t)32 * XS * 8 + caddr_bytes + (size_t)warp * WARP_BYTES;
    unsigned char* ring = wbase;
    double* scratch = reinterpret_cast<double*>(wbase + NBUF * REC);

    // this warp's records [b0, b1) and first row tile
    const int gpart = t.part * nwarp + warp;
    const int b0 = partRec[gpart], b1 = partRec[gpart + 1];
    int rt = partRow[gpart];

    auto issue_rec = [&](int b) {
        if (b < b1) {
            const unsigned char* src = recs + (size_t)b * REC;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (b % NBUF) * REC);
            for (int g = lane; g < REC / 16; g += 32) cp_async16(dst + g * 16, src + g * 16);
        }
        cp_async_commit();
    };

    if (tid <= p) sbase[tid] = groups[t.group].base[tid];
    if (tid == 0) sS = groups[t.group].S;
    for (int tt = tid; tt < TL; tt += nth) {
        const int a = tt % na, rest = tt / na;
        const int m = rest % K, bl = rest / K;
        tab_p[tt] = (short)(a + na * bl);
        tab_m[tt] = (short)m;
        tab_g[tt] = t.ebase + a + A * m + K * A * bl;
    }
    // unused pole columns of the x tile must be finite (they are multiplied, never stored)
    for (int i = tid; i < 32 * XS; i += nth) xsT[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NBUF - 1; ++i) issue_rec(b0 + i);
    __syncthreads();
    const int S = sS;
    const int ncell = NQ * nr;
    for (int c = tid; c < ncell; c += nth) {
        const int qq = c / nr, r = c - qq * nr;
        caddr[c] = cell_addr(sbase, S, qq, t.r0 + r, KDp);
    }
    __syncthreads();

    // ---- stage the x tile in (asynchronous 8-byte copies into xsT[pole][row])
    for (int c = warp; c < ncell; c += nwarp) {
        const int qq = c / nr, r = c - qq * nr;
        const double* src = X + caddr[c];
        const unsigned dst = (unsigned)__cvta_generic_to_shared(xsT + (size_t)(r * PIt) * XS + qq * K);
        for (int tt = lane; tt < TL; tt += 32) cp_async8(dst + (tab_p[tt] * XS + tab_m[tt]) * 8, src + tab_g[tt]);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- stream this warp's tile records through the tensor pipe
    const int lr = lane >> 2, lc = lane & 3;          // fragment row / col of this lane
    double acc[MB][4][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nbk = 0; nbk < 4; ++nbk) acc[mb][nbk][0] = acc[mb][nbk][1] = 0.0;
    for (int b = b0; b < b1; ++b) {
        issue_rec(b + NBUF - 1);
        cp_async_wait<NBUF - 1>();         // record b has landed (this thread's copies) ...
        __syncwarp();                      // ... and every lane's
        const unsigned char* rec = ring + (b % NBUF) * REC;
        const int2 meta = *reinterpret_cast<const int2*>(rec + TR * TCc * 8);
        const double* afr = reinterpret_cast<const double*>(rec) + lane;
        const double* xcol = xsT + (size_t)lr * XS + meta.x * TCc + lc;     // B[k = lc][n = lr] of pole block 0
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            double bf[4];
#pragma unroll
            for (int nbk = 0; nbk < 4; ++nbk) bf[nbk] = xcol[(size_t)(nbk * 8) * XS + kb * 4];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const double a = afr[(mb * KB + kb) * 32];
#pragma unroll
                for (int nbk = 0; nbk < 4; ++nbk) dmma_m8n8k4(acc[mb][nbk][0], acc[mb][nbk][1], a, bf[nbk]);
            }
        }
        if (meta.y & 1) {              // end of row tile rt: scratch[row][pole], then coalesced writes
            __syncwarp();
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nbk = 0; nbk < 4; ++nbk)
                    *reinterpret_cast<double2*>(scratch + (mb * 8 + lr) * 32 + nbk * 8 + 2 * lc) =
                        make_double2(acc[mb][nbk][0], acc[mb][nbk][1]);
            __syncwarp();
            for (int cc = 0; cc < CPT; ++cc) {
                const int q = rt * CPT + cc;
                for (int r = 0; r < nr; ++r) {
                    double* dstg = Y + caddr[q * nr + r];
                    const double* sc = scratch + cc * K * 32 + r * PIt;
                    if (beta == 0.0) {
                        for (int tt = lane; tt < TL; tt += 32) dstg[tab_g[tt]] = alpha * sc[tab_m[tt] * 32 + tab_p[tt]];
                    } else {
                        for (int tt = lane; tt < TL; tt += 32) {
                            const int g = tab_g[tt];
                            dstg[g] = fma(alpha, sc[tab_m[tt] * 32 + tab_p[tt]], beta * dstg[g]);
                        }
                    }
                }
            }
            ++rt;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nbk = 0; nbk < 4; ++nbk) acc[mb][nbk][0] = acc[mb][nbk][1] = 0.0;
        }
        __syncwarp();                  // the ring slot is refilled by the next iteration's issue
    }
    cp_async_wait<0>();
}

#endif
