// gsg_b200.cu -- plan construction, kernel dispatch and the C ABI of libgsgb200.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see __graft_entry__.build()).
#include "../../include/gsg_b200.h"
#include "host_setup.hpp"
#include "kernels.cuh"

#include <nvtx3/nvToolsExt.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace gsgk;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define GSG_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            return fail(GSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
        }                                                                                      \
    } while (0)

#define GSG_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != 0) return rc__;\
    } while (0)

// NVTX range (tracing): visible in Nsight Systems / ncu --nvtx, a no-op without a tool attached
struct nvtx_range {
    explicit nvtx_range(const char* name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
    nvtx_range(const nvtx_range&) = delete;
    nvtx_range& operator=(const nvtx_range&) = delete;
};

template <class T>
struct DevBuf {                       // move-only owner of a device allocation
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            if (p) cudaFree(p);
            p = o.p; n = o.n;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { if (p) cudaFree(p); }
    int upload(const std::vector<T>& h) {
        if (p) { cudaFree(p); p = nullptr; }
        n = h.size();
        if (n == 0) return 0;
        GSG_CUDA(cudaMalloc(&p, n * sizeof(T)));
        GSG_CUDA(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
        // a pageable host-to-device cudaMemcpy may return once the data is staged, before the DMA has landed; the plans'
        // streams are non-blocking (not ordered behind the legacy stream), so a table built lazily right before a launch
        // (flat tables, tensor_construct's cell table) must be waited for explicitly
        GSG_CUDA(cudaStreamSynchronize(0));
        return 0;
    }
    // grow-only; new allocations are zero-filled (the padding slots of state vectors stay finite)
    int resize(size_t m) {
        if (m <= n && p) return 0;
        if (p) { cudaFree(p); p = nullptr; }
        n = m;
        GSG_CUDA(cudaMalloc(&p, n * sizeof(T)));
        GSG_CUDA(cudaMemset(p, 0, n * sizeof(T)));
        // cudaMemset on device memory is asynchronous (legacy stream) and the plans' streams are non-blocking: without
        // this wait a kernel on the plan's stream could run before the fill and have its output zeroed afterwards
        GSG_CUDA(cudaStreamSynchronize(0));
        return 0;
    }
};

constexpr int SHORT_MAX_NP = 32;      // register-resident poles up to this length
constexpr int SHORT_MAX_P = 3;
constexpr size_t GENERIC_SMEM_BUDGET = 96 * 1024;
constexpr size_t SMEM_OPTIN_MAX = 227 * 1024;
constexpr int SHORT_TILE_DOUBLES = 4096;
constexpr int TMA_STAGE_TARGET_DOUBLES = 5840;

enum class Kind { SHORT_TMA, LONG, LONG2, CONSTH, GENERIC, ROWTILE };   // LONG: transient tag while a class is being placed

struct SweepClass {          // one launch of a sweep
    int p = 0;               // pole length class (SHORT_TMA: the largest short class it holds)
    Kind kind = Kind::GENERIC;
    int NPOLE = 0, Amin = 0; // generic kernel parameters
    int nwarps = 8;          // long kernel: warps per CTA
    int rsplit = 1;          // long kernel: row parts (CTAs) per pole set
    int cpl = 1;             // register-tiled long kernel: poles per lane
    int npass = 1;           // ... and column passes
    int nbuf = 4;            // ... and record-ring depth (2 with 16 warps)
    std::vector<std::unique_ptr<DevBuf<int>>> passBlk, passRow;
    DevBuf<TileL2> l2tiles;
    size_t smem = 0;
    DevBuf<TileDev> tiles;
    DevBuf<TileS> stiles;
    DevBuf<TileS2> s2tiles;     // second-generation streaming kernel
    DevBuf<RTWork> rtwork;      // row-tile kernel: (item, tile) work list, heaviest tiles first
    DevBuf<long long> rtsrc;    // ... and per work item the offsets in x of the tile's RT_MAXX cells
    DevBuf<int> rtsched;        // ... and the persistent CTAs' item lists: sched[k * rtgrid + b] = k-th item of CTA b, -1 = end
    int rtgrid = 0;
    DevBuf<int> rtzero;         // ... and the multi-cells that receive partial sums (zeroed before a beta = 0 sweep)
    bool stream2 = false;
    ShortParams sprm{};
    int ntiles = 0;
    double short_dofs = 0;   // SHORT_TMA: DOFs this launch processes (for the roofline figure)
};

struct Direction {
    int A = 1;               // K^(d-1)
    DevBuf<GroupDev> groups;
    std::vector<GroupDev> groups_h;
    DevBuf<CellOfs> celltab;     // register-tiled long kernel: per group, one entry per 1-D cell
    DevBuf<int> offtab;          // in-cell offset of pole j: a + K*A*b
    DevBuf<int> rt_offtab;       // row-tile kernel: the same offsets in an order whose consecutive poles are free of
                                 // shared-memory bank conflicts (b fastest while A < 16: the stride K*A is odd)
    std::vector<SweepClass> classes;
};

}  // namespace

#include "rowtile.inl"

struct gsg_plan {
    gsg::IndexSet S;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;

    // 1-D matrix: block CSR (shared by all p, principal sub-blocks) + dense blocks for short poles
    DevBuf<int> b_rowptr, b_col;
    DevBuf<double> b_val;
    int KK2 = 0;
    std::vector<double> dense_host;                        // concatenated, passed as kernel parameter
    int hoff[5] = {0, 0, 0, 0, 0};
    int htotal = 0, short_pmax = -1;
    // register-tiled long kernel: column passes.  The principal sub-block of class p is cut into
    // npass column slices (so that a 32*C-pole x tile of one slice fits shared memory); each slice is
    // its own record stream with block columns relative to the slice
    struct ColPass {
        DevBuf<unsigned char> recs;
        std::vector<int> row_start;      // first record of each block-row (+ end)
        int qc0 = 0, nqc = 0;
    };
    std::map<std::pair<int, int>, std::vector<std::unique_ptr<ColPass>>> lpass;   // key (p, npass)
    std::vector<int> h_rowptr, h_col;    // host copy of the block CSR
    std::vector<double> h_val;
    // constant-bank kernel: per p, the pattern blocks' values in pattern order (empty = not usable)
    std::vector<std::vector<double>> consth_vals;

    std::vector<Direction> dirs;
    // direction-pair fusion (streaming classes): pair j = dimensions (2j, 2j+1).  pair_np = largest n' of the
    // 2-D sub-planes handled by PAIR tiles (-1 = off); dirs_red = the directions WITHOUT the pole groups those
    // tiles cover (used only by the fused gradient / advection right-hand side)
    int pair_np = -1;
    std::vector<Direction> dirs_red;
    std::vector<SweepClass> pairs;

    // fork/join streams so the launches of one sweep run concurrently
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> ev_done;
    cudaEvent_t ev_fork = nullptr, ev_pair_fork = nullptr, ev_pair_done = nullptr;

    DevBuf<int> cell_block;          // block of every multi-cell (device tensor_construct)
    DevBuf<double> w1d;              // D concatenated 1-D coefficient vectors
    // reconstruct tables
    DevBuf<unsigned char> r_level;
    DevBuf<long long> r_offset;
    DevBuf<double> r_leg, r_dg;

    // workspaces (device layout)
    DevBuf<double> wx, wy, wk, wacc, ww, wtmp, wred;
    DevBuf<double> wpts, wout;
    DevBuf<unsigned> rkeys, rkeys2, rperm, rperm2;     // reconstruct: Morton keys / permutation (sorted points)
    DevBuf<unsigned char> rtemp;
    DevBuf<int> tile_counter;     // dynamic tile schedulers of the persistent kernels: one slot per launch, round robin
    int ctr_next = 0;
    // concurrent right-hand side (rhs_concurrent): pool of high-priority streams for the long-pole launches of ALL
    // directions, forked after the first (beta = 0) pieces and joined at the end
    std::vector<cudaStream_t> pool;        // [0, 16): high priority (long-pole launches); [16, 24): low priority
    std::vector<cudaEvent_t> pool_ev;      //           (SM-filling persistent kernels: their ramp-up / drain overlap)
    cudaEvent_t ev_p1 = nullptr;
    long long* dbg = nullptr;     // optional clock-stamp buffer (gsg_debug_stamps)
    DevBuf<long long> dbgbuf;

    // row-tile kernel for the long-pole classes (rowtile.inl): tile programs of the classes rt_pmin..n
    bool rt_on = false;
    int rt_pmin = 0, rt_C = 1, rt_PW = 1, rt_RG = 1;
    size_t rt_smem = 0;
    RTProgram rt_prog;
    DevBuf<RTTile> rt_tiles;
    DevBuf<RTGroup> rt_groups;
    DevBuf<unsigned char> rt_recs;

    // flat path (kernels.cuh, sweep_flat_kernel): one launch per right-hand side for small index sets.
    // flat_mode: 0 = never, 1 = whenever supported, 2 = automatic (N * D <= FLAT_AUTO_MAX)
    int flat_mode = 2;
    DevBuf<int> flat_cd;                                   // FLATCD ints per (multi-cell, direction)
    DevBuf<int> sq_rowend;
    // pre-squared Laplacian blocks S_p = (H[0:N',0:N'])^2 of the short classes p <= sq_pmax (dense there anyway;
    // for the long classes the square is completely dense -- level 0 touches every cell -- so those stay D(D x))
    int sq_pmax = -1;
    std::vector<double> dense_sq_host;                     // same layout as dense_host
    DevBuf<int> sq_rowptr, sq_col;
    DevBuf<double> sq_val;
    int sq_cls_row0[MAXL + 1] = {0};
    bool use_sq = false;        // the streaming launches take the squared dense blocks (set by laplacian())
    int kind_filter = 0;        // 0 = every class, 1 = streaming classes only, 2 = long classes only

    // RK4 driver: 0 = automatic (linear right-hand sides use the Taylor form), 1 = always staged
    int rk4_mode = 0;
    cudaGraphExec_t step_exec = nullptr;     // last captured RK4 step (released with the plan / next capture)
    DevBuf<double> wv4;

    // multi-GPU block partition (gsg_plan_set_partition): part_bits dimensions D, D-1, ... each split the
    // multi-level blocks into {level == 0} and {level >= 1}; rank bit j = 1 owns level_{D-j} == 0
    int part_rank = 0, part_bits = 0;

    // optional timing of the dominant (streaming) kernel with CUDA events on its stream
    bool prof_on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    size_t prof_used = 0, prof_cap = 0;
    double prof_dofs = 0;         // DOFs processed by the profiled launches
};

struct gsg_csr {
    int device = 0;
    int64_t m = 0, n = 0, nnz = 0;
    DevBuf<long long> rowptr;
    DevBuf<int> col;
    DevBuf<double> val;
    DevBuf<double> wx, wy;
};

namespace {

int pow_int(int b, int e) {
    int r = 1;
    while (e-- > 0) r *= b;
    return r;
}

bool short_supported(int K, int p) {
    return K >= 1 && K <= 5 && p <= SHORT_MAX_P && (K << p) <= SHORT_MAX_NP;
}

// ------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------
// K x K block CSR (block columns ascending, blocks row-major, padded to KK2 doubles) of a dense N1 x N1 matrix with
// the stored-block mask blk
void block_csr_from_dense(const std::vector<double>& Hd, const std::vector<char>& blk, int64_t N1, int K, int NQ, int KK2,
                          std::vector<int>& rowptr, std::vector<int>& col, std::vector<double>& val) {
    rowptr.assign(NQ + 1, 0);
    col.clear();
    val.clear();
    for (int q = 0; q < NQ; ++q) {
        for (int qc = 0; qc < NQ; ++qc) {
            if (!blk[(size_t)q * NQ + qc]) continue;
            col.push_back(qc);
            const size_t o = val.size();
            val.resize(o + KK2, 0.0);
            for (int mo = 0; mo < K; ++mo)
                for (int mi = 0; mi < K; ++mi)
                    val[o + mo * K + mi] = Hd[(size_t)(q * K + mo) * N1 + (qc * K + mi)];
        }
        rowptr[q + 1] = (int)col.size();
    }
}

// Row-tile kernel for the long-pole classes: eligible when an item has enough poles to fill the lanes (k^(D-1) >= 64,
// i.e. the big index sets: D = 6 / 5 at k = 3, D = 4 at k = 4, 5) and the program fits.  GSG_ROWTILE=0 turns it off
// (the constant-bank / register-tiled kernels then serve those classes), GSG_RT_BUDGET_KB sets the shared-memory
// budget of a tile (default 225 KB: one CTA per SM, the fewest partial sums), GSG_RT_C the poles per lane.
int build_rowtile_program(gsg_plan& P) {
    P.rt_on = false;
    const int K = P.S.k, n = P.S.n;
    const int PI = (int)(P.S.kD / K);
    const char* env = getenv("GSG_ROWTILE");
    if (env && atoi(env) == 0) return 0;
    if (K > 5 || PI < 64 || PI > 512) return 0;
    int pmin = 0;
    while (pmin <= n && short_supported(K, pmin)) ++pmin;
    if (pmin > n) return 0;
    size_t budget = SMEM_OPTIN_MAX - 1024;                // measured at D=6, k=3, n=8: one big tile per SM beats two of half the size
    if (const char* e = getenv("GSG_RT_BUDGET_KB")) budget = (size_t)atoi(e) * 1024;
    budget = std::min(budget, SMEM_OPTIN_MAX - 1024);
    P.rt_C = (PI >= 192 && K <= 3) ? 4 : (PI > 96 && K <= 4) ? 2 : 1;       // k = 5 with two poles per lane spills
    if (const char* e = getenv("GSG_RT_C")) {
        const int c = atoi(e);
        if (c == 1 || (c == 2 && K <= 4) || (c == 4 && K <= 3)) P.rt_C = c;
    }
    P.rt_PW = (PI + 32 * P.rt_C - 1) / (32 * P.rt_C);
    if (P.rt_PW > 8) return 0;
    P.rt_RG = std::max(1, std::min(RT_MAXRG, 8 / P.rt_PW));
    if (rt_build_program(P.h_rowptr, P.h_col, P.h_val, P.KK2, K, (int)P.S.kDp, n, pmin, budget, P.rt_RG, P.rt_prog) != 0) {
        P.rt_prog = RTProgram();
        return 0;                                         // no tiling fits: the other kernels serve these classes
    }
    size_t need = 0;
    for (const RTTile& T : P.rt_prog.tiles)
        need = std::max(need, (size_t)64 + (size_t)(2 * P.rt_RG + T.nx) * P.S.kDp * 8 + (size_t)T.rec_bytes);
    P.rt_smem = need;
    GSG_TRY(P.rt_tiles.upload(P.rt_prog.tiles));
    GSG_TRY(P.rt_groups.upload(P.rt_prog.groups));
    GSG_TRY(P.rt_recs.upload(P.rt_prog.blob));
    P.rt_pmin = pmin;
    P.rt_on = true;
    return 0;
}

int build_matrix(gsg_plan& P, int64_t Hn, const int64_t* colptr, const int64_t* rowval, const double* nzval) {
    const int K = P.S.k;
    const int n = P.S.n;
    const int64_t N1 = int64_t(K) << n;
    if (Hn != N1) return fail(GSG_ERR_ARG, "H must be k*2^n square");
    const int NQ = 1 << n;
    // dense copy is fine: (k 2^n)^2 doubles (4.7 MB at k=3, n=8); guard the size
    if (N1 > 16384) return fail(GSG_ERR_UNSUPPORTED, "k*2^n > 16384 not supported");
    std::vector<double> Hd((size_t)N1 * N1, 0.0);
    std::vector<char> blk((size_t)NQ * NQ, 0);
    for (int64_t j = 0; j < N1; ++j) {
        if (colptr[j] < 1 || colptr[j + 1] < colptr[j]) return fail(GSG_ERR_ARG, "bad H colptr (expect 1-based)");
        for (int64_t pp = colptr[j] - 1; pp < colptr[j + 1] - 1; ++pp) {
            const int64_t i = rowval[pp] - 1;
            if (i < 0 || i >= N1) return fail(GSG_ERR_ARG, "bad H rowval (expect 1-based)");
            Hd[(size_t)i * N1 + j] += nzval[pp];
            blk[(size_t)(i / K) * NQ + (j / K)] = 1;
        }
    }
    P.KK2 = (K * K + 1) & ~1;
    std::vector<int> rowptr, col;
    std::vector<double> val;
    block_csr_from_dense(Hd, blk, N1, K, NQ, P.KK2, rowptr, col, val);
    P.h_rowptr = rowptr;
    P.h_col = col;
    P.h_val = val;
    GSG_TRY(P.b_rowptr.upload(rowptr));
    GSG_TRY(P.b_col.upload(col));
    GSG_TRY(P.b_val.upload(val));

    // constant-bank kernel (k = 3, N' = 48 and 96): the stored blocks must lie inside the structural
    // pattern the kernel was unrolled for; otherwise that class falls back to the register-tiled kernel
    P.consth_vals.assign(n + 1, {});
    if (K == 3 && !getenv("GSG_NO_CONSTH")) {
        // N' = 96 is unrolled too (134 KB of code) but measured slower than the register-tiled kernel:
        // warps of different CTAs run out of phase and thrash the instruction cache (ncu: no_instruction
        // 4.6 stalls per issue).  Opt-in for experiments only.
        const int pmax_consth = getenv("GSG_CONSTH5") ? 5 : 4;
        for (int p = 4; p <= pmax_consth && p <= n; ++p) {
            const int nq = 1 << p;
            bool inside = true;
            std::vector<double> vals;
            for (int q = 0; q < nq && inside; ++q)
                for (int qc = 0; qc < nq; ++qc) {
                    const bool st = blk[(size_t)q * NQ + qc] != 0;
                    if (pat::touch(q, qc)) {
                        for (int mo = 0; mo < K; ++mo)
                            for (int mi = 0; mi < K; ++mi)
                                vals.push_back(Hd[(size_t)(q * K + mo) * N1 + (qc * K + mi)]);
                    } else if (st) {
                        inside = false;
                        break;
                    }
                }
            if (inside && (int)vals.size() == pat::nblocks(nq) * K * K) P.consth_vals[p] = std::move(vals);
        }
    }

    // dense principal sub-blocks for the register-resident short classes
    std::vector<double> all;
    for (int p = 0; p <= n && p <= SHORT_MAX_P; ++p) {
        const int NP = K << p;
        if (NP > SHORT_MAX_NP) break;
        std::vector<double> d((size_t)NP * NP);
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j) d[(size_t)i * NP + j] = Hd[(size_t)i * N1 + j];
        P.hoff[p] = (int)all.size();
        all.insert(all.end(), d.begin(), d.end());
        if (all.size() & 1) all.push_back(0.0);
        P.short_pmax = p;
    }
    P.htotal = (int)all.size();
    // the kernels' parameter struct HDense<K> always has room for every class up to ShortDims<K>::pmax():
    // with n < pmax the unused classes are zero-filled
    {
        size_t want = 0;
        for (int p = 0; p <= SHORT_MAX_P && (K << p) <= SHORT_MAX_NP; ++p) want += (((size_t)(K << p) * (K << p)) + 1) & ~(size_t)1;
        if (all.size() < want) all.resize(want, 0.0);
    }
    P.dense_host = all;

    // pre-squared blocks of the short classes: (D_op * D_op) restricted to a pole is the square of the pole's
    // principal sub-block (poles are closed under D_op), summed over the inner index in ascending order as
    // SparseArrays' A*B does (src/multidim_derivative.jl:76); no FMA contraction (host_setup.cpp)
    P.sq_pmax = P.short_pmax;
    P.dense_sq_host.assign(all.size(), 0.0);
    {
        std::vector<int> rp, cl;
        std::vector<double> vl;
        for (int p = 0; p <= P.sq_pmax; ++p) {
            const int NP = K << p, nq = 1 << p;
            std::vector<double> a((size_t)NP * NP), sq((size_t)NP * NP);
            for (int i = 0; i < NP; ++i)
                for (int j = 0; j < NP; ++j) a[(size_t)i * NP + j] = Hd[(size_t)i * N1 + j];
            gsg::dense_square(a.data(), NP, sq.data());
            std::copy(sq.begin(), sq.end(), P.dense_sq_host.begin() + P.hoff[p]);
            P.sq_cls_row0[p] = (int)rp.size();
            for (int q = 0; q < nq; ++q) {
                rp.push_back((int)cl.size());
                for (int qc = 0; qc < nq; ++qc) {
                    cl.push_back(qc);
                    const size_t o = vl.size();
                    vl.resize(o + P.KK2, 0.0);
                    for (int mo = 0; mo < K; ++mo)
                        for (int mi = 0; mi < K; ++mi) vl[o + mo * K + mi] = sq[(size_t)(q * K + mo) * NP + (qc * K + mi)];
                }
            }
            rp.push_back((int)cl.size());
        }
        {
            std::vector<int> re(rp.size(), 0);
            for (size_t i = 0; i + 1 < rp.size(); ++i) re[i] = rp[i + 1];
            GSG_TRY(P.sq_rowend.upload(re));
        }
        GSG_TRY(P.sq_rowptr.upload(rp));
        GSG_TRY(P.sq_col.upload(cl));
        GSG_TRY(P.sq_val.upload(vl));
    }
    GSG_TRY(build_rowtile_program(P));
    return 0;
}

// record streams of class p cut into npass column slices (cached per plan)
int get_col_passes(gsg_plan& P, int p, int npass, const std::vector<std::unique_ptr<gsg_plan::ColPass>>** out) {
    auto key = std::make_pair(p, npass);
    auto it = P.lpass.find(key);
    if (it == P.lpass.end()) {
        const int K = P.S.k, KK = K * K;
        const int REC = (KK * 8 + 8 + 15) & ~15;
        const int nq = 1 << p;
        std::vector<std::unique_ptr<gsg_plan::ColPass>> passes;
        for (int ps = 0; ps < npass; ++ps) {
            std::unique_ptr<gsg_plan::ColPass> cp(new gsg_plan::ColPass());
            cp->nqc = nq / npass;
            cp->qc0 = ps * cp->nqc;
            std::vector<unsigned char> buf;
            cp->row_start.assign(nq + 1, 0);
            int nrec = 0;
            for (int q = 0; q < nq; ++q) {
                cp->row_start[q] = nrec;
                std::vector<int> sel;
                for (int b = P.h_rowptr[q]; b < P.h_rowptr[q + 1]; ++b)
                    if (P.h_col[b] >= cp->qc0 && P.h_col[b] < cp->qc0 + cp->nqc) sel.push_back(b);
                const int emit = std::max<int>((int)sel.size(), 1);       // every row owns an end-of-row record
                for (int i = 0; i < emit; ++i) {
                    buf.resize((size_t)(nrec + 1) * REC, 0);
                    unsigned char* rec = buf.data() + (size_t)nrec * REC;
                    int meta[2] = {0, i == emit - 1 ? 1 : 0};
                    if (i < (int)sel.size()) {
                        std::memcpy(rec, P.h_val.data() + (size_t)sel[i] * P.KK2, (size_t)KK * 8);
                        meta[0] = P.h_col[sel[i]] - cp->qc0;
                    }
                    std::memcpy(rec + KK * 8, meta, 8);
                    ++nrec;
                }
            }
            cp->row_start[nq] = nrec;
            buf.resize((size_t)(nrec + 2 * LONG_CH) * REC, 0);     // over-read slack for whole-chunk copies
            GSG_TRY(cp->recs.upload(buf));
            passes.push_back(std::move(cp));
        }
        it = P.lpass.emplace(key, std::move(passes)).first;
    }
    *out = &it->second;
    return 0;
}

// pole groups of direction d (0-based), keyed by the other dims' levels, in layout order of their level_d = 0 block.
// Pure host code (also used by the CPU-side table checks, gsg_debug_flat_tables).
int host_groups(const gsg::IndexSet& S, int d, int part_rank, int part_bits, int exclude_np, std::vector<GroupDev>& groups) {
    const int D = S.D, n = S.n;
    groups.clear();
    for (const gsg::Block& b0 : S.blocks) {
        if (b0.level[d] != 0) continue;
        GroupDev g;
        std::memset(&g, 0, sizeof(g));
        int s = 0;
        for (int i = 0; i < D; ++i) s += b0.level[i];
        g.p = (S.scheme == 1) ? n : n - s;
        std::vector<int> lv = b0.level;
        for (int ld = 0; ld <= g.p; ++ld) {
            lv[d] = ld;
            auto it = S.by_level.find(lv);
            if (it == S.by_level.end()) return fail(GSG_ERR_ARG, "internal: missing block");
            g.base[ld] = S.blocks[it->second].poffset;
        }
        long long Slo = 1, Shi = 1;
        for (int i = 0; i < d; ++i) Slo *= b0.cells[i];
        for (int i = d + 1; i < D; ++i) Shi *= b0.cells[i];
        if (Slo * Shi > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "too many items in a pole group");
        g.S = (int)Slo;
        g.nitems = (int)(Slo * Shi);
        // block partition: this rank sweeps a group iff it owns it.  Along a partition dimension a pole
        // with p >= 1 straddles the pair of ranks that differ in that bit; the rank holding level >= 1
        // (bit 0) sweeps it after receiving the level-0 cells, p == 0 poles stay with the bit-1 rank.
        bool mine = true;
        for (int j = 0; j < part_bits; ++j) {
            const int e = D - 1 - j;
            const int mybit = (part_rank >> j) & 1;
            if (e == d) mine = mine && (g.p == 0 ? mybit == 1 : mybit == 0);
            else mine = mine && ((b0.level[e] == 0) == (mybit == 1));
        }
        if (!mine) continue;
        if (exclude_np >= 0) {            // groups covered by the PAIR tiles of this direction's pair
            const int pa = d & ~1, pb = pa + 1;
            int s_other = 0;
            for (int i = 0; i < D; ++i)
                if (i != pa && i != pb) s_other += b0.level[i];
            if (n - s_other <= exclude_np) continue;
        }
        groups.push_back(g);
    }
    return 0;
}

// flat sweep tables: for every multi-cell (device-layout cell index) and direction d its pole group, item and 1-D cell
int host_flat_cells(const gsg::IndexSet& S, int d, const std::vector<GroupDev>& groups, std::vector<FlatCell>& cells /* ncells * D */) {
    const int D = S.D;
    const long long KDp = S.kDp;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const GroupDev& g = groups[gi];
        for (int ld = 0; ld <= g.p; ++ld) {
            const int Cd = ld <= 1 ? 1 : 1 << (ld - 1);
            const int q0 = ld == 0 ? 0 : 1 << (ld - 1);
            if (g.base[ld] % KDp != 0) return fail(GSG_ERR_UNSUPPORTED, "internal: unaligned block");
            const long long cell0 = g.base[ld] / KDp;
            for (int r = 0; r < g.nitems; ++r) {
                const long long lo = r % g.S, hi = r / g.S;
                for (int cd = 0; cd < Cd; ++cd) {
                    const long long ci = cell0 + lo + (long long)g.S * (cd + (long long)Cd * hi);
                    if (ci < 0 || ci >= S.ncells_total) return fail(GSG_ERR_UNSUPPORTED, "internal: flat cell index out of range");
                    cells[(size_t)ci * D + d] = FlatCell{(int)gi, r, q0 + cd};
                }
            }
        }
    }
    return 0;
}

int build_direction(gsg_plan& P, int d /*0-based*/, Direction& dir, int exclude_np /* -1: keep every group */) {
    const gsg::IndexSet& S = P.S;
    const int K = S.k, n = S.n;
    dir.A = pow_int(K, d);
    const int KD = (int)S.kD, KDp = (int)S.kDp;
    const int PI = KD / K;

    std::vector<GroupDev> groups;
    GSG_TRY(host_groups(S, d, P.part_rank, P.part_bits, exclude_np, groups));
    dir.groups_h = groups;
    GSG_TRY(dir.groups.upload(groups));

    // ---- merged TMA class for all register-resident pole lengths
    const bool use_tma = P.short_pmax >= 0 && K <= 5;
    if (use_tma) {
        SweepClass c;
        c.kind = Kind::SHORT_TMA;
        c.p = P.short_pmax;
        const int cmin = 1 << P.short_pmax;
        int CT = std::max(cmin, (TMA_STAGE_TARGET_DOUBLES / KDp) / cmin * cmin);   // multi-cells per tile
        const size_t fixed = 128 + std::max(8 * sizeof(TileS), 4 * sizeof(TileS2)) + (size_t)PI * 4 + 128;
        int ns = 4;
        if (const char* e = getenv("GSG_STREAM_NS")) ns = std::max(2, std::min(4, atoi(e)));
        while (ns > 2 && fixed + (size_t)ns * CT * KDp * 8 > 200 * 1024) --ns;
        while (CT > cmin && fixed + (size_t)ns * CT * KDp * 8 > 200 * 1024) CT -= cmin;
        if (fixed + (size_t)ns * CT * KDp * 8 <= SMEM_OPTIN_MAX) {
            std::vector<TileS> tl;
            for (int p = P.short_pmax; p >= 0; --p) {          // largest pole length first
                const int nr_max = CT >> p;
                for (size_t gi = 0; gi < groups.size(); ++gi) {
                    if (groups[gi].p != p) continue;
                    for (int r0 = 0; r0 < groups[gi].nitems; r0 += nr_max) {
                        TileS t;
                        std::memset(&t, 0, sizeof(t));
                        for (int l = 0; l <= p && l < 4; ++l) t.base[l] = groups[gi].base[l];
                        t.S = groups[gi].S;
                        t.r0 = r0;
                        t.nr = (short)std::min(nr_max, groups[gi].nitems - r0);
                        t.P = (short)p;
                        tl.push_back(t);
                    }
                }
            }
            c.ntiles = (int)tl.size();
            for (const TileS& t : tl) c.short_dofs += (double)(1 << t.P) * t.nr * KD;
            // second-generation descriptors: ready-made runs of memory-contiguous multi-cells
            c.stream2 = !getenv("GSG_STREAM_V1") && CT <= TILE2_MAXRUN && ns <= 4;
            if (c.stream2) {
                std::vector<TileS2> tl2(tl.size());
                for (size_t ti = 0; ti < tl.size() && c.stream2; ++ti) {
                    const TileS& t = tl[ti];
                    TileS2& o = tl2[ti];
                    std::memset(&o, 0, sizeof(o));
                    o.P = t.P;
                    o.nr = t.nr;
                    o.ncell = (int)t.nr << t.P;
                    int nruns = 0;
                    for (int q = 0; q < (1 << t.P); ++q) {
                        const int ld = q == 0 ? 0 : 32 - __builtin_clz((unsigned)q);
                        const int cd = q == 0 ? 0 : q - (1 << (ld - 1));
                        const int Cd = ld <= 1 ? 1 : 1 << (ld - 1);
                        long long prev = -2;
                        for (int r = 0; r < t.nr; ++r) {
                            const int item = t.r0 + r, lo = item % t.S, hi = item / t.S;
                            const long long ofs = t.base[ld] + (long long)KDp * (lo + (long long)t.S * (cd + (long long)Cd * hi));
                            if (ofs % KDp != 0 || ofs / KDp > 0x7fffffffLL) { c.stream2 = false; break; }
                            const long long cell = ofs / KDp;
                            if (cell == prev + 1 && nruns > 0) {
                                ++o.run[nruns - 1].n;
                            } else {
                                if (nruns == TILE2_MAXRUN) { c.stream2 = false; break; }
                                o.run[nruns].cell0 = (int)cell;
                                o.run[nruns].scell = (short)(q * t.nr + r);
                                o.run[nruns].n = 1;
                                ++nruns;
                            }
                            prev = cell;
                        }
                        if (!c.stream2) break;
                    }
                    o.nruns = nruns;
                }
                if (c.stream2) GSG_TRY(c.s2tiles.upload(tl2));
            }
            if (c.ntiles > 0) {
                GSG_TRY(c.stiles.upload(tl));
                c.sprm.KD = KD;
                c.sprm.KDp = KDp;
                c.sprm.A = dir.A;
                c.sprm.stage_doubles = CT * KDp;
                c.sprm.nstage = ns;
                c.smem = fixed + (size_t)ns * CT * KDp * 8;
                dir.classes.push_back(std::move(c));
            }
        }
    }
    const bool tma_active = !dir.classes.empty();

    std::vector<CellOfs> celltab;                      // filled by the register-tiled long classes
    {
        std::vector<int> offtab(PI);
        for (int j = 0; j < PI; ++j) offtab[j] = (j % dir.A) + K * dir.A * (j / dir.A);
        GSG_TRY(dir.offtab.upload(offtab));
    }

    // ---- row-tile class: every long-pole class p >= rt_pmin in ONE launch, CTA = (item, tile of the class's program)
    const bool rowtile = P.rt_on && tma_active;
    if (rowtile) {
        const std::vector<int> ot = rt_pole_order(K, dir.A, PI, 32 * P.rt_C * P.rt_PW);
        GSG_TRY(dir.rt_offtab.upload(ot));
        SweepClass c;
        c.kind = Kind::ROWTILE;
        c.p = n;
        std::vector<RTWork> work;
        std::vector<int> zero;
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            const GroupDev& g = groups[gi];
            if (g.p < P.rt_pmin) continue;
            const int NQ = 1 << g.p;
            const int ctab = (int)celltab.size();
            const long long KS = (long long)KDp * g.S;
            for (int q = 0; q < NQ; ++q) {
                const int ld = q == 0 ? 0 : 32 - __builtin_clz((unsigned)q);
                const int cd = q == 0 ? 0 : q - (1 << (ld - 1));
                const int Cd = ld <= 1 ? 1 : 1 << (ld - 1);
                celltab.push_back(CellOfs{g.base[ld] + KS * cd, KS * Cd});
            }
            for (int r = 0; r < g.nitems; ++r) {
                const int lo = r % g.S, hi = r / g.S;
                for (int t = 0; t < P.rt_prog.cls_count[g.p]; ++t) {
                    const int ti = P.rt_prog.cls_first[g.p] + t;
                    const RTTile& T = P.rt_prog.tiles[ti];
                    work.push_back(RTWork{ctab, lo, hi, ti, T.nx, T.rec_ofs, T.rec_bytes, 0});
                }
                for (int q : P.rt_prog.cls_partial_q[g.p]) {
                    const CellOfs& co = celltab[ctab + q];
                    zero.push_back((int)((co.bq + (long long)KDp * lo + co.kc * hi) / KDp));
                }
            }
        }
        if (!work.empty()) {
            std::stable_sort(work.begin(), work.end(), [&](const RTWork& a, const RTWork& b) {
                return P.rt_prog.tiles[a.tile].rec_bytes > P.rt_prog.tiles[b.tile].rec_bytes;
            });
            c.ntiles = (int)work.size();
            c.smem = P.rt_smem;
            std::vector<long long> src(work.size() * (size_t)RT_MAXX, 0);
            for (size_t wi = 0; wi < work.size(); ++wi) {
                const RTWork& w = work[wi];
                const RTTile& T = P.rt_prog.tiles[w.tile];
                for (int i = 0; i < T.nx; ++i) {
                    const CellOfs& co = celltab[w.ctab + T.xq[i]];
                    src[wi * RT_MAXX + i] = co.bq + (long long)KDp * w.lo + co.kc * w.hi;
                }
            }
            GSG_TRY(c.rtsrc.upload(src));
            GSG_TRY(c.rtwork.upload(work));
            // persistent CTAs: the items (sorted by cost) are dealt to the least loaded CTA; cost = blocks + a fixed part
            // (measured: ~1.6 us + 0.075 us per block of a tile)
            static const int grid_env = getenv("GSG_RT_GRID") ? atoi(getenv("GSG_RT_GRID")) : -1;
            const int G = grid_env == 0 ? c.ntiles : std::max(1, std::min(c.ntiles, grid_env > 0 ? grid_env : P.sm_count));
            std::vector<std::vector<int>> mine(G);
            std::vector<long long> load(G, 0);
            for (size_t wi = 0; wi < work.size(); ++wi) {
                int b = 0;
                for (int h = 1; h < G; ++h)
                    if (load[h] < load[b]) b = h;
                mine[b].push_back((int)wi);
                load[b] += P.rt_prog.tiles[work[wi].tile].rec_bytes / (K * K * 8) + 22;
            }
            size_t kmax = 0;
            for (const auto& m : mine) kmax = std::max(kmax, m.size());
            std::vector<int> sched((kmax + 1) * (size_t)G, -1);
            for (int b = 0; b < G; ++b)
                for (size_t k2 = 0; k2 < mine[b].size(); ++k2) sched[k2 * G + b] = mine[b][k2];
            GSG_TRY(c.rtsched.upload(sched));
            c.rtgrid = G;
            GSG_TRY(c.rtzero.upload(zero));
            dir.classes.push_back(std::move(c));
        }
    }

    for (int p = 0; p <= n; ++p) {
        SweepClass c;
        c.p = p;
        const int NQ = 1 << p, NP = K * NQ;
        const int REC = (K * K * 8 + 8 + 15) & ~15;
        if (short_supported(K, p) && tma_active) continue;        // served by the streaming class above
        if (rowtile && p >= P.rt_pmin) continue;                  // served by the row-tile class above
        // register-tiled / constant-bank kernels for the long classes of k <= 5; everything else (k > 5, short
        // classes whose stage does not fit shared memory) goes to the generic block-CSR kernel
        c.kind = (K <= 5 && !short_supported(K, p)) ? Kind::LONG : Kind::GENERIC;
        if (c.kind == Kind::LONG && K == 3 && (p == 4 || p == 5) && !P.consth_vals[p].empty()) {
            // constant-bank kernel: one warp per 32 consecutive (flattened) poles of one group
            std::vector<TileL2> tl2;
            for (size_t gi = 0; gi < groups.size(); ++gi) {
                if (groups[gi].p != p) continue;
                const GroupDev& g = groups[gi];
                const long long np = (long long)g.nitems * PI;
                if (np > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "too many poles in a pole group");
                const int ctab = (int)celltab.size();
                for (int q = 0; q < NQ; ++q) {
                    const int ld = q == 0 ? 0 : 32 - __builtin_clz((unsigned)q);
                    const int cd = q == 0 ? 0 : q - (1 << (ld - 1));
                    const int Cd = ld <= 1 ? 1 : 1 << (ld - 1);
                    const long long KS = (long long)KDp * g.S;
                    celltab.push_back(CellOfs{g.base[ld] + KS * cd, KS * Cd});
                }
                for (long long p0 = 0; p0 < np; p0 += 32) {
                    TileL2 t;
                    t.ctab = ctab;
                    t.S = g.S;
                    t.item0 = (int)(p0 / PI);
                    t.j0 = (int)(p0 % PI);
                    t.lo0 = t.item0 % g.S;
                    t.hi0 = t.item0 / g.S;
                    t.npoles = (int)std::min<long long>(32, np - p0);
                    t.part = 0;
                    tl2.push_back(t);
                }
            }
            if (tl2.empty()) continue;
            c.kind = Kind::CONSTH;
            c.ntiles = (int)tl2.size();
            c.smem = (size_t)CONSTH_WARPS * ((size_t)NP * 32 * 8 + (size_t)NQ * sizeof(CellOfs));
            GSG_TRY(c.l2tiles.upload(tl2));
            dir.classes.push_back(std::move(c));
            continue;
        }
        if (c.kind == Kind::LONG) {
            // register-tiled long kernel: tiles of 32*C consecutive (flattened) poles of one group
            long long maxpoles = 0, ngrp = 0;
            for (size_t gi = 0; gi < groups.size(); ++gi)
                if (groups[gi].p == p) { maxpoles = std::max(maxpoles, (long long)groups[gi].nitems * PI); ++ngrp; }
            if (ngrp == 0) continue;
            int C = K <= 3 ? 4 : 2;
            if (const char* e = getenv("GSG_LONG_C")) C = atoi(e);
            while (C > 1 && 32 * (C >> 1) >= maxpoles) C >>= 1;                   // no wider than the groups
            // the x tile of one column pass (NP / npass rows x 32*C poles) must fit shared memory: cut the
            // matrix into column slices rather than narrow the tile (shared-memory wavefronts per DFMA fall
            // with C; a pass costs one more read-modify-write of the class's y, which is small)
            size_t xcap = 200 * 1024;
            if (const char* e = getenv("GSG_LONG_XCAP")) xcap = (size_t)atoi(e) * 1024;
            int npass = 1;
            while (npass < NQ && (size_t)(NP / npass) * 32 * C * 8 > xcap) npass <<= 1;
            // Column passes are opt-in (GSG_LONG_PASSES=1): measured slower end to end at D=6 (the passes of
            // one class run back to back on a handful of SMs and stretch the sweep's tail); by default the
            // tile is narrowed instead.
            if (!getenv("GSG_LONG_PASSES")) {
                npass = 1;
                while (C > 1 && (size_t)NP * 32 * C * 8 > xcap) C >>= 1;
            }
            // experiment: where C = 4 needs a tile above 100 KB (one 8-warp CTA per SM), take C = 2 with a 2-deep
            // ring instead so that two CTAs (16 warps) share the SM
            bool two_per_sm = false;
            if (getenv("GSG_LONG_HALF") && K <= 3 && C == 4 && (size_t)(NP / npass) * 32 * C * 8 > 100 * 1024 &&
                (size_t)(NP / npass) * 32 * 2 * 8 <= 100 * 1024) {
                C = 2;
                two_per_sm = true;
            }
            const size_t xtile = (size_t)(NP / npass) * 32 * C * 8;
            // ~190 registers per thread at C = 4: small CTAs (several per SM) while the tile is small
            int nw2 = (C >= 4 && xtile <= 50 * 1024) ? 4 : 8;
            // C <= 2 (<= 128 registers): a CTA with a big x tile owns its SM, and 8 warps leave the record loop
            // latency-bound (~2.7x its shared-memory bound at N' = 384) -- run 16 warps with a 2-deep ring
            int nbuf = LONG_NBUF;
            if (K <= 3 && C <= 2 && xtile > 100 * 1024 && NQ >= 64 && !getenv("GSG_LONG_NO16")) { nw2 = 16; nbuf = 2; }
            if (two_per_sm) { nw2 = 8; nbuf = 2; }
            if (const char* e = getenv("GSG_LONG_NW")) nw2 = std::min(atoi(e), nbuf == 2 ? 16 : 8);
            nw2 = std::max(1, std::min(nw2, nbuf == 2 ? 16 : 8));
            const size_t ring_bytes = (size_t)nbuf * LONG_CH * REC;
            while (nw2 > 2 && xtile + nw2 * ring_bytes + 2048 > SMEM_OPTIN_MAX) nw2 >>= 1;
            const size_t smem2 = xtile + nw2 * ring_bytes;
            c.nbuf = nbuf;
            if (smem2 + 2048 <= SMEM_OPTIN_MAX && (C == 1 || C == 2 || C == 4)) {
                c.kind = Kind::LONG2;
                c.cpl = C;
                c.nwarps = nw2;
                c.smem = smem2;
                std::vector<TileL2> base;
                const int PT = 32 * C;
                for (size_t gi = 0; gi < groups.size(); ++gi) {
                    if (groups[gi].p != p) continue;
                    const GroupDev& g = groups[gi];
                    const long long np = (long long)g.nitems * PI;
                    if (np > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "too many poles in a pole group");
                    // the group's cell table
                    const int ctab = (int)celltab.size();
                    for (int q = 0; q < NQ; ++q) {
                        const int ld = q == 0 ? 0 : 32 - __builtin_clz((unsigned)q);
                        const int cd = q == 0 ? 0 : q - (1 << (ld - 1));
                        const int Cd = ld <= 1 ? 1 : 1 << (ld - 1);
                        const long long KS = (long long)KDp * g.S;
                        celltab.push_back(CellOfs{g.base[ld] + KS * cd, KS * Cd});
                    }
                    for (long long p0 = 0; p0 < np; p0 += PT) {
                        TileL2 t;
                        t.ctab = ctab;
                        t.S = g.S;
                        t.item0 = (int)(p0 / PI);
                        t.j0 = (int)(p0 % PI);
                        t.lo0 = t.item0 % g.S;
                        t.hi0 = t.item0 / g.S;
                        t.npoles = (int)std::min<long long>(PT, np - p0);
                        t.part = 0;
                        base.push_back(t);
                    }
                }
                // row parts: the x tile is staged once per part, so as few parts as keep one CTA's record
                // stream short enough to finish well inside the sweep (the long classes run beside the
                // streaming kernel; their SM time, not their latency, is what counts)
                const std::vector<std::unique_ptr<gsg_plan::ColPass>>* passes = nullptr;
                GSG_TRY(get_col_passes(P, p, npass, &passes));
                c.npass = npass;
                int maxrec = 0;
                for (const auto& cp : *passes) maxrec = std::max(maxrec, cp->row_start[NQ]);
                const int reccap = std::max(128, 1400 / npass);       // passes run back to back: keep each short
                int rsplit = (maxrec + reccap - 1) / reccap;
                if (const char* e = getenv("GSG_LONG_RSPLIT")) rsplit = atoi(e);
                rsplit = std::max(1, std::min(rsplit, std::max(1, NQ / (2 * nw2))));
                c.rsplit = rsplit;
                const int G = rsplit * nw2;
                for (const auto& cp : *passes) {
                    const std::vector<int>& rs = cp->row_start;
                    const int nrec = rs[NQ];
                    std::vector<int> pb(G + 1, nrec), pr(G + 1, NQ);
                    pb[0] = 0; pr[0] = 0;
                    const long long total = (long long)nrec + NQ;       // +1 per row: epilogue cost
                    int q = 0;
                    for (int g = 1; g < G; ++g) {
                        const long long target = total * g / G;
                        while (q < NQ && (long long)rs[q] + q < target) ++q;
                        pr[g] = std::max(q, pr[g - 1]);
                        pb[g] = rs[pr[g]];
                    }
                    c.passBlk.emplace_back(new DevBuf<int>());
                    c.passRow.emplace_back(new DevBuf<int>());
                    GSG_TRY(c.passBlk.back()->upload(pb));
                    GSG_TRY(c.passRow.back()->upload(pr));
                }
                std::vector<TileL2> full;
                full.reserve(base.size() * rsplit);
                for (int part = 0; part < rsplit; ++part)
                    for (TileL2 t1 : base) { t1.part = part; full.push_back(t1); }
                c.ntiles = (int)full.size();
                GSG_TRY(c.l2tiles.upload(full));
                dir.classes.push_back(std::move(c));
                continue;
            }
        }
        c.kind = Kind::GENERIC;               // no register-tiled configuration fits: generic kernel
        std::vector<TileDev> tl;
        {
            // largest power-of-K pole sub-range that fits the budget
            int npole = 1;
            if (K >= 2)
                while (npole * K <= PI && (size_t)16 * NP * npole * K <= GENERIC_SMEM_BUDGET) npole *= K;
            if ((size_t)16 * NP * npole > 200 * 1024) return fail(GSG_ERR_UNSUPPORTED, "pole too long for shared memory");
            int nr_max = 1;
            if (npole == PI) {
                nr_max = (int)std::max<size_t>(1, GENERIC_SMEM_BUDGET / ((size_t)16 * NP * npole));
                nr_max = std::min(nr_max, std::max(1, 512 / npole));
            }
            c.NPOLE = npole;
            c.Amin = std::min(dir.A, npole);
            const int nsub = PI / npole;
            for (size_t gi = 0; gi < groups.size(); ++gi) {
                if (groups[gi].p != p) continue;
                for (int r0 = 0; r0 < groups[gi].nitems; r0 += nr_max) {
                    const int nr = std::min(nr_max, groups[gi].nitems - r0);
                    for (int sr = 0; sr < nsub; ++sr) {
                        int ebase;
                        if (dir.A >= npole) {
                            const int chunks = dir.A / npole;
                            ebase = (sr % chunks) * npole + K * dir.A * (sr / chunks);
                        } else {
                            ebase = K * dir.A * (sr * (npole / dir.A));
                        }
                        tl.push_back(TileDev{(int)gi, r0, nr, ebase});
                    }
                }
            }
            c.smem = (size_t)16 * NP * npole * nr_max + (size_t)2 * K * npole * sizeof(int) + 16;
        }
        c.ntiles = (int)tl.size();
        if (c.ntiles == 0) continue;
        GSG_TRY(c.tiles.upload(tl));
        dir.classes.push_back(std::move(c));
    }
    GSG_TRY(dir.celltab.upload(celltab));
    // launch order: the persistent streaming kernel goes on the main stream (index 0); the long
    // classes follow, longest poles first, each on its own forked stream
    std::stable_sort(dir.classes.begin(), dir.classes.end(), [](const SweepClass& a, const SweepClass& b) {
        const int ka = a.kind == Kind::SHORT_TMA ? 1 : 0, kb = b.kind == Kind::SHORT_TMA ? 1 : 0;
        if (ka != kb) return ka > kb;
        return a.p > b.p;
    });
    return 0;
}

// PAIR tiles (direction-pair fusion of the streaming classes, kernels.cuh): whole 2-D sub-planes with n' <= pair_np
int build_pairs(gsg_plan& P) {
    const gsg::IndexSet& S = P.S;
    const int D = S.D, K = S.k, n = S.n;
    P.pair_np = -1;
    P.pairs.clear();
    P.dirs_red.clear();
    if (getenv("GSG_NO_PAIR") || S.scheme != 0 || D < 2 || K > 5 || P.short_pmax < 0) return 0;
    int npmax = K <= 3 ? 2 : (K == 4 ? 1 : 0);
    npmax = std::min(npmax, std::min(P.short_pmax, n));
    if (const char* e = getenv("GSG_PAIR_NP")) npmax = std::min(npmax, atoi(e));
    if (npmax < 0) return 0;
    const int KD = (int)S.kD, KDp = (int)S.kDp;
    const int CT = 8;                                             // multi-cells per ring stage
    if ((long long)CT * KDp * 8 * 2 > 200 * 1024) return 0;
    const int npairs = D / 2;
    P.pairs.resize(npairs);
    for (int j = 0; j < npairs; ++j) {
        const int da = 2 * j, db = da + 1;
        SweepClass& c = P.pairs[j];
        c.kind = Kind::SHORT_TMA;
        c.stream2 = true;
        std::vector<TileS2> tl;
        // block partition: a pair that contains a partition dimension is not fused (its poles straddle ranks);
        // otherwise a sub-plane belongs to the rank that owns its blocks (ownership depends on the other levels)
        bool pair_has_partition_dim = false;
        for (int jb = 0; jb < P.part_bits; ++jb) {
            const int e = D - 1 - jb;
            if (e == da || e == db) pair_has_partition_dim = true;
        }
        if (pair_has_partition_dim) { c.ntiles = 0; continue; }
        for (int np = npmax; np >= 0; --np) {                     // largest sub-planes first
            const int NC = pairp::ncell(np), nr_max = CT / NC;
            // plane groups: levels of the other dims with n - sum == np; representative = the (0, 0) block
            for (const gsg::Block& b0 : S.blocks) {
                if (b0.level[da] != 0 || b0.level[db] != 0) continue;
                int s_other = 0;
                for (int i = 0; i < D; ++i) s_other += b0.level[i];
                if (n - s_other != np) continue;
                bool mine = true;
                for (int jb = 0; jb < P.part_bits; ++jb) {
                    const int e = D - 1 - jb;
                    mine = mine && ((b0.level[e] == 0) == (((P.part_rank >> jb) & 1) == 1));
                }
                if (!mine) continue;
                long long nitems = 1;
                for (int i = 0; i < D; ++i)
                    if (i != da && i != db) nitems *= b0.cells[i];
                // per slot: block and the cell strides of the other dims inside it
                std::vector<const gsg::Block*> sblk(NC);
                std::vector<long long> sfix(NC);                  // cell offset contributed by (c_a, c_b)
                for (int qb = 0; qb < (1 << np); ++qb)
                    for (int qa = 0; qa < (1 << np); ++qa) {
                        const int sidx = pairp::slot(np, qa, qb);
                        if (sidx < 0) continue;
                        std::vector<int> lv = b0.level;
                        const int la = pairp::lvl(qa), lb = pairp::lvl(qb);
                        lv[da] = la; lv[db] = lb;
                        auto it = S.by_level.find(lv);
                        if (it == S.by_level.end()) return fail(GSG_ERR_ARG, "internal: missing block (pair)");
                        const gsg::Block& bk = S.blocks[it->second];
                        const int ca = la <= 1 ? 0 : qa - (1 << (la - 1)), cb = lb <= 1 ? 0 : qb - (1 << (lb - 1));
                        long long stride = 1, fix = 0;
                        for (int i = 0; i < D; ++i) {
                            if (i == da) fix += ca * stride;
                            if (i == db) fix += cb * stride;
                            stride *= bk.cells[i];
                        }
                        sblk[sidx] = &bk;
                        sfix[sidx] = fix;
                    }
                for (long long r0 = 0; r0 < nitems; r0 += nr_max) {
                    const int nr = (int)std::min<long long>(nr_max, nitems - r0);
                    TileS2 o;
                    std::memset(&o, 0, sizeof(o));
                    o.P = (short)np;
                    o.nr = (short)nr;
                    o.ncell = NC * nr;
                    int nruns = 0;
                    for (int sidx = 0; sidx < NC; ++sidx) {
                        const gsg::Block& bk = *sblk[sidx];
                        long long prev = -2;
                        for (int r = 0; r < nr; ++r) {
                            // item index -> cells of the other dims (first other dim fastest) -> offset in this block
                            long long rem = r0 + r, stride = 1, lin = sfix[sidx];
                            for (int i = 0; i < D; ++i) {
                                if (i != da && i != db) {
                                    const long long ci = rem % b0.cells[i];
                                    rem /= b0.cells[i];
                                    lin += ci * stride;
                                }
                                stride *= bk.cells[i];
                            }
                            if (bk.poffset % KDp != 0) return fail(GSG_ERR_UNSUPPORTED, "internal: unaligned block");
                            const long long cell = bk.poffset / KDp + lin;
                            if (cell > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "too many multi-cells");
                            if (cell == prev + 1 && nruns > 0) {
                                ++o.run[nruns - 1].n;
                            } else {
                                if (nruns == TILE2_MAXRUN) return fail(GSG_ERR_UNSUPPORTED, "internal: too many runs");
                                o.run[nruns].cell0 = (int)cell;
                                o.run[nruns].scell = (short)(sidx * nr + r);
                                o.run[nruns].n = 1;
                                ++nruns;
                            }
                            prev = cell;
                        }
                    }
                    o.nruns = nruns;
                    tl.push_back(o);
                    c.short_dofs += (double)NC * nr * KD;
                }
            }
        }
        c.ntiles = (int)tl.size();
        if (c.ntiles == 0) {
            if (P.part_bits == 0) { P.pairs.clear(); return 0; }
            continue;                       // this rank owns no sub-plane of the pair
        }
        GSG_TRY(c.s2tiles.upload(tl));
        const int PI = KD / K;
        const size_t fixed = 128 + std::max(8 * sizeof(TileS), 4 * sizeof(TileS2)) + (size_t)PI * 4 + 128;
        int ns = 4;
        while (ns > 2 && fixed + (size_t)ns * CT * KDp * 8 > 200 * 1024) --ns;
        c.sprm.KD = KD;
        c.sprm.KDp = KDp;
        c.sprm.A = pow_int(K, da);
        c.sprm.stage_doubles = CT * KDp;
        c.sprm.nstage = ns;
        c.smem = fixed + (size_t)ns * CT * KDp * 8;
        c.p = npmax;
    }
    P.pair_np = npmax;
    P.dirs_red.resize(2 * npairs);
    for (int d = 0; d < 2 * npairs; ++d) {
        bool fused = true;                  // pairs holding a partition dimension keep their complete tables
        for (int jb = 0; jb < P.part_bits; ++jb) {
            const int e = D - 1 - jb;
            if (e == (d & ~1) || e == (d | 1)) fused = false;
        }
        GSG_TRY(build_direction(P, d, P.dirs_red[d], fused ? npmax : -1));
    }
    return 0;
}

// tile range of a launch (every launch covers its whole tile list)
inline void tile_range(const gsg_plan&, int ntiles, int& begin, int& count) {
    begin = 0;
    count = ntiles;
}

int launch_check(const char* what, int K, const SweepClass& c) {
    cudaError_t e = cudaGetLastError();
    static const bool debug_sync = getenv("GSG_DEBUG_SYNC") != nullptr;
    if (e == cudaSuccess && debug_sync) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s<K=%d> p=%d ntiles=%d smem=%zu NPOLE=%d Amin=%d: %s", what, K, c.p, c.ntiles,
                 c.smem, c.NPOLE, c.Amin, cudaGetErrorString(e));
        return fail(GSG_ERR_CUDA, buf);
    }
    return 0;
}

// Kernel attributes are per device: the per-kernel `configured` cache passed in by the launchers only
// remembers the largest size set on the device that set it, so it is keyed by the current device here
// (one process may hold plans on several GPUs).
template <class Kern>
int ensure_smem(Kern kern, size_t smem, size_t& configured_unused) {
    (void)configured_unused;
    // process-wide (the attribute is): a per-thread cache let a second host thread LOWER the limit a first
    // thread had raised, and the first thread's next large launch failed with "invalid argument"
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    GSG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& configured = done[std::make_pair(reinterpret_cast<const void*>(kern), dev)];
    if (configured == 0) {
        // every sweep kernel asks for the maximum shared-memory carve-out: kernels with different
        // carve-outs cannot be resident on one SM at the same time
        GSG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = 1;
    }
    if (smem > 32 * 1024 && smem > configured) {      // static + dynamic must stay under 48 KB without opt-in
        GSG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    return 0;
}

template <int K>
int launch_short_tma(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                     double* y, double alpha, double beta) {
    if constexpr (K >= 1 && K <= 5) {
        auto kern = sweep_short_tma_kernel<K>;
        static thread_local size_t configured = 0;
        GSG_TRY(ensure_smem(kern, c.smem, configured));
        int tb, tn;
        tile_range(pl, c.ntiles, tb, tn);
        if (tn == 0) return 0;
        // one CTA per SM; a launch with little work (a rank's share of a partitioned plan) takes fewer CTAs, >= 12 tiles
        // each, so that the independent launches of a right-hand side run side by side instead of queueing for SMs
        static const int sgrid_env = getenv("GSG_STREAM_GRID") ? atoi(getenv("GSG_STREAM_GRID")) : 0;   // SM split experiment
        const int grid = std::max(1, std::min(sgrid_env > 0 ? sgrid_env : pl.sm_count, std::min(tn, std::max(8, tn / 12))));
        static_assert(sizeof(HDense<K>) + 256 < 32000, "dense blocks must fit the kernel parameter space");
        HDense<K> hd;
        const std::vector<double>& dense = pl.use_sq ? pl.dense_sq_host : pl.dense_host;     // Laplacian: pre-squared blocks
        if ((int)dense.size() != ShortDims<K>::htotal())
            return fail(GSG_ERR_UNSUPPORTED, "internal: dense block table size mismatch");
        std::memcpy(hd.v, dense.data(), sizeof(hd.v));
        int* const counter = pl.tile_counter.p + (pl.ctr_next++ & 63);       // launches may overlap: one slot each
        GSG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
        const bool prof = pl.prof_on && pl.prof_used < std::min(pl.prof_cap, pl.prof_ev.size());
        if (prof) GSG_CUDA(cudaEventRecord(pl.prof_ev[pl.prof_used].first, st));
        if (c.stream2) {
            auto kern2 = sweep_stream_kernel<K, false>;
            static thread_local size_t configured2 = 0;
            GSG_TRY(ensure_smem(kern2, c.smem, configured2));
            kern2<<<grid, STREAM_THREADS, c.smem, st>>>(x, y, alpha, 0.0, beta != 0.0 ? 1 : 0, c.s2tiles.p + tb, tn, hd,
                                                         c.sprm, counter, pl.dbg);
        } else
        kern<<<grid, 32 * (SHORT_TMA_COMPUTE_WARPS + 1), c.smem, st>>>(x, y, alpha, beta != 0.0 ? 1 : 0, dir.groups.p,
                                                                         c.stiles.p + tb, tn, hd, c.sprm,
                                                                         counter, pl.dbg);
        if (prof) {
            GSG_CUDA(cudaEventRecord(pl.prof_ev[pl.prof_used].second, st));
            ++pl.prof_used;
            pl.prof_dofs += c.short_dofs * (double)tn / (double)c.ntiles;
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return launch_check("sweep_short_tma", K, c);
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: TMA short kernel not instantiated");
}

// PAIR tiles of pair j: y = alpha_a D_a x + alpha_b D_b x (+ y) on the sub-planes they cover
template <int K>
int launch_pair(gsg_plan& pl, cudaStream_t st, int j, const double* x, double* y, double alpha_a, double alpha_b,
                double beta) {
    if constexpr (K >= 1 && K <= 5) {
        const SweepClass& c = pl.pairs[j];
        auto kern = sweep_stream_kernel<K, true>;
        static thread_local size_t configured = 0;
        GSG_TRY(ensure_smem(kern, c.smem, configured));
        static const int sgrid_env = getenv("GSG_STREAM_GRID") ? atoi(getenv("GSG_STREAM_GRID")) : 0;
        const int grid = std::max(1, std::min(sgrid_env > 0 ? sgrid_env : pl.sm_count, std::min(c.ntiles, std::max(8, c.ntiles / 12))));
        HDense<K> hd;
        if ((int)pl.dense_host.size() != ShortDims<K>::htotal())
            return fail(GSG_ERR_UNSUPPORTED, "internal: dense block table size mismatch");
        std::memcpy(hd.v, pl.dense_host.data(), sizeof(hd.v));
        int* const counter = pl.tile_counter.p + (pl.ctr_next++ & 63);
        GSG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
        const bool prof = pl.prof_on && pl.prof_used < std::min(pl.prof_cap, pl.prof_ev.size());
        if (prof) GSG_CUDA(cudaEventRecord(pl.prof_ev[pl.prof_used].first, st));
        kern<<<grid, StreamCfg<true>::THREADS, c.smem, st>>>(x, y, alpha_a, alpha_b, beta != 0.0 ? 1 : 0, c.s2tiles.p, c.ntiles, hd,
                                                    c.sprm, counter, nullptr);
        if (prof) {
            GSG_CUDA(cudaEventRecord(pl.prof_ev[pl.prof_used].second, st));
            ++pl.prof_used;
            pl.prof_dofs += 2.0 * c.short_dofs;        // two directional applies' worth of DOFs in one pass
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return launch_check("sweep_stream(pair)", K, c);
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: pair kernel not instantiated");
}

template <int K, int C, int NB>
int launch_long2_kc(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                    double* y, double alpha, double beta) {
    auto kern = sweep_long2_kernel<K, C, NB>;
    static thread_local size_t configured = 0;
    GSG_TRY(ensure_smem(kern, c.smem, configured));
    int tb, tn;
    tile_range(pl, c.ntiles, tb, tn);
    if (tn == 0) return 0;
    const int PI = (int)pl.S.kD / K;
    auto it = pl.lpass.find(std::make_pair(c.p, c.npass));
    if (it == pl.lpass.end()) return fail(GSG_ERR_UNSUPPORTED, "internal: column passes missing");
    for (int ps = 0; ps < c.npass; ++ps) {        // passes accumulate into the same y: same stream, in order
        const gsg_plan::ColPass& cp = *it->second[ps];
        kern<<<tn, c.nwarps * 32, c.smem, st>>>(x, y, alpha, (beta != 0.0 || ps > 0) ? 1 : 0, dir.celltab.p,
                                                 dir.offtab.p, c.l2tiles.p + tb, cp.recs.p, c.passBlk[ps]->p,
                                                 c.passRow[ps]->p, c.p, cp.qc0, cp.nqc, (int)pl.S.kDp, dir.A, PI,
                                                 pl.dbg);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return launch_check("sweep_long2", K, c);
}

template <int K>
int launch_long2_k(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                   double* y, double alpha, double beta) {
    if constexpr (K >= 1 && K <= 5) {
        if (c.nbuf == 2) {
            if constexpr (K <= 3) {
                switch (c.cpl) {
                    case 1: return launch_long2_kc<K, 1, 2>(pl, st, dir, c, x, y, alpha, beta);
                    case 2: return launch_long2_kc<K, 2, 2>(pl, st, dir, c, x, y, alpha, beta);
                }
            }
            return fail(GSG_ERR_UNSUPPORTED, "internal: 16-warp long kernel not instantiated");
        }
        switch (c.cpl) {
            case 1: return launch_long2_kc<K, 1, 4>(pl, st, dir, c, x, y, alpha, beta);
            case 2: return launch_long2_kc<K, 2, 4>(pl, st, dir, c, x, y, alpha, beta);
            case 4: if constexpr (K <= 3) return launch_long2_kc<K, 4, 4>(pl, st, dir, c, x, y, alpha, beta); break;
        }
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: register-tiled long kernel not instantiated");
}

template <int K, int PP>
int launch_consth_kp(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                     double* y, double alpha, double beta) {
    auto kern = sweep_consth_kernel<K, PP>;
    static thread_local size_t configured = 0;
    GSG_TRY(ensure_smem(kern, c.smem, configured));
    int tb, tn;
    tile_range(pl, c.ntiles, tb, tn);
    if (tn == 0) return 0;
    static_assert(sizeof(HBlocks<K, PP>) + 128 < 32764, "pattern blocks must fit the kernel parameter space");
    HBlocks<K, PP> hb;
    const std::vector<double>& vals = pl.consth_vals[PP];
    if (vals.size() * sizeof(double) != sizeof(hb.v)) return fail(GSG_ERR_UNSUPPORTED, "internal: pattern block table size mismatch");
    std::memcpy(hb.v, vals.data(), sizeof(hb.v));
    const int PI = (int)pl.S.kD / K;
    const int grid = (tn + CONSTH_WARPS - 1) / CONSTH_WARPS;
    kern<<<grid, 32 * CONSTH_WARPS, c.smem, st>>>(x, y, alpha, beta != 0.0 ? 1 : 0, dir.celltab.p, dir.offtab.p,
                                                   c.l2tiles.p + tb, tn, (int)pl.S.kDp, dir.A, PI, hb, pl.dbg);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_check("sweep_consth", K, c);
}

int launch_consth(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                  double* y, double alpha, double beta) {
    if (pl.S.k == 3 && c.p == 4) return launch_consth_kp<3, 4>(pl, st, dir, c, x, y, alpha, beta);
    if (pl.S.k == 3 && c.p == 5) return launch_consth_kp<3, 5>(pl, st, dir, c, x, y, alpha, beta);
    return fail(GSG_ERR_UNSUPPORTED, "internal: constant-bank kernel not instantiated");
}

template <int K>
int launch_generic_k(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                     double* y, double alpha, double beta) {
    auto kern = sweep_generic_kernel<K>;
    static thread_local size_t configured = 0;
    GSG_TRY(ensure_smem(kern, c.smem, configured));
    Bcsr M{pl.b_rowptr.p, pl.b_col.p, pl.b_val.p, pl.KK2};
    if (pl.use_sq) {                       // Laplacian: the class's own pre-squared rows
        if (c.p > pl.sq_pmax) return fail(GSG_ERR_UNSUPPORTED, "internal: no squared blocks for this class");
        M = Bcsr{pl.sq_rowptr.p + pl.sq_cls_row0[c.p], pl.sq_col.p, pl.sq_val.p, pl.KK2};
    }
    int tb, tn;
    tile_range(pl, c.ntiles, tb, tn);
    if (tn == 0) return 0;
    kern<<<tn, 256, c.smem, st>>>(x, y, alpha, beta, dir.groups.p, c.tiles.p + tb, M, pl.S.k, c.p, (int)pl.S.kDp,
                                   dir.A, c.NPOLE, c.Amin);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_check("sweep_generic", K, c);
}

template <int K>
int launch_rowtile_k(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                     double* y, double alpha, double beta) {
    if constexpr (K >= 1 && K <= 5) {
        if (c.ntiles == 0) return 0;
        const int KDp = (int)pl.S.kDp;
        if (beta == 0.0 && c.rtzero.n > 0) {      // rows that only receive partial sums start from zero
            const int nz = (int)c.rtzero.n;
            zero_cells_kernel<<<std::min(nz, pl.sm_count * 8), 128, 0, st>>>(y, c.rtzero.p, nz, KDp);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        const int threads = 32 * pl.rt_PW * pl.rt_RG;
        auto go = [&](auto kern) -> int {
            static thread_local size_t configured = 0;
            GSG_TRY(ensure_smem(kern, c.smem, configured));
            // persistent: one CTA per SM walks its list of work items (GSG_RT_GRID=0 at plan creation: one CTA per item)
            kern<<<c.rtgrid, threads, c.smem, st>>>(x, y, alpha, beta != 0.0 ? 1 : 0, dir.celltab.p, dir.rt_offtab.p, c.rtwork.p,
                                                     c.rtsched.p, c.rtsrc.p, pl.rt_tiles.p, pl.rt_groups.p, pl.rt_recs.p, KDp, dir.A,
                                                     pl.rt_PW, pl.rt_RG, pl.dbg);
            return 0;
        };
        if (threads > 256) return fail(GSG_ERR_UNSUPPORTED, "internal: row-tile kernel launched with more than 8 warps");
        if (pl.rt_C == 4) { if constexpr (K <= 3) GSG_TRY(go(sweep_rowtile_kernel<K, 4, 256>)); else return fail(GSG_ERR_UNSUPPORTED, "internal: row-tile C = 4 for k > 3"); }
        else if (pl.rt_C == 2) { if constexpr (K <= 4) GSG_TRY(go(sweep_rowtile_kernel<K, 2, 256>)); else return fail(GSG_ERR_UNSUPPORTED, "internal: row-tile C = 2 for k > 4"); }
        else GSG_TRY(go(sweep_rowtile_kernel<K, 1, 256>));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return launch_check("sweep_rowtile", K, c);
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: row-tile kernel not instantiated");
}

#define GSG_K_SWITCH(FN, ...)                       \
    switch (K) {                                    \
        case 1: return FN<1>(__VA_ARGS__);          \
        case 2: return FN<2>(__VA_ARGS__);          \
        case 3: return FN<3>(__VA_ARGS__);          \
        case 4: return FN<4>(__VA_ARGS__);          \
        case 5: return FN<5>(__VA_ARGS__);          \
        default: break;                             \
    }

int launch_class(gsg_plan& pl, cudaStream_t st, const Direction& dir, const SweepClass& c, const double* x,
                 double* y, double alpha, double beta) {
    const int K = pl.S.k;
    switch (c.kind) {
        case Kind::SHORT_TMA: GSG_K_SWITCH(launch_short_tma, pl, st, dir, c, x, y, alpha, beta); break;
        case Kind::LONG2: GSG_K_SWITCH(launch_long2_k, pl, st, dir, c, x, y, alpha, beta); break;
        case Kind::CONSTH: return launch_consth(pl, st, dir, c, x, y, alpha, beta);
        case Kind::ROWTILE: GSG_K_SWITCH(launch_rowtile_k, pl, st, dir, c, x, y, alpha, beta); break;
        case Kind::GENERIC:
            GSG_K_SWITCH(launch_generic_k, pl, st, dir, c, x, y, alpha, beta);
            return launch_generic_k<0>(pl, st, dir, c, x, y, alpha, beta);
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: no kernel for this class");
}

int elementwise_grid(const gsg_plan& pl, int64_t N) {
    const int64_t want = (N + 255) / 256;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)pl.sm_count * 16));
}

// ---- flat path: one launch per right-hand side for small index sets (kernels.cuh, sweep_flat_kernel) ----------
constexpr int64_t FLAT_AUTO_MAX = 6000000;       // automatic mode: N * D up to this (launch-latency regime)

bool flat_supported(const gsg_plan& pl) {
    return pl.S.k >= 1 && pl.S.k <= 5 && pl.S.D <= FLAT_MAXD && pl.part_bits == 0 && pl.S.ncells_total < 0x7fffffffLL &&
           (size_t)FLAT_WARPS * pl.S.kDp * sizeof(double) + (size_t)pl.S.D * pl.S.kD * sizeof(int) <= 96 * 1024 && pl.S.n <= 15;
}

bool flat_on(const gsg_plan& pl) {
    if (pl.flat_mode == 0 || !flat_supported(pl)) return false;
    return pl.flat_mode == 1 || pl.S.N * pl.S.D <= FLAT_AUTO_MAX;
}

int flat_tables(gsg_plan& pl) {
    if (pl.flat_cd.p) return 0;
    const gsg::IndexSet& S = pl.S;
    const int D = S.D;
    const long long KDp = S.kDp;
    std::vector<FlatCell> cells((size_t)S.ncells_total * D, FlatCell{-1, 0, 0});
    for (int d = 0; d < D; ++d) GSG_TRY(host_flat_cells(S, d, pl.dirs[d].groups_h, cells));
    std::vector<int> tab((size_t)S.ncells_total * D * FLATCD, 0);
    for (int64_t ci = 0; ci < S.ncells_total; ++ci)
        for (int d = 0; d < D; ++d) {
            const FlatCell& fc = cells[(size_t)ci * D + d];
            if (fc.group < 0) return fail(GSG_ERR_UNSUPPORTED, "internal: flat table does not cover every cell");
            const GroupDev& g = pl.dirs[d].groups_h[fc.group];
            int* o = tab.data() + ((size_t)ci * D + d) * FLATCD;
            // derivative blocks: class p ends row q at the first block column >= 2^p (principal sub-block)
            int e = pl.h_rowptr[fc.q];
            while (e < pl.h_rowptr[fc.q + 1] && pl.h_col[e] < (1 << g.p)) ++e;
            o[0] = pl.h_rowptr[fc.q];
            o[1] = e;
            o[2] = g.S;
            o[3] = (g.p << 16) | fc.q;
            const long long lo = fc.r % g.S, hi = fc.r / g.S;
            for (int l = 0; l <= g.p; ++l) {
                const long long Cd = l <= 1 ? 1 : 1LL << (l - 1);
                o[4 + l] = (int)(g.base[l] / KDp + lo + (long long)g.S * Cd * hi);
            }
        }
    return pl.flat_cd.upload(tab);
}

template <int K>
int launch_flat_k(gsg_plan& pl, const FlatDirs& fd, const FlatMat& M, const double* x, double* y, double beta, int pmin, int pmax) {
    if constexpr (K >= 1 && K <= 5) {
        const int PI = (int)pl.S.kD / K;
        int PIp = 1;
        while (PIp < PI && PIp < 32) PIp <<= 1;
        const int nch = (PI + PIp - 1) / PIp;
        const int nd = std::max(1, fd.ndir);
        // one multi-cell per CTA (GSG_FLAT_CPC: more, for experiments): the kernel cuts long rows -- the coarse cells --
        // into record slices that the CTA's 8 warps share
        static const int cpc_env = getenv("GSG_FLAT_CPC") ? atoi(getenv("GSG_FLAT_CPC")) : 1;
        int cpc = std::max(1, std::min(8, cpc_env));
        while (cpc > 1 && ((size_t)FLAT_WARPS * cpc * pl.S.kDp * sizeof(double) > 40 * 1024 || cpc * nd > 32)) --cpc;
        const int ncells = (int)pl.S.ncells_total;
        const int grid = (ncells + cpc - 1) / cpc;
        const size_t smem = (size_t)FLAT_WARPS * cpc * pl.S.kDp * sizeof(double) + (size_t)nd * PI * sizeof(int);
        auto kern = sweep_flat_kernel<K>;       // default carve-out (the L1 serves the gathers); opt in above 48 KB only (k^D ~ 700)
        if (smem > 48 * 1024) {
            static std::mutex mu;
            static std::map<std::pair<const void*, int>, size_t> done;
            std::lock_guard<std::mutex> lock(mu);
            size_t& have = done[std::make_pair(reinterpret_cast<const void*>(kern), pl.device)];
            if (smem > have) {
                GSG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                have = smem;
            }
        }
        const bool prof = pl.prof_on && pl.prof_used < std::min(pl.prof_cap, pl.prof_ev.size());
        if (prof) GSG_CUDA(cudaEventRecord(pl.prof_ev[pl.prof_used].first, pl.stream));
        kern<<<grid, FLAT_THREADS, smem, pl.stream>>>(x, y, beta, fd, pl.flat_cd.p, pl.S.D, ncells, cpc, M, (int)pl.S.kD,
                                                       (int)pl.S.kDp, PI, PIp, pmin, pmax);
        if (prof) {
            GSG_CUDA(cudaEventRecord(pl.prof_ev[pl.prof_used].second, pl.stream));
            ++pl.prof_used;
            pl.prof_dofs += (double)pl.S.N * fd.ndir;        // directional applies' worth of DOFs
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GSG_CUDA(cudaGetLastError());
        return 0;
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: flat kernel not instantiated");
}

// y = beta * y + sum_{d in mask, c_d != 0} c_d M_d x on the poles of class pmin..pmax; M = the derivative blocks or
// (sq) the pre-squared blocks of the short classes
int flat_apply(gsg_plan& pl, const double* c, unsigned mask, const double* x, double* y, double beta, bool sq = false,
               int pmin = 0, int pmax = MAXL) {
    nvtx_range nvtx_r("rhs: flat");
    GSG_TRY(flat_tables(pl));
    const int K = pl.S.k;
    FlatDirs fd;
    std::memset(&fd, 0, sizeof(fd));
    for (int d = 0; d < pl.S.D; ++d) {
        if (!((mask >> d) & 1) || c[d] == 0.0) continue;
        fd.c[fd.ndir] = c[d];
        fd.A[fd.ndir] = pl.dirs[d].A;
        fd.d[fd.ndir] = d;
        ++fd.ndir;
    }
    FlatMat M;
    std::memset(&M, 0, sizeof(M));
    M.KK2 = pl.KK2;
    if (sq) {
        if (pmax > pl.sq_pmax) return fail(GSG_ERR_UNSUPPORTED, "internal: squared blocks exist for the short classes only");
        M.sq = 1;
        M.rowptr = pl.sq_rowptr.p; M.rowend = pl.sq_rowend.p; M.col = pl.sq_col.p; M.val = pl.sq_val.p;
        for (int p = 0; p <= pl.sq_pmax; ++p) M.cls_row0[p] = pl.sq_cls_row0[p];
    } else {
        M.col = pl.b_col.p; M.val = pl.b_val.p;       // row ranges come with the cell table
    }
    GSG_K_SWITCH(launch_flat_k, pl, fd, M, x, y, beta, pmin, pmax);
    return fail(GSG_ERR_UNSUPPORTED, "internal: flat kernel for k > 5");
}

// y = alpha * D_d x + beta * y   (d 0-based; device layout); x and y must not alias.  The launches
// of one sweep write disjoint parts of y, so they are forked onto auxiliary streams and joined.
int sweep(gsg_plan& pl, int d, double alpha, const double* x, double beta, double* y, bool reduced = false) {
    static const char* const names[16] = {"sweep d=1", "sweep d=2", "sweep d=3", "sweep d=4", "sweep d=5", "sweep d=6",
                                          "sweep d=7", "sweep d=8", "sweep d=9", "sweep d=10", "sweep d=11", "sweep d=12",
                                          "sweep", "sweep", "sweep", "sweep"};
    nvtx_range nvtx_r(names[d & 15]);
    if (!reduced && pl.kind_filter == 0 && !pl.use_sq && flat_on(pl)) {
        double c[FLAT_MAXD] = {0};
        c[d] = alpha;
        if (alpha == 0.0) {                  // y = beta * y
            if (beta == 0.0) GSG_CUDA(cudaMemsetAsync(y, 0, (size_t)pl.S.Npad * sizeof(double), pl.stream));
            else if (beta != 1.0) {
                scale_kernel<<<elementwise_grid(pl, pl.S.Npad), 256, 0, pl.stream>>>(pl.S.Npad, y, beta);
                g_launches.fetch_add(1, std::memory_order_relaxed);
            }
            return 0;
        }
        return flat_apply(pl, c, 1u << d, x, y, beta);
    }
    const Direction& dir = reduced ? pl.dirs_red[d] : pl.dirs[d];
    const size_t nc = dir.classes.size();
    if (beta != 0.0 && beta != 1.0) {     // kernels implement beta in {0, 1}
        scale_kernel<<<elementwise_grid(pl, pl.S.Npad), 256, 0, pl.stream>>>(pl.S.Npad, y, beta);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        beta = 1.0;
    }
    static const bool no_fork = getenv("GSG_NO_FORK") != nullptr;
    const bool fork = nc > 1 && !no_fork;
    if (fork) {
        GSG_CUDA(cudaEventRecord(pl.ev_fork, pl.stream));
        for (size_t i = 1; i < nc; ++i) GSG_CUDA(cudaStreamWaitEvent(pl.aux[i], pl.ev_fork, 0));
    }
    // forked launches first: the long-pole CTAs should be resident before the persistent
    // streaming kernel occupies every SM
    static const int only = getenv("GSG_ONLY_CLASS") ? atoi(getenv("GSG_ONLY_CLASS")) : -1;   // timing aid
    static const int cmask = getenv("GSG_CLASS_MASK") ? atoi(getenv("GSG_CLASS_MASK")) : -1;  // timing aid (bit i = class i)
    static const bool stream_first = getenv("GSG_STREAM_FIRST") != nullptr;
    // Laplacian passes (kind_filter): 1 = only the classes that have pre-squared blocks, 2 = only the others
    auto keep = [&](size_t i) {
        const bool shortc = dir.classes[i].p <= pl.sq_pmax;
        return pl.kind_filter == 0 || (pl.kind_filter == 1 ? shortc : !shortc);
    };
    if (stream_first && nc > 0 && keep(0) && (only < 0 || only == 0) && (cmask < 0 || (cmask & 1)))
        GSG_TRY(launch_class(pl, pl.stream, dir, dir.classes[0], x, y, alpha, beta));
    // GSG_CONSTH_LAST: the constant-bank class is launched AFTER the streaming kernel (its small CTAs can
    // share an SM with a streaming CTA when the latter runs three stages)
    static const bool consth_last = getenv("GSG_CONSTH_LAST") != nullptr;
    for (size_t i = 1; i < nc; ++i) {
        if (only >= 0 && (int)i != only) continue;
        if (cmask >= 0 && !((cmask >> i) & 1)) continue;
        if (!keep(i)) continue;
        if (consth_last && dir.classes[i].kind == Kind::CONSTH) continue;
        cudaStream_t st = fork ? pl.aux[i] : pl.stream;
        GSG_TRY(launch_class(pl, st, dir, dir.classes[i], x, y, alpha, beta));
    }
    if (!stream_first && nc > 0 && keep(0) && (only < 0 || only == 0) && (cmask < 0 || (cmask & 1))) GSG_TRY(launch_class(pl, pl.stream, dir, dir.classes[0], x, y, alpha, beta));
    if (consth_last)
        for (size_t i = 1; i < nc; ++i) {
            if (dir.classes[i].kind != Kind::CONSTH || !keep(i)) continue;
            if (only >= 0 && (int)i != only) continue;
            if (cmask >= 0 && !((cmask >> i) & 1)) continue;
            cudaStream_t st = fork ? pl.aux[i] : pl.stream;
            GSG_TRY(launch_class(pl, st, dir, dir.classes[i], x, y, alpha, beta));
        }
    if (fork) {
        for (size_t i = 1; i < nc; ++i) {
            GSG_CUDA(cudaEventRecord(pl.ev_done[i], pl.aux[i]));
            GSG_CUDA(cudaStreamWaitEvent(pl.stream, pl.ev_done[i], 0));
        }
    }
    return 0;
}

// y = sum_d c_d D_d x with the streaming classes of each direction pair fused (PAIR tiles) -- every c_d != 0
// y = beta0 * y + sum_{d in mask} c_d D_d x with the streaming classes of each direction pair fused (PAIR tiles)
// where both directions of the pair are in the mask; every c_d in the mask must be non-zero
int grad_fused(gsg_plan& pl, const double* c, const double* x, double* y, unsigned mask = ~0u, double beta0 = 0.0) {
    if (flat_on(pl)) return flat_apply(pl, c, mask, x, y, beta0);
    nvtx_range nvtx_r("rhs: fused gradient");
    const int K = pl.S.k, D = pl.S.D;
    const int npairs = (int)pl.pairs.size();
    bool first = true;
    auto next_beta = [&]() { const double b = first ? beta0 : 1.0; first = false; return b; };
    for (int j = 0; j < npairs; ++j) {
        const int da = 2 * j, db = da + 1;
        const bool ina = (mask >> da) & 1, inb = (mask >> db) & 1;
        if (!(ina && inb)) {              // at most one direction of the pair: plain sweeps (complete tables)
            if (ina) GSG_TRY(sweep(pl, da, c[da], x, next_beta(), y));
            if (inb) GSG_TRY(sweep(pl, db, c[db], x, next_beta(), y));
            continue;
        }
        const double beta = next_beta();
        // The PAIR tiles and the two reduced sweeps of this pair write disjoint cells (covered / not covered);
        // pairs are processed one after the other because the next pair's tiles cover other cells.
        // (Measured: running the PAIR kernel on its own stream beside the sweeps is slower, 4.47 vs 4.27 ms per
        // step -- it competes for the same L2 throughput and delays the long-pole CTAs; opt-in GSG_PAIR_OVERLAP.)
        static const bool overlap = getenv("GSG_PAIR_OVERLAP") != nullptr;
        const bool has_tiles = pl.pairs[j].ntiles > 0;
        cudaStream_t ps = overlap ? pl.aux.back() : pl.stream;
        if (overlap) {
            GSG_CUDA(cudaEventRecord(pl.ev_pair_fork, pl.stream));
            GSG_CUDA(cudaStreamWaitEvent(ps, pl.ev_pair_fork, 0));
            GSG_TRY(sweep(pl, da, c[da], x, beta, y, true));
        }
        if (has_tiles) switch (K) {
            case 1: GSG_TRY(launch_pair<1>(pl, ps, j, x, y, c[da], c[db], beta)); break;
            case 2: GSG_TRY(launch_pair<2>(pl, ps, j, x, y, c[da], c[db], beta)); break;
            case 3: GSG_TRY(launch_pair<3>(pl, ps, j, x, y, c[da], c[db], beta)); break;
            case 4: GSG_TRY(launch_pair<4>(pl, ps, j, x, y, c[da], c[db], beta)); break;
            case 5: GSG_TRY(launch_pair<5>(pl, ps, j, x, y, c[da], c[db], beta)); break;
            default: return fail(GSG_ERR_UNSUPPORTED, "internal: pair fusion for k > 5");
        }
        if (!overlap) GSG_TRY(sweep(pl, da, c[da], x, beta, y, true));
        GSG_TRY(sweep(pl, db, c[db], x, 1.0, y, true));
        if (overlap) {
            GSG_CUDA(cudaEventRecord(pl.ev_pair_done, ps));
            GSG_CUDA(cudaStreamWaitEvent(pl.stream, pl.ev_pair_done, 0));
        }
    }
    for (int d = 2 * npairs; d < D; ++d)
        if ((mask >> d) & 1) GSG_TRY(sweep(pl, d, c[d], x, next_beta(), y));
    return 0;
}

bool can_fuse(const gsg_plan& pl, const double* c, unsigned mask = ~0u) {
    if (pl.pairs.empty() || pl.pair_np < 0) return false;
    for (int d = 0; d < pl.S.D; ++d)
        if (((mask >> d) & 1) && c[d] == 0.0) return false;
    return true;
}

// ---- concurrent right-hand side -----------------------------------------------------------------------------
// y = beta0 * y + sum_{d in mask} c_d D_d x with every accumulate launch a reduction at the L2 (bulk reduce-add in
// the streaming kernels, RED.ADD.F64 in the long-pole kernels), so that all pieces that ACCUMULATE commute and can
// be in flight together.  Phase 1 (only when beta0 == 0): the pieces that cover every cell exactly once write with
// beta = 0 -- the first pair's PAIR tiles + its first reduced sweep, or the first plain sweep.  Phase 2: everything
// else with beta = 1; the long-pole launches of ALL remaining directions go to a pool of high-priority streams at
// once (their 30-50 us CTA lifetimes overlap each other and the streaming kernels instead of adding up sweep by
// sweep), the SM-filling persistent kernels (PAIR tiles, streaming classes) follow one another on the main stream.
// `pre_wait[d]` (optional): an event the pieces of direction d must wait for (multi-GPU: the pulled level-0 cells).
struct PoolCtx {
    gsg_plan& pl;
    unsigned used = 0;
    int next = 0, next_lo = 0;
    explicit PoolCtx(gsg_plan& p) : pl(p) {}
    int take_index(int i, cudaStream_t* out, cudaEvent_t after, cudaEvent_t after2) {
        if (!((used >> i) & 1u) && after) GSG_CUDA(cudaStreamWaitEvent(pl.pool[i], after, 0));
        if (after2) GSG_CUDA(cudaStreamWaitEvent(pl.pool[i], after2, 0));
        used |= 1u << i;
        *out = pl.pool[i];
        return 0;
    }
    int take(cudaStream_t* out, cudaEvent_t after, cudaEvent_t after2 = nullptr) { return take_index(next++ & 15, out, after, after2); }
    int take_lo(cudaStream_t* out, cudaEvent_t after, cudaEvent_t after2 = nullptr) { return take_index(16 + (next_lo++ & 7), out, after, after2); }
    int join() {
        for (int i = 0; i < 24; ++i)
            if ((used >> i) & 1u) {
                GSG_CUDA(cudaEventRecord(pl.pool_ev[i], pl.pool[i]));
                GSG_CUDA(cudaStreamWaitEvent(pl.stream, pl.pool_ev[i], 0));
            }
        used = 0;
        return 0;
    }
};

// one sweep's launches with beta = 1: long-pole classes to the pool, the streaming class to `deferred`
int sweep_scatter(gsg_plan& pl, PoolCtx& ctx, int d, double alpha, const double* x, double* y, bool reduced,
                  cudaEvent_t after, cudaEvent_t after2, std::vector<std::pair<const Direction*, const SweepClass*>>& deferred) {
    const Direction& dir = reduced ? pl.dirs_red[d] : pl.dirs[d];
    // a row-tile launch that fills the machine (persistent, one CTA per SM) is an SM-filling kernel like the streaming
    // one: queued with them, in order, instead of in front of them (GSG_RT_POOL=1: the high-priority pool as before)
    static const bool rt_pool = getenv("GSG_RT_POOL") != nullptr;
    for (const SweepClass& c : dir.classes) {
        if (c.kind == Kind::SHORT_TMA || (c.kind == Kind::ROWTILE && !rt_pool && 2 * c.rtgrid >= pl.sm_count)) {
            deferred.emplace_back(&dir, &c);
            continue;
        }
        cudaStream_t st;
        GSG_TRY(ctx.take(&st, after, after2));
        GSG_TRY(launch_class(pl, st, dir, c, x, y, alpha, 1.0));
    }
    return 0;
}

int rhs_concurrent(gsg_plan& pl, const double* c, unsigned mask, const double* x, double* y, double beta0,
                   const cudaEvent_t* pre_wait = nullptr) {
    if (!pre_wait && flat_on(pl)) return flat_apply(pl, c, mask, x, y, beta0);
    nvtx_range nvtx_r("rhs: concurrent");
    const int K = pl.S.k, D = pl.S.D;
    if (beta0 != 0.0 && beta0 != 1.0) return fail(GSG_ERR_ARG, "beta must be 0 or 1");
    const int npairs = (pl.pair_np >= 0) ? (int)pl.pairs.size() : 0;
    auto pair_fused = [&](int j) {
        const int da = 2 * j, db = da + 1;
        return j < npairs && ((mask >> da) & 1) && ((mask >> db) & 1) && c[da] != 0.0 && c[db] != 0.0 &&
               !(pre_wait && (pre_wait[da] || pre_wait[db]));
    };
    auto do_pair = [&](int j, double beta, cudaStream_t st) -> int {
        if (pl.pairs[j].ntiles == 0) return 0;
        const int da = 2 * j, db = da + 1;
        switch (K) {
            case 1: return launch_pair<1>(pl, st, j, x, y, c[da], c[db], beta);
            case 2: return launch_pair<2>(pl, st, j, x, y, c[da], c[db], beta);
            case 3: return launch_pair<3>(pl, st, j, x, y, c[da], c[db], beta);
            case 4: return launch_pair<4>(pl, st, j, x, y, c[da], c[db], beta);
            case 5: return launch_pair<5>(pl, st, j, x, y, c[da], c[db], beta);
        }
        return fail(GSG_ERR_UNSUPPORTED, "internal: pair fusion for k > 5");
    };
    // ---- phase 1: initialise y (beta0 == 0)
    int first_dir = -1;                       // direction whose (reduced) sweep ran in phase 1
    int first_pair = -1;
    if (beta0 == 0.0) {
        for (int d = 0; d < D && first_dir < 0; ++d) {
            if (!((mask >> d) & 1) || c[d] == 0.0 || (pre_wait && pre_wait[d])) continue;
            const int j = d / 2;
            if ((d & 1) == 0 && pair_fused(j)) {
                first_pair = j;
                GSG_TRY(do_pair(j, 0.0, pl.stream));
                GSG_TRY(sweep(pl, d, c[d], x, 0.0, y, true));
            } else if ((d & 1) == 1 && pair_fused(j)) {
                continue;                      // (cannot happen: the even member comes first)
            } else {
                GSG_TRY(sweep(pl, d, c[d], x, 0.0, y, false));
            }
            first_dir = d;
        }
        if (first_dir < 0) {                   // nothing to initialise from: y = 0 on the plan's cells
            int d0 = 0;
            while (d0 < D && pre_wait && pre_wait[d0]) ++d0;
            if (d0 == D) return fail(GSG_ERR_UNSUPPORTED, "internal: no local direction to initialise the right-hand side");
            GSG_TRY(sweep(pl, d0, 0.0, x, 0.0, y, false));
        }
    }
    // ---- phase 2: everything else accumulates; long-pole launches first (pool), then the SM-filling kernels
    GSG_CUDA(cudaEventRecord(pl.ev_p1, pl.stream));
    PoolCtx ctx(pl);
    std::vector<std::pair<const Direction*, const SweepClass*>> deferred;
    std::vector<int> deferred_dir;
    std::vector<int> pairs_todo;
    for (int d = 0; d < D; ++d) {
        if (!((mask >> d) & 1) || c[d] == 0.0) continue;
        const int j = d / 2;
        const bool fused = pair_fused(j);
        if (d == first_dir) continue;
        if (fused && (d & 1) == 0 && j != first_pair) pairs_todo.push_back(j);
        const size_t n0 = deferred.size();
        GSG_TRY(sweep_scatter(pl, ctx, d, c[d], x, y, fused, pl.ev_p1, pre_wait ? pre_wait[d] : nullptr, deferred));
        for (size_t i = n0; i < deferred.size(); ++i) deferred_dir.push_back(d);
    }
    // SM-filling persistent kernels: one low-priority pool stream each, so that the drain of one overlaps the ramp-up
    // of the next (GSG_RHS_ONE_STREAM: all on the main stream, in order)
    static const bool one_stream_env = getenv("GSG_RHS_ONE_STREAM") != nullptr;
    // (event-timed launches stay on the main stream: a start event on a side stream would also time the wait for SMs)
    const bool one_stream = one_stream_env || (pl.prof_on && pl.prof_used < std::min(pl.prof_cap, pl.prof_ev.size()));
    // order of the SM-filling launches: the row-tile launches of all directions, the PAIR launches, the streaming launches
    // (GSG_RT_INTERLEAVE=1: direction by direction)
    static const bool rt_interleave = getenv("GSG_RT_INTERLEAVE") != nullptr;
    std::vector<size_t> dorder;
    for (int pass = 0; pass < 2; ++pass)
        for (size_t i = 0; i < deferred.size(); ++i) {
            const bool is_rt = deferred[i].second->kind == Kind::ROWTILE;
            if (rt_interleave ? pass == 0 : (pass == 0) == is_rt) dorder.push_back(i);
        }
    auto launch_deferred = [&](size_t i) -> int {
        const int d = deferred_dir[i];
        cudaStream_t st = pl.stream;
        if (!one_stream) GSG_TRY(ctx.take_lo(&st, pl.ev_p1, pre_wait ? pre_wait[d] : nullptr));
        else if (pre_wait && pre_wait[d]) GSG_CUDA(cudaStreamWaitEvent(pl.stream, pre_wait[d], 0));
        return launch_class(pl, st, *deferred[i].first, *deferred[i].second, x, y, c[d], 1.0);
    };
    size_t di = 0;
    for (; di < dorder.size() && !rt_interleave && deferred[dorder[di]].second->kind == Kind::ROWTILE; ++di) GSG_TRY(launch_deferred(dorder[di]));
    for (int j : pairs_todo) {
        cudaStream_t st = pl.stream;
        if (!one_stream) GSG_TRY(ctx.take_lo(&st, pl.ev_p1));
        GSG_TRY(do_pair(j, 1.0, st));
    }
    for (; di < dorder.size(); ++di) GSG_TRY(launch_deferred(dorder[di]));
    return ctx.join();
}

// k = -sum_d a_d D_d w
int advect_rhs(gsg_plan& pl, const double* a, const double* w, double* k) {
    {
        double c[16];
        for (int d = 0; d < pl.S.D; ++d) c[d] = -a[d];
        static const bool serial = getenv("GSG_RHS_SERIAL") != nullptr;
        if (!serial) return rhs_concurrent(pl, c, ~0u, w, k, 0.0);
        if (can_fuse(pl, c)) return grad_fused(pl, c, w, k);
    }
    bool first = true;
    for (int d = 0; d < pl.S.D; ++d) {
        if (a[d] == 0.0 && !first) continue;
        GSG_TRY(sweep(pl, d, -a[d], w, first ? 0.0 : 1.0, k));
        first = false;
    }
    return 0;
}

// k = sum_d D_d (D_d u)
// The reference applies the explicit product: lap += D_op * D_op, then lap * x (src/multidim_derivative.jl:71-79).
// Restricted to a pole, D_op * D_op is the square of the pole's principal sub-block, so the short classes
// (p <= sq_pmax, dense blocks anyway) take ONE sweep with the pre-squared blocks S_p -- half the traffic and the
// reference's rounding.  For the long classes the square is completely dense (level 0 touches every cell: 65 536
// blocks instead of 5 308 at k = 3, n = 8), so those poles stay two sparse applications D_d (D_d x).
int laplacian(gsg_plan& pl, const double* u, double* k, double* tmp) {
    nvtx_range nvtx_r("rhs: laplacian");
    const int D = pl.S.D;
    static const bool nosq = getenv("GSG_LAP_NOSQ") != nullptr;
    if (nosq || pl.sq_pmax < 0) {
        for (int d = 0; d < D; ++d) {
            GSG_TRY(sweep(pl, d, 1.0, u, 0.0, tmp));
            GSG_TRY(sweep(pl, d, 1.0, tmp, d == 0 ? 0.0 : 1.0, k));
        }
        return 0;
    }
    const bool any_long = pl.S.n > pl.sq_pmax;
    const bool any_short = pl.S.scheme == 0 || pl.S.n <= pl.sq_pmax;      // full scheme: every pole has p = n
    if (flat_on(pl)) {
        double ones[FLAT_MAXD];
        for (int d = 0; d < FLAT_MAXD; ++d) ones[d] = 1.0;
        // every cell is written here (zero where a cell has no short pole), the long passes accumulate
        GSG_TRY(flat_apply(pl, ones, ~0u, u, k, 0.0, true, 0, pl.sq_pmax));
        if (any_long)
            for (int d = 0; d < D; ++d) {
                GSG_TRY(flat_apply(pl, ones, 1u << d, u, tmp, 0.0, false, pl.sq_pmax + 1, MAXL));
                GSG_TRY(flat_apply(pl, ones, 1u << d, tmp, k, 1.0, false, pl.sq_pmax + 1, MAXL));
            }
        return 0;
    }
    struct Restore {
        gsg_plan& pl;
        ~Restore() { pl.use_sq = false; pl.kind_filter = 0; }
    } restore{pl};
    for (int d = 0; d < D; ++d) {
        const double beta = d == 0 ? 0.0 : 1.0;
        if (any_short) {
            pl.use_sq = true;
            pl.kind_filter = 1;
            GSG_TRY(sweep(pl, d, 1.0, u, beta, k));
        }
        if (any_long) {
            pl.use_sq = false;
            pl.kind_filter = 2;
            GSG_TRY(sweep(pl, d, 1.0, u, 0.0, tmp));
            GSG_TRY(sweep(pl, d, 1.0, tmp, beta, k));
        }
    }
    return 0;
}

// Runs `one_step` nsteps times on pl.stream.  The first step runs eagerly (it also configures the
// kernels' attributes); the second is captured into a CUDA graph (the sweeps' stream fork/join becomes
// graph dependencies) and the rest replay it, so the ~270 launches of a step cost no host time.
template <class Step>
int run_steps(gsg_plan& pl, int64_t nsteps, Step one_step) {
    // Opt-in (GSG_GRAPH=1): measured SLOWER at D=6 (6.2 vs 5.4 ms per step) -- inside a graph the
    // independent kernel nodes of a sweep lose their launch order and stream priorities, the persistent
    // streaming kernel takes every SM first and the long-pole kernels run after it instead of beside it.
    // The flat path (small index sets: a step is 5 launches of a few microseconds each) is launch-latency bound and has
    // no stream fork/join inside a step: there the replay is the default (GSG_NO_GRAPH=1 turns it off).
    static const bool graph_env = getenv("GSG_GRAPH") != nullptr, no_graph_env = getenv("GSG_NO_GRAPH") != nullptr;
    const bool use_graph = nsteps >= 3 && !pl.prof_on && !pl.dbg && (graph_env || (flat_on(pl) && !no_graph_env));
    if (!use_graph) {
        for (int64_t s = 0; s < nsteps; ++s) GSG_TRY(one_step());
        return 0;
    }
    GSG_TRY(one_step());
    if (pl.step_exec) { cudaGraphExecDestroy(pl.step_exec); pl.step_exec = nullptr; }
    const int64_t l0 = g_launches.load();
    cudaGraph_t graph = nullptr;
    GSG_CUDA(cudaStreamBeginCapture(pl.stream, cudaStreamCaptureModeThreadLocal));
    const int rc = one_step();
    const cudaError_t ce = cudaStreamEndCapture(pl.stream, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(GSG_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    const int64_t per_step = g_launches.load() - l0;
    g_launches.fetch_sub(per_step, std::memory_order_relaxed);          // the capture launched nothing
    const cudaError_t ie = cudaGraphInstantiate(&pl.step_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(GSG_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    for (int64_t s = 1; s < nsteps; ++s) GSG_CUDA(cudaGraphLaunch(pl.step_exec, pl.stream));
    g_launches.fetch_add(per_step * (nsteps - 1), std::memory_order_relaxed);
    return 0;
}

// Classical RK4 in its staged form; works for any right-hand side.
template <class Rhs>
int rk4_loop(gsg_plan& pl, int64_t len, double* y, double dt, int64_t nsteps, Rhs rhs) {
    GSG_TRY(pl.wk.resize(len));
    GSG_TRY(pl.wacc.resize(len));
    GSG_TRY(pl.ww.resize(len));
    double* k = pl.wk.p;
    double* acc = pl.wacc.p;
    double* w = pl.ww.p;
    const int grid = elementwise_grid(pl, len);
    return run_steps(pl, nsteps, [&]() -> int {
        nvtx_range nvtx_r("rk4 step (staged)");
        GSG_TRY(rhs(y, k));                                                        // k1
        rk_stage_kernel<<<grid, 256, 0, pl.stream>>>(len, y, k, acc, w, 0.5 * dt, dt / 6.0, 1);
        GSG_TRY(rhs(w, k));                                                        // k2
        rk_stage_kernel<<<grid, 256, 0, pl.stream>>>(len, y, k, acc, w, 0.5 * dt, dt / 3.0, 0);
        GSG_TRY(rhs(w, k));                                                        // k3
        rk_stage_kernel<<<grid, 256, 0, pl.stream>>>(len, y, k, acc, w, dt, dt / 3.0, 0);
        GSG_TRY(rhs(w, k));                                                        // k4
        rk_final_kernel<<<grid, 256, 0, pl.stream>>>(len, y, k, acc, dt / 6.0);
        g_launches.fetch_add(4, std::memory_order_relaxed);
        GSG_CUDA(cudaGetLastError());
        return 0;
    });
}

// Classical RK4 for a LINEAR, time-independent right-hand side f(u) = L u.  The four stages are then
// k1 = L u, k2 = k1 + dt/2 L k1, ... and the update collapses to
//   u += dt L u + dt^2/2 L^2 u + dt^3/6 L^3 u + dt^4/24 L^4 u
// -- the same four operator applications, but one combine pass (48 B per DOF) instead of three stage
// updates and a final one (144 B per DOF).  Identical to the staged form up to rounding.
template <class Rhs>
int rk4_linear_loop(gsg_plan& pl, int64_t len, double* y, double dt, int64_t nsteps, Rhs rhs) {
    GSG_TRY(pl.wk.resize(len));
    GSG_TRY(pl.wacc.resize(len));
    GSG_TRY(pl.ww.resize(len));
    GSG_TRY(pl.wv4.resize(len));
    double* v1 = pl.wk.p;
    double* v2 = pl.wacc.p;
    double* v3 = pl.ww.p;
    double* v4 = pl.wv4.p;
    const int grid = elementwise_grid(pl, len);
    return run_steps(pl, nsteps, [&]() -> int {
        nvtx_range nvtx_r("rk4 step (Taylor form)");
        GSG_TRY(rhs(y, v1));
        GSG_TRY(rhs(v1, v2));
        GSG_TRY(rhs(v2, v3));
        GSG_TRY(rhs(v3, v4));
        rk4_taylor_kernel<<<grid, 256, 0, pl.stream>>>(len, y, v1, v2, v3, v4, dt, dt * dt / 2.0, dt * dt * dt / 6.0,
                                                        dt * dt * dt * dt / 24.0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GSG_CUDA(cudaGetLastError());
        return 0;
    });
}

int check_plan(const gsg_plan* p) {
    if (!p) return fail(GSG_ERR_ARG, "null plan");
    GSG_CUDA(cudaSetDevice(p->device));
    return 0;
}

// reference vector layout <-> device layout (multi-cells padded from kD to kDp doubles)
int copy_in(gsg_plan& pl, double* dev, const double* src, cudaMemcpyKind kind) {
    const gsg::IndexSet& S = pl.S;
    if (S.kD == S.kDp) {
        GSG_CUDA(cudaMemcpyAsync(dev, src, (size_t)S.N * sizeof(double), kind, pl.stream));
    } else {
        GSG_CUDA(cudaMemcpy2DAsync(dev, (size_t)S.kDp * sizeof(double), src, (size_t)S.kD * sizeof(double),
                                   (size_t)S.kD * sizeof(double), (size_t)S.ncells_total, kind, pl.stream));
    }
    return 0;
}

int copy_out(gsg_plan& pl, double* dst, const double* dev, cudaMemcpyKind kind) {
    const gsg::IndexSet& S = pl.S;
    if (S.kD == S.kDp) {
        GSG_CUDA(cudaMemcpyAsync(dst, dev, (size_t)S.N * sizeof(double), kind, pl.stream));
    } else {
        GSG_CUDA(cudaMemcpy2DAsync(dst, (size_t)S.kD * sizeof(double), dev, (size_t)S.kDp * sizeof(double),
                                   (size_t)S.kD * sizeof(double), (size_t)S.ncells_total, kind, pl.stream));
    }
    return 0;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int gsg_version(void) { return 101; }

const char* gsg_last_error(void) { return g_err.c_str(); }

int64_t gsg_launch_count(void) { return g_launches.load(); }

int gsg_device_info(int device, char* buf, size_t buflen) {
    if (!buf || buflen == 0) return fail(GSG_ERR_ARG, "null buffer");
    cudaDeviceProp prop;
    GSG_CUDA(cudaGetDeviceProperties(&prop, device));
    snprintf(buf, buflen, "%s;%d;%d.%d;%zu", prop.name, prop.multiProcessorCount, prop.major, prop.minor,
             (size_t)prop.totalGlobalMem);
    return 0;
}

// ---- host-side setup mirrors ---------------------------------------------------------------
static int check_dkn(int D, int k, int n, int scheme) {
    if (D < 1 || D > 12) return fail(GSG_ERR_ARG, "D out of range [1,12]");
    if (k < 1 || k > gsg::K_MAX) return fail(GSG_ERR_ARG, "DomainError: k out of range [1,10]");
    if (n < 0 || n > gsg::N_MAX_LEVEL) return fail(GSG_ERR_ARG, "n out of range [0,16]");
    if (scheme != 0 && scheme != 1) return fail(GSG_ERR_ARG, "ArgumentError: scheme must be 0 (sparse) or 1 (full)");
    return 0;
}

int gsg_get_size(int D, int k, int n, int scheme, int64_t* size_out) {
    GSG_TRY(check_dkn(D, k, n, scheme));
    if (!size_out) return fail(GSG_ERR_ARG, "null output");
    const int64_t N = gsg::get_size(D, k, n, scheme);
    if (N < 0) return fail(GSG_ERR_UNSUPPORTED, "index set too large (more than 4e6 multi-levels or 2^40 DOFs)");
    *size_out = N;
    return 0;
}

int gsg_basis_v(int k, int level, int cell, int mode, const double* x, int64_t npts, double* out) {
    if (k < 1 || k > gsg::K_MAX || mode < 1 || mode > k) return fail(GSG_ERR_ARG, "DomainError: mode/k");
    if (level < 0 || cell < 1) return fail(GSG_ERR_ARG, "bad level/cell");
    for (int64_t i = 0; i < npts; ++i) out[i] = gsg::v_fn(k, level, cell, mode, x[i]);
    return 0;
}

int gsg_cell_index(double x, int level, int64_t* cell_out) {
    if (!cell_out) return fail(GSG_ERR_ARG, "null output");
    *cell_out = gsg::cell_index(x, level);
    return 0;
}

int gsg_basis_tables(int k, double* leg_out, double* dg_out) {
    if (k < 1 || k > gsg::K_MAX) return fail(GSG_ERR_ARG, "DomainError: k out of range [1,10]");
    const auto& L = gsg::leg_coeffs();
    const auto& T = gsg::dg_coeffs(k);
    if (leg_out)
        for (size_t i = 0; i < L.size(); ++i) std::copy(L[i].begin(), L[i].end(), leg_out + i * L[i].size());
    if (dg_out)
        for (int i = 0; i < k; ++i) std::copy(T[i].begin(), T[i].end(), dg_out + (size_t)i * 2 * k);
    return 0;
}

int gsg_dlf_matrix(int k, int n, int basis, int64_t* nnz_inout, int64_t* colptr, int64_t* rowval, double* nzval) {
    GSG_TRY(check_dkn(1, k, n, 0));
    if (!nnz_inout) return fail(GSG_ERR_ARG, "null nnz");
    if (basis != 0 && basis != 1) return fail(GSG_ERR_ARG, "ArgumentError: basis must be 0 (hier) or 1 (pos)");
    // cache the last matrix: the two-call pattern asks twice
    static thread_local int ck = -1, cn = -1, cb = -1;
    static thread_local gsg::Csc cached;
    if (ck != k || cn != n || cb != basis) {
        cached = basis == 0 ? gsg::periodic_hier_DLF_matrix(k, n) : gsg::periodic_pos_DLF_matrix(k, n);
        ck = k; cn = n; cb = basis;
    }
    if (!nzval) {
        *nnz_inout = cached.nnz();
        return 0;
    }
    if (*nnz_inout < cached.nnz()) return fail(GSG_ERR_ARG, "nnz buffer too small");
    for (int64_t j = 0; j <= cached.n; ++j) colptr[j] = cached.colptr[j] + 1;
    for (int64_t p = 0; p < cached.nnz(); ++p) {
        rowval[p] = cached.rowval[p] + 1;
        nzval[p] = cached.nzval[p];
    }
    *nnz_inout = cached.nnz();
    return 0;
}

int gsg_tensor_construct(int D, int k, int n, int scheme, const double* const* vcoeffs_1d, double* out) {
    GSG_TRY(check_dkn(D, k, n, scheme));
    if (!vcoeffs_1d || !out) return fail(GSG_ERR_ARG, "null pointer");
    gsg::IndexSet S;
    if (!S.build(D, k, n, scheme)) return fail(GSG_ERR_UNSUPPORTED, "index set too large");
    gsg::tensor_construct(S, vcoeffs_1d, out);
    return 0;
}

// structural block pattern of periodic_DLF_matrix(k, n) in the hierarchical basis: out[q * 2^n + r] = 1 iff
// the closed supports of 1-D cells q and r intersect or touch periodically (what the constant-bank kernel
// is unrolled for)
int gsg_block_pattern(int n, unsigned char* out) {
    if (n < 0 || n > gsg::N_MAX_LEVEL || !out) return fail(GSG_ERR_ARG, "bad argument");
    const int nq = 1 << n;
    for (int q = 0; q < nq; ++q)
        for (int r = 0; r < nq; ++r) out[(size_t)q * nq + r] = pat::touch(q, r) ? 1 : 0;
    return 0;
}

// ---- plan -------------------------------------------------------------------------------------
int gsg_plan_create(int D, int k, int n, int scheme, int64_t H_n, const int64_t* H_colptr,
                    const int64_t* H_rowval, const double* H_nzval, int device, gsg_plan** plan_out) {
    GSG_TRY(check_dkn(D, k, n, scheme));
    if (!H_colptr || !H_rowval || !H_nzval || !plan_out) return fail(GSG_ERR_ARG, "null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(GSG_ERR_CUDA, "no CUDA device: libgsgb200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GSG_ERR_ARG, "bad device ordinal");
    GSG_CUDA(cudaSetDevice(device));
    std::unique_ptr<gsg_plan> P(new gsg_plan());
    P->device = device;
    {   // size guards BEFORE the index set is enumerated
        double kd = 1.0;
        for (int i = 0; i < D; ++i) kd *= k;
        if (kd > (double)(1 << 20)) return fail(GSG_ERR_UNSUPPORTED, "k^D too large");
        if (((int64_t)k << n) > 16384) return fail(GSG_ERR_UNSUPPORTED, "k*2^n > 16384 not supported");
    }
    if (!P->S.build(D, k, n, scheme)) return fail(GSG_ERR_UNSUPPORTED, "index set too large (more than 4e6 multi-levels or 2^40 DOFs)");
    if (P->S.Npad / P->S.kDp > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "too many multi-cells");
    cudaDeviceProp prop;
    GSG_CUDA(cudaGetDeviceProperties(&prop, device));
    P->sm_count = prop.multiProcessorCount;
    GSG_CUDA(cudaStreamCreateWithFlags(&P->own_stream, cudaStreamNonBlocking));
    P->stream = P->own_stream;
    GSG_CUDA(cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming));
    GSG_CUDA(cudaEventCreateWithFlags(&P->ev_pair_fork, cudaEventDisableTiming));
    GSG_CUDA(cudaEventCreateWithFlags(&P->ev_pair_done, cudaEventDisableTiming));
    int prio_lo = 0, prio_hi = 0;     // long-pole CTAs first: they run beside the persistent streaming kernel
    GSG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (getenv("GSG_NO_PRIO")) prio_hi = prio_lo;
    P->aux.assign(n + 3, nullptr);
    P->ev_done.assign(n + 3, nullptr);
    for (int i = 0; i < n + 3; ++i) {
        GSG_CUDA(cudaStreamCreateWithPriority(&P->aux[i], cudaStreamNonBlocking, prio_hi));
        GSG_CUDA(cudaEventCreateWithFlags(&P->ev_done[i], cudaEventDisableTiming));
    }
    GSG_TRY(P->tile_counter.resize(64));
    P->pool.assign(24, nullptr);
    P->pool_ev.assign(24, nullptr);
    for (int i = 0; i < 24; ++i) {
        GSG_CUDA(cudaStreamCreateWithPriority(&P->pool[i], cudaStreamNonBlocking, i < 16 ? prio_hi : prio_lo));
        GSG_CUDA(cudaEventCreateWithFlags(&P->pool_ev[i], cudaEventDisableTiming));
    }
    GSG_CUDA(cudaEventCreateWithFlags(&P->ev_p1, cudaEventDisableTiming));
    if (const char* e = getenv("GSG_FLAT")) P->flat_mode = std::max(0, std::min(2, atoi(e)));
    GSG_TRY(build_matrix(*P, H_n, H_colptr, H_rowval, H_nzval));
    P->dirs.resize(D);
    for (int d = 0; d < D; ++d) GSG_TRY(build_direction(*P, d, P->dirs[d], -1));
    GSG_TRY(build_pairs(*P));

    // reconstruct tables
    std::vector<unsigned char> lv;
    std::vector<long long> off;
    for (const gsg::Block& b : P->S.blocks) {
        for (int i = 0; i < D; ++i) lv.push_back((unsigned char)b.level[i]);
        off.push_back(b.poffset);
    }
    GSG_TRY(P->r_level.upload(lv));
    GSG_TRY(P->r_offset.upload(off));
    const auto& L = gsg::leg_coeffs();
    std::vector<double> legflat, dgflat;
    for (const auto& row : L) legflat.insert(legflat.end(), row.begin(), row.end());
    for (const auto& row : gsg::dg_coeffs(k)) dgflat.insert(dgflat.end(), row.begin(), row.end());
    GSG_TRY(P->r_leg.upload(legflat));
    GSG_TRY(P->r_dg.upload(dgflat));
    *plan_out = P.release();
    return 0;
}

int gsg_plan_destroy(gsg_plan* plan) {
    if (!plan) return 0;
    cudaSetDevice(plan->device);
    cudaDeviceSynchronize();
    for (cudaStream_t st : plan->aux) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : plan->ev_done) if (ev) cudaEventDestroy(ev);
    for (cudaStream_t st : plan->pool) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : plan->pool_ev) if (ev) cudaEventDestroy(ev);
    if (plan->ev_p1) cudaEventDestroy(plan->ev_p1);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    if (plan->ev_pair_fork) cudaEventDestroy(plan->ev_pair_fork);
    if (plan->ev_pair_done) cudaEventDestroy(plan->ev_pair_done);
    if (plan->own_stream) cudaStreamDestroy(plan->own_stream);
    for (auto& pr : plan->prof_ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (plan->step_exec) cudaGraphExecDestroy(plan->step_exec);
    delete plan;
    return 0;
}

int gsg_plan_size(const gsg_plan* plan, int64_t* size_out) {
    if (!plan || !size_out) return fail(GSG_ERR_ARG, "null pointer");
    *size_out = plan->S.N;
    return 0;
}

int gsg_plan_dev_size(const gsg_plan* plan, int64_t* size_out) {
    if (!plan || !size_out) return fail(GSG_ERR_ARG, "null pointer");
    *size_out = plan->S.Npad;
    return 0;
}

int gsg_plan_set_stream(gsg_plan* plan, void* stream) {
    if (!plan) return fail(GSG_ERR_ARG, "null plan");
    plan->stream = stream ? (cudaStream_t)stream : plan->own_stream;
    return 0;
}

int gsg_plan_sync(gsg_plan* plan) {
    GSG_TRY(check_plan(plan));
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    return 0;
}

int gsg_pack_dev(gsg_plan* plan, const double* ref_layout_dev, double* dev_layout_dev) {
    GSG_TRY(check_plan(plan));
    if (!ref_layout_dev || !dev_layout_dev) return fail(GSG_ERR_ARG, "null pointer");
    return copy_in(*plan, dev_layout_dev, ref_layout_dev, cudaMemcpyDeviceToDevice);
}

int gsg_unpack_dev(gsg_plan* plan, const double* dev_layout_dev, double* ref_layout_dev) {
    GSG_TRY(check_plan(plan));
    if (!ref_layout_dev || !dev_layout_dev) return fail(GSG_ERR_ARG, "null pointer");
    return copy_out(*plan, ref_layout_dev, dev_layout_dev, cudaMemcpyDeviceToDevice);
}

int gsg_plan_flat_active(const gsg_plan* plan, int* active_out) {
    if (!plan || !active_out) return fail(GSG_ERR_ARG, "null pointer");
    *active_out = flat_on(*plan) ? 1 : 0;
    return 0;
}

// one line per direction: the launches of a sweep as "kind(p, tiles)" -- which kernel serves which pole classes
int gsg_plan_describe(const gsg_plan* plan, char* buf, size_t buflen) {
    if (!plan || !buf || buflen == 0) return fail(GSG_ERR_ARG, "null pointer");
    static const char* const names[] = {"stream", "long", "long2", "consth", "generic", "rowtile"};
    std::string out;
    char tmp[160];
    snprintf(tmp, sizeof tmp, "D=%d k=%d n=%d N=%lld flat=%d pair_np=%d rowtile=%d(pmin=%d C=%d PW=%d RG=%d smem=%zu tiles=%zu)\n",
             plan->S.D, plan->S.k, plan->S.n, (long long)plan->S.N, flat_on(*plan) ? 1 : 0, plan->pair_np, plan->rt_on ? 1 : 0,
             plan->rt_pmin, plan->rt_C, plan->rt_PW, plan->rt_RG, plan->rt_smem, plan->rt_prog.tiles.size());
    out += tmp;
    for (size_t d = 0; d < plan->dirs.size(); ++d) {
        snprintf(tmp, sizeof tmp, "d=%zu:", d + 1);
        out += tmp;
        for (const SweepClass& c : plan->dirs[d].classes) {
            snprintf(tmp, sizeof tmp, " %s(p=%d,%d)", names[(int)c.kind], c.p, c.ntiles);
            out += tmp;
        }
        out += "\n";
    }
    snprintf(buf, buflen, "%s", out.c_str());
    return 0;
}

int gsg_plan_set_flat(gsg_plan* plan, int mode) {
    if (!plan || mode < 0 || mode > 2) return fail(GSG_ERR_ARG, "flat mode must be 0 (tiled), 1 (flat) or 2 (automatic)");
    plan->flat_mode = mode;
    return 0;
}

int gsg_debug_flat_tables(int D, int k, int n, int scheme, int d, int64_t* groups_out, int32_t* cells_out,
                          int64_t* ngroups_out, int64_t* ncells_out) {
    GSG_TRY(check_dkn(D, k, n, scheme));
    if (d < 1 || d > D) return fail(GSG_ERR_ARG, "axis d out of range [1,D]");
    gsg::IndexSet S;
    if (!S.build(D, k, n, scheme)) return fail(GSG_ERR_UNSUPPORTED, "index set too large");
    std::vector<GroupDev> groups;
    GSG_TRY(host_groups(S, d - 1, 0, 0, -1, groups));
    if (ngroups_out) *ngroups_out = (int64_t)groups.size();
    if (ncells_out) *ncells_out = S.ncells_total;
    if (groups_out)
        for (size_t g = 0; g < groups.size(); ++g) {
            int64_t* o = groups_out + g * 20;
            for (int l = 0; l <= MAXL; ++l) o[l] = groups[g].base[l];
            o[17] = groups[g].p; o[18] = groups[g].S; o[19] = groups[g].nitems;
        }
    if (cells_out) {
        std::vector<FlatCell> cells((size_t)S.ncells_total * D, FlatCell{-1, 0, 0});
        GSG_TRY(host_flat_cells(S, d - 1, groups, cells));
        for (int64_t c = 0; c < S.ncells_total; ++c) {
            const FlatCell& fc = cells[(size_t)c * D + (d - 1)];
            cells_out[3 * c] = fc.group; cells_out[3 * c + 1] = fc.r; cells_out[3 * c + 2] = fc.q;
        }
    }
    return 0;
}

// CPU-side check of the row-tile program (no device needed): the tile program of pole class p for the library's own
// H = periodic_DLF_matrix(k, n), multi-cells of k^D doubles.  Two-call pattern: NULL outputs return the counts
// {tiles, groups, blob bytes}.  Layouts: include/gsg_b200.h.
int gsg_debug_rowtile_program(int D, int k, int n, int p, int64_t budget_bytes, int nrg, int32_t* tiles_out, int32_t* groups_out,
                              unsigned char* blob_out, int64_t* counts_out) {
    GSG_TRY(check_dkn(D, k, n, 0));
    if (p < 0 || p > n || !counts_out) return fail(GSG_ERR_ARG, "bad argument");
    const gsg::Csc H = gsg::periodic_hier_DLF_matrix(k, n);
    const int64_t N1 = H.n;
    const int NQ = 1 << n;
    std::vector<double> Hd((size_t)N1 * N1, 0.0);
    std::vector<char> blk((size_t)NQ * NQ, 0);
    for (int64_t j = 0; j < N1; ++j)
        for (int64_t pp = H.colptr[j]; pp < H.colptr[j + 1]; ++pp) {
            const int64_t i = H.rowval[pp];
            Hd[(size_t)i * N1 + j] += H.nzval[pp];
            blk[(size_t)(i / k) * NQ + (j / k)] = 1;
        }
    const int KK2 = (k * k + 1) & ~1;
    std::vector<int> rowptr, col;
    std::vector<double> val;
    block_csr_from_dense(Hd, blk, N1, k, NQ, KK2, rowptr, col, val);
    int64_t kD = 1;
    for (int i = 0; i < D; ++i) kD *= k;
    const int KDp = (int)((kD + 1) & ~int64_t(1));
    RTProgram prog;
    GSG_TRY(rt_build_program(rowptr, col, val, KK2, k, KDp, p, p, (size_t)budget_bytes, nrg, prog));
    counts_out[0] = (int64_t)prog.tiles.size();
    counts_out[1] = (int64_t)prog.groups.size();
    counts_out[2] = (int64_t)prog.blob.size();
    if (tiles_out)
        for (size_t t = 0; t < prog.tiles.size(); ++t) {
            const RTTile& T = prog.tiles[t];
            int32_t* o = tiles_out + t * (8 + RT_MAXX);
            o[0] = T.nx; o[1] = T.rec_ofs; o[2] = T.rec_bytes; o[3] = T.grp0;
            for (int g = 0; g < RT_MAXRG; ++g) o[4 + g] = T.rg_end[g];
            for (int i = 0; i < RT_MAXX; ++i) o[8 + i] = T.xq[i];
        }
    if (groups_out) std::memcpy(groups_out, prog.groups.data(), prog.groups.size() * sizeof(RTGroup));
    if (blob_out) std::memcpy(blob_out, prog.blob.data(), prog.blob.size());
    return 0;
}

// CPU-side check of the row-tile kernel's lane order: the nslots = 32 * C * PW table entries for in-cell stride A
// (padding lanes as ~offset)
int gsg_debug_rowtile_pole_order(int k, int A, int PI, int nslots, int32_t* out) {
    if (k < 1 || A < 1 || PI < 1 || PI % A != 0 || nslots < PI || nslots % 32 != 0 || !out) return fail(GSG_ERR_ARG, "bad argument");
    const std::vector<int> t = rt_pole_order(k, A, PI, nslots);
    std::memcpy(out, t.data(), t.size() * sizeof(int));
    return 0;
}

int gsg_plan_set_rk4_mode(gsg_plan* plan, int mode) {
    if (!plan || (mode != 0 && mode != 1)) return fail(GSG_ERR_ARG, "rk4 mode must be 0 (automatic) or 1 (staged)");
    plan->rk4_mode = mode;
    return 0;
}

// ---- multi-GPU block partition ----------------------------------------------------------------------
int gsg_plan_set_partition(gsg_plan* plan, int rank, int nranks) {
    GSG_TRY(check_plan(plan));
    int bits = 0;
    while ((1 << bits) < nranks) ++bits;
    if (nranks < 1 || (1 << bits) != nranks || (bits > 0 && bits >= plan->S.D) || rank < 0 || rank >= nranks)
        return fail(GSG_ERR_ARG, "partition: nranks must be a power of two < 2^D (one direction stays local) and 0 <= rank < nranks");
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    plan->part_rank = rank;
    plan->part_bits = bits;
    plan->dirs.clear();
    plan->dirs.resize(plan->S.D);
    for (int d = 0; d < plan->S.D; ++d) GSG_TRY(build_direction(*plan, d, plan->dirs[d], -1));
    GSG_TRY(build_pairs(*plan));
    return 0;
}

static int block_owner(const gsg_plan& pl, const gsg::Block& b, int skip_dim /* -1: none */) {
    int o = 0;
    for (int j = 0; j < pl.part_bits; ++j) {
        const int e = pl.S.D - 1 - j;
        if (e == skip_dim) continue;
        if (b.level[e] == 0) o |= 1 << j;
    }
    return o;
}

// kind 0: the blocks this rank owns.  kind 1 (d = 1-based partition dimension): the level_d == 0 blocks of
// the poles that straddle this rank and its partner along d -- the partner pair exchanges exactly these
// (their owner sends the stage input, the sweeping rank returns its contribution).  Offsets and sizes are
// in doubles of the DEVICE layout.  Two-call pattern: offsets == NULL returns the count only.
int gsg_plan_partition_blocks(gsg_plan* plan, int kind, int d, int64_t* offsets, int64_t* sizes, int64_t* count,
                              int* partner_out) {
    if (!plan || !count) return fail(GSG_ERR_ARG, "null pointer");
    const gsg_plan& pl = *plan;
    const int D = pl.S.D;
    int bitj = -1;
    if (kind == 1) {
        if (d < 1 || d > D) return fail(GSG_ERR_ARG, "axis d out of range [1,D]");
        bitj = D - d;                                  // dimension d (1-based) carries bit D - d
        if (bitj >= pl.part_bits) {                    // not a partition dimension: nothing to exchange
            *count = 0;
            if (partner_out) *partner_out = -1;
            return 0;
        }
        if (partner_out) *partner_out = pl.part_rank ^ (1 << bitj);
    } else if (kind != 0) {
        return fail(GSG_ERR_ARG, "kind must be 0 or 1");
    }
    const int others_mask = kind == 1 ? ~(1 << bitj) : ~0;
    int64_t nout = 0;
    for (const gsg::Block& b : pl.S.blocks) {
        bool take;
        if (kind == 0) {
            take = block_owner(pl, b, -1) == pl.part_rank;
        } else {
            int s = 0;
            for (int i = 0; i < D; ++i) s += b.level[i];
            const bool straddles = pl.S.scheme == 1 ? pl.S.n >= 1 : s < pl.S.n;      // pole has p >= 1
            take = b.level[d - 1] == 0 && straddles &&
                   ((block_owner(pl, b, d - 1) ^ pl.part_rank) & others_mask & ((1 << pl.part_bits) - 1)) == 0;
        }
        if (!take) continue;
        if (offsets) {
            offsets[nout] = b.poffset;
            sizes[nout] = b.ncells * pl.S.kDp;
        }
        ++nout;
    }
    *count = nout;
    return 0;
}

// u[c] += c1 v1[c] + ... + c4 v4[c] on the listed multi-cells (device layout; cells = multi-cell indices)
int gsg_rk4_taylor_cells_dev(gsg_plan* plan, const int* cells_dev, int64_t ncells, double* u, const double* v1,
                             const double* v2, const double* v3, const double* v4, double c1, double c2, double c3,
                             double c4) {
    GSG_TRY(check_plan(plan));
    if (!cells_dev || !u || !v1 || !v2 || !v3 || !v4 || ncells < 0) return fail(GSG_ERR_ARG, "bad argument");
    if (ncells == 0) return 0;
    const int grid = (int)std::min<int64_t>(ncells, (int64_t)plan->sm_count * 16);
    rk4_taylor_cells_kernel<<<grid, 256, 0, plan->stream>>>(cells_dev, ncells, (int)plan->S.kDp, u, v1, v2, v3, v4, c1,
                                                            c2, c3, c4);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

int gsg_rk_stage_dev(gsg_plan* plan, int64_t len, const double* u, const double* k, double* acc, double* w,
                     double cw, double ca, int first) {
    GSG_TRY(check_plan(plan));
    if (len < 0 || !u || !k || !acc || !w) return fail(GSG_ERR_ARG, "bad argument");
    if (len == 0) return 0;
    rk_stage_kernel<<<elementwise_grid(*plan, len), 256, 0, plan->stream>>>(len, u, k, acc, w, cw, ca, first);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

int gsg_rk_final_dev(gsg_plan* plan, int64_t len, double* u, const double* k, const double* acc, double ca) {
    GSG_TRY(check_plan(plan));
    if (len < 0 || !u || !k || !acc) return fail(GSG_ERR_ARG, "bad argument");
    if (len == 0) return 0;
    rk_final_kernel<<<elementwise_grid(*plan, len), 256, 0, plan->stream>>>(len, u, k, acc, ca);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

int gsg_profile_enable(gsg_plan* plan, int on) {
    GSG_TRY(check_plan(plan));
    if (on && plan->prof_ev.empty()) {
        plan->prof_ev.resize(4096);
        for (auto& pr : plan->prof_ev) {
            GSG_CUDA(cudaEventCreate(&pr.first));
            GSG_CUDA(cudaEventCreate(&pr.second));
        }
    }
    plan->prof_on = on != 0;
    plan->prof_cap = on > 1 ? (size_t)on : plan->prof_ev.size();     // on > 1: time only the first `on` launches
    plan->prof_used = 0;
    plan->prof_dofs = 0;
    return 0;
}

int gsg_profile_read(gsg_plan* plan, int64_t* launches_out, double* total_ms_out, double* dofs_out) {
    GSG_TRY(check_plan(plan));
    GSG_CUDA(cudaDeviceSynchronize());
    double total = 0;
    for (size_t i = 0; i < plan->prof_used; ++i) {
        float ms = 0;
        GSG_CUDA(cudaEventElapsedTime(&ms, plan->prof_ev[i].first, plan->prof_ev[i].second));
        total += ms;
    }
    if (launches_out) *launches_out = (int64_t)plan->prof_used;
    if (total_ms_out) *total_ms_out = total;
    if (dofs_out) *dofs_out = plan->prof_dofs;
    return 0;
}

// development aid: enable (n > 0) / read back the per-phase clock stamps of the TMA kernel's CTA 0
int gsg_debug_stamps(gsg_plan* plan, long long* out, int n) {
    GSG_TRY(check_plan(plan));
    if (!plan->dbg) {
        GSG_TRY(plan->dbgbuf.resize(8192));
        cudaMemset(plan->dbgbuf.p, 0, 8192 * sizeof(long long));
        plan->dbg = plan->dbgbuf.p;
        return 0;
    }
    GSG_CUDA(cudaDeviceSynchronize());
    if (out && n > 0) GSG_CUDA(cudaMemcpy(out, plan->dbg, sizeof(long long) * std::min(n, 8192), cudaMemcpyDeviceToHost));
    return 0;
}

// development aid: launch a spinner (threads, dynamic smem, duration) on the main (which = 0) or first
// auxiliary stream (which = 1); its CTAs stamp {smid, start, end} into the debug buffer at `slot`
int gsg_debug_spin(gsg_plan* plan, int which, int grid, int threads, int smem, int ns, int slot) {
    GSG_TRY(check_plan(plan));
    if (!plan->dbg) return fail(GSG_ERR_ARG, "enable gsg_debug_stamps first");
    GSG_CUDA(cudaFuncSetAttribute(debug_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaStream_t st = which ? plan->aux[1] : plan->stream;
    debug_spin_kernel<<<grid, threads, smem, st>>>(plan->dbg, slot, ns);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

// ---- device-pointer operator apply (device layout) ------------------------------------------------
int gsg_apply_D_dev(gsg_plan* plan, int d, double alpha, const double* x_dev, double beta, double* y_dev) {
    GSG_TRY(check_plan(plan));
    if (d < 1 || d > plan->S.D) return fail(GSG_ERR_ARG, "axis d out of range [1,D]");
    if (!x_dev || !y_dev || x_dev == y_dev) return fail(GSG_ERR_ARG, "x and y must be distinct non-null device vectors");
    return sweep(*plan, d - 1, alpha, x_dev, beta, y_dev);
}

int gsg_apply_grad_dev(gsg_plan* plan, const double* a, const double* x_dev, double* y_dev) {
    GSG_TRY(check_plan(plan));
    if (!a || !x_dev || !y_dev || x_dev == y_dev) return fail(GSG_ERR_ARG, "bad pointers");
    static const bool serial = getenv("GSG_RHS_SERIAL") != nullptr;
    if (!serial) return rhs_concurrent(*plan, a, ~0u, x_dev, y_dev, 0.0);
    if (can_fuse(*plan, a)) return grad_fused(*plan, a, x_dev, y_dev);
    for (int d = 0; d < plan->S.D; ++d) GSG_TRY(sweep(*plan, d, a[d], x_dev, d == 0 ? 0.0 : 1.0, y_dev));
    return 0;
}

// y = beta * y + sum over the directions d (1-based) whose bit (d-1) is set in dmask of c[d-1] * D_d x; beta in
// {0, 1}.  Direction pairs that lie inside the mask are swept fused (one load of x, one store of y).
int gsg_apply_dirs_dev(gsg_plan* plan, const double* c, unsigned dmask, double beta, const double* x_dev,
                       double* y_dev) {
    GSG_TRY(check_plan(plan));
    if (!c || !x_dev || !y_dev || x_dev == y_dev) return fail(GSG_ERR_ARG, "bad pointers");
    if (beta != 0.0 && beta != 1.0) return fail(GSG_ERR_ARG, "beta must be 0 or 1");
    const int D = plan->S.D;
    dmask &= (D >= 32 ? ~0u : ((1u << D) - 1u));
    if (dmask == 0) return 0;
    static const bool serial = getenv("GSG_RHS_SERIAL") != nullptr;
    if (!serial) {
        bool any = false;
        for (int d = 0; d < D; ++d) any = any || (((dmask >> d) & 1) && c[d] != 0.0);
        if (any || beta == 1.0) return rhs_concurrent(*plan, c, dmask, x_dev, y_dev, beta);
    }
    if (can_fuse(*plan, c, dmask)) return grad_fused(*plan, c, x_dev, y_dev, dmask, beta);
    bool first = true;
    for (int d = 0; d < D; ++d) {
        if (!((dmask >> d) & 1)) continue;
        GSG_TRY(sweep(*plan, d, c[d], x_dev, first ? beta : 1.0, y_dev));
        first = false;
    }
    return 0;
}

int gsg_apply_laplacian_dev(gsg_plan* plan, const double* x_dev, double* y_dev, double* tmp_dev) {
    GSG_TRY(check_plan(plan));
    if (!x_dev || !y_dev || !tmp_dev) return fail(GSG_ERR_ARG, "bad pointers");
    if (x_dev == y_dev || x_dev == tmp_dev || y_dev == tmp_dev)
        return fail(GSG_ERR_ARG, "x, y and tmp must be three distinct device vectors");
    return laplacian(*plan, x_dev, y_dev, tmp_dev);
}

// ---- host-pointer operator apply ------------------------------------------------------------------
static int stage_in(gsg_plan* plan, const double* x) {
    const size_t Np = (size_t)plan->S.Npad;
    GSG_TRY(plan->wx.resize(Np));
    GSG_TRY(plan->wy.resize(Np));
    return copy_in(*plan, plan->wx.p, x, cudaMemcpyHostToDevice);
}

static int stage_out(gsg_plan* plan, double* y) {
    GSG_TRY(copy_out(*plan, y, plan->wy.p, cudaMemcpyDeviceToHost));
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    return 0;
}

int gsg_apply_D(gsg_plan* plan, int d, const double* x, double* y) {
    GSG_TRY(check_plan(plan));
    if (d < 1 || d > plan->S.D) return fail(GSG_ERR_ARG, "axis d out of range [1,D]");
    if (!x || !y) return fail(GSG_ERR_ARG, "null vector");
    GSG_TRY(stage_in(plan, x));
    GSG_TRY(sweep(*plan, d - 1, 1.0, plan->wx.p, 0.0, plan->wy.p));
    return stage_out(plan, y);
}

int gsg_apply_grad(gsg_plan* plan, const double* a, const double* x, double* y) {
    GSG_TRY(check_plan(plan));
    if (!a || !x || !y) return fail(GSG_ERR_ARG, "null pointer");
    GSG_TRY(stage_in(plan, x));
    GSG_TRY(gsg_apply_grad_dev(plan, a, plan->wx.p, plan->wy.p));
    return stage_out(plan, y);
}

int gsg_apply_laplacian(gsg_plan* plan, const double* x, double* y) {
    GSG_TRY(check_plan(plan));
    if (!x || !y) return fail(GSG_ERR_ARG, "null pointer");
    GSG_TRY(stage_in(plan, x));
    GSG_TRY(plan->wtmp.resize((size_t)plan->S.Npad));
    GSG_TRY(laplacian(*plan, plan->wx.p, plan->wy.p, plan->wtmp.p));
    return stage_out(plan, y);
}

// ---- RK4 ------------------------------------------------------------------------------------------
int gsg_rk4_advect_dev(gsg_plan* plan, const double* a, double* y_dev, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!a || !y_dev || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    std::vector<double> av(a, a + plan->S.D);
    gsg_plan& pl = *plan;
    auto rhs = [&](const double* w, double* k) { return advect_rhs(pl, av.data(), w, k); };
    if (pl.rk4_mode == 0) return rk4_linear_loop(pl, plan->S.Npad, y_dev, dt, nsteps, rhs);
    return rk4_loop(pl, plan->S.Npad, y_dev, dt, nsteps, rhs);
}

int gsg_rk4_advect(gsg_plan* plan, const double* a, double* y, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!a || !y || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    GSG_TRY(plan->wx.resize((size_t)plan->S.Npad));
    GSG_TRY(copy_in(*plan, plan->wx.p, y, cudaMemcpyHostToDevice));
    GSG_TRY(gsg_rk4_advect_dev(plan, a, plan->wx.p, dt, nsteps));
    GSG_TRY(copy_out(*plan, y, plan->wx.p, cudaMemcpyDeviceToHost));
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    return 0;
}

int gsg_rk4_wave_dev(gsg_plan* plan, double* u_dev, double* v_dev, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!u_dev || !v_dev || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    gsg_plan& pl = *plan;
    const int64_t Np = pl.S.Npad;
    // state y = [u; v] kept contiguous in a workspace so the stage kernels see one vector
    GSG_TRY(pl.wy.resize(2 * (size_t)Np));
    GSG_TRY(pl.wtmp.resize((size_t)Np));
    double* y = pl.wy.p;
    GSG_CUDA(cudaMemcpyAsync(y, u_dev, Np * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    GSG_CUDA(cudaMemcpyAsync(y + Np, v_dev, Np * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    auto rhs = [&](const double* w, double* k) -> int {
        // [u; v]' = [v; L u]   (src/pdes.jl:22-49)
        GSG_CUDA(cudaMemcpyAsync(k, w + Np, Np * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
        return laplacian(pl, w, k + Np, pl.wtmp.p);
    };
    int rc = pl.rk4_mode == 0 ? rk4_linear_loop(pl, 2 * Np, y, dt, nsteps, rhs) : rk4_loop(pl, 2 * Np, y, dt, nsteps, rhs);
    if (rc) return rc;
    GSG_CUDA(cudaMemcpyAsync(u_dev, y, Np * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    GSG_CUDA(cudaMemcpyAsync(v_dev, y + Np, Np * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    return 0;
}

int gsg_rk4_wave(gsg_plan* plan, double* u, double* v, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!u || !v || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    const size_t Np = (size_t)plan->S.Npad;
    GSG_TRY(plan->wx.resize(2 * Np));
    double* du = plan->wx.p;
    double* dv = plan->wx.p + Np;
    GSG_TRY(copy_in(*plan, du, u, cudaMemcpyHostToDevice));
    GSG_TRY(copy_in(*plan, dv, v, cudaMemcpyHostToDevice));
    GSG_TRY(gsg_rk4_wave_dev(plan, du, dv, dt, nsteps));
    GSG_TRY(copy_out(*plan, u, du, cudaMemcpyDeviceToHost));
    GSG_TRY(copy_out(*plan, v, dv, cudaMemcpyDeviceToHost));
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    return 0;
}

int gsg_energy(gsg_plan* plan, const double* u, const double* udot, double* energy_out) {
    GSG_TRY(check_plan(plan));
    if (!u || !udot || !energy_out) return fail(GSG_ERR_ARG, "null pointer");
    gsg_plan& pl = *plan;
    const size_t Np = (size_t)pl.S.Npad;
    GSG_TRY(pl.wx.resize(Np));
    GSG_TRY(pl.wy.resize(Np));
    GSG_TRY(pl.wred.resize(1));
    GSG_CUDA(cudaMemsetAsync(pl.wred.p, 0, sizeof(double), pl.stream));
    const int grid = elementwise_grid(pl, (int64_t)Np);
    // padding slots are zero (zero-filled allocation, never written), so they add nothing
    GSG_TRY(copy_in(pl, pl.wx.p, udot, cudaMemcpyHostToDevice));
    sumsq_kernel<<<grid, 256, 0, pl.stream>>>((long long)Np, pl.wx.p, pl.wred.p);
    GSG_TRY(copy_in(pl, pl.wx.p, u, cudaMemcpyHostToDevice));
    for (int d = 0; d < pl.S.D; ++d) {
        GSG_TRY(sweep(pl, d, 1.0, pl.wx.p, 0.0, pl.wy.p));
        sumsq_kernel<<<grid, 256, 0, pl.stream>>>((long long)Np, pl.wy.p, pl.wred.p);
    }
    g_launches.fetch_add(pl.S.D + 1, std::memory_order_relaxed);
    GSG_CUDA(cudaMemcpyAsync(energy_out, pl.wred.p, sizeof(double), cudaMemcpyDeviceToHost, pl.stream));
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    return 0;
}

// ---- tensor_construct on the device ----------------------------------------------------------------------
// out_dev (DEVICE layout, length gsg_plan_dev_size; padding slots untouched) = tensor_construct(D, k, n, [v_1..v_D])
// from D host vectors of length k*2^n: what every D >= 3 example builds its initial data with
// (examples/traveling_wave.jl:18-48, examples/vlasov_evolve.jl:27) -- one small upload + one expansion kernel
// instead of an N-entry host loop and a 276 MB copy.
int gsg_tensor_construct_dev(gsg_plan* plan, const double* const* vcoeffs_1d, double* out_dev) {
    GSG_TRY(check_plan(plan));
    if (!vcoeffs_1d || !out_dev) return fail(GSG_ERR_ARG, "null pointer");
    gsg_plan& pl = *plan;
    const gsg::IndexSet& S = pl.S;
    const int n1d = S.k << S.n;
    if (!pl.cell_block.p) {
        std::vector<int> cb((size_t)S.ncells_total);
        size_t c = 0;
        for (size_t b = 0; b < S.blocks.size(); ++b)
            for (int64_t i = 0; i < S.blocks[b].ncells; ++i) cb[c++] = (int)b;
        GSG_TRY(pl.cell_block.upload(cb));
    }
    GSG_TRY(pl.w1d.resize((size_t)S.D * n1d));
    for (int d = 0; d < S.D; ++d) {
        if (!vcoeffs_1d[d]) return fail(GSG_ERR_ARG, "null 1-D coefficient vector");
        GSG_CUDA(cudaMemcpyAsync(pl.w1d.p + (size_t)d * n1d, vcoeffs_1d[d], sizeof(double) * n1d, cudaMemcpyHostToDevice, pl.stream));
    }
    const int grid = (int)std::min<int64_t>(S.ncells_total, (int64_t)pl.sm_count * 16);
    tensor_construct_kernel<<<grid, 256, 0, pl.stream>>>(pl.cell_block.p, S.ncells_total, pl.r_level.p, pl.r_offset.p, pl.w1d.p,
                                                          S.D, S.k, n1d, (int)S.kD, (int)S.kDp, out_dev);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    GSG_CUDA(cudaStreamSynchronize(pl.stream));       // the host vectors may be pageable: do not return before the copies
    return 0;
}

// ---- reconstruct -----------------------------------------------------------------------------------
int gsg_reconstruct_dev(gsg_plan* plan, const double* vcoeffs_dev, const double* points_dev, int64_t npts,
                        double* out_dev) {
    GSG_TRY(check_plan(plan));
    if (!vcoeffs_dev || !points_dev || !out_dev || npts < 0) return fail(GSG_ERR_ARG, "bad argument");
    if (npts == 0) return 0;
    gsg_plan& pl = *plan;
    ReconTables T;
    T.blk_level = pl.r_level.p;
    T.blk_offset = pl.r_offset.p;
    T.leg = pl.r_leg.p;
    T.dg = pl.r_dg.p;
    T.nblocks = (int)pl.S.blocks.size();
    T.D = pl.S.D;
    T.k = pl.S.k;
    T.n = pl.S.n;
    T.KD = (int)pl.S.kD;
    T.KDp = (int)pl.S.kDp;
    T.legw = 2 * (gsg::K_MAX + 1);
    // third-generation kernel (k <= 5, D <= 4): Morton-sorted points, one thread per 2 points, separable contraction
    if (T.k >= 2 && T.k <= 5 && T.D >= 1 && T.D <= 4 && npts >= 64 && npts < (1LL << 31) && !getenv("GSG_RECON_V2")) {
        nvtx_range nvtx_r("reconstruct (sorted, separable)");
        const int bits = std::min(10, 32 / T.D);
        GSG_TRY(pl.rkeys.resize((size_t)npts)); GSG_TRY(pl.rkeys2.resize((size_t)npts));
        GSG_TRY(pl.rperm.resize((size_t)npts)); GSG_TRY(pl.rperm2.resize((size_t)npts));
        recon_keys_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, pl.stream>>>(points_dev, npts, T.D, bits, pl.rkeys.p, pl.rperm.p);
        size_t tbytes = 0;
        GSG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, pl.rkeys.p, pl.rkeys2.p, pl.rperm.p, pl.rperm2.p, (int)npts, 0,
                                                 bits * T.D, pl.stream));
        GSG_TRY(pl.rtemp.resize(tbytes + 16));
        GSG_CUDA(cub::DeviceRadixSort::SortPairs(pl.rtemp.p, tbytes, pl.rkeys.p, pl.rkeys2.p, pl.rperm.p, pl.rperm2.p, (int)npts, 0,
                                                 bits * T.D, pl.stream));
        g_launches.fetch_add(2, std::memory_order_relaxed);
        constexpr int C = 2;
        const unsigned grid3 = (unsigned)((npts + 128 * C - 1) / (128 * C));
        const auto& L = gsg::leg_coeffs();
        const auto& G = gsg::dg_coeffs(T.k);
        int rc3 = -1;
        auto launch3 = [&](auto kc, auto dc) -> int {
            constexpr int K = decltype(kc)::value, DD = decltype(dc)::value;
            ReconBasis<K> Bs;
            for (int m = 0; m < K; ++m) {
                for (int i = 0; i < K; ++i) Bs.leg[m][i] = L[m][i];
                for (int i = 0; i < 2 * K; ++i) Bs.dg[m][i] = G[m][i];
            }
            reconstruct3_kernel<K, DD, C><<<grid3, 128, 0, pl.stream>>>(T, Bs, vcoeffs_dev, points_dev, pl.rperm2.p, npts, out_dev);
            return 0;
        };
#define GSG_R3(KK, DD) if (T.k == KK && T.D == DD) rc3 = launch3(std::integral_constant<int, KK>{}, std::integral_constant<int, DD>{});
        GSG_R3(2, 1) GSG_R3(2, 2) GSG_R3(2, 3) GSG_R3(2, 4) GSG_R3(3, 1) GSG_R3(3, 2) GSG_R3(3, 3) GSG_R3(3, 4)
        GSG_R3(4, 1) GSG_R3(4, 2) GSG_R3(4, 3) GSG_R3(4, 4) GSG_R3(5, 1) GSG_R3(5, 2) GSG_R3(5, 3) GSG_R3(5, 4)
#undef GSG_R3
        if (rc3 == 0) {
            g_launches.fetch_add(1, std::memory_order_relaxed);
            GSG_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const int nwarp = 8;
    // second-generation kernel: factored mode products (low / high half of the dimensions)
    {
        const int nlow = (T.D + 1) / 2;
        long long KL = 1, KH = 1;
        for (int d = 0; d < nlow; ++d) KL *= T.k;
        for (int d = nlow; d < T.D; ++d) KH *= T.k;
        const int n1 = T.n + 1, ntab = T.D * n1 * T.k;
        const size_t lohi_bytes = ((size_t)T.KD * 4 + 15) & ~(size_t)15;
        const size_t pw = ((size_t)T.nblocks * 8 + (size_t)(ntab + KL + KH) * 8 + (size_t)T.D * n1 * 4 + 15) & ~(size_t)15;
        const size_t pdig_bytes = ((size_t)(KL + KH) * 4 + 15) & ~(size_t)15;
        const size_t lvl_bytes = ((size_t)T.nblocks * T.D + 15) & ~(size_t)15;
        const size_t smem2 = lohi_bytes + pdig_bytes + lvl_bytes + pw * nwarp;
        if (KL < 65536 && KH < 65536 && nlow <= 6 && smem2 <= 200 * 1024) {
            static thread_local size_t configured2 = 0;
            GSG_TRY(ensure_smem(reconstruct2_kernel, smem2, configured2));
            const int64_t want2 = (npts + nwarp - 1) / nwarp;
            const int grid2 = (int)std::max<int64_t>(1, std::min<int64_t>(want2, (int64_t)pl.sm_count * 8));
            reconstruct2_kernel<<<grid2, nwarp * 32, smem2, pl.stream>>>(T, vcoeffs_dev, points_dev, npts, out_dev, (int)KL,
                                                                          (int)KH, nlow);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            GSG_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const size_t per_warp = ((size_t)T.D * (T.n + 1) * T.k + (size_t)T.D * (T.n + 1)) * sizeof(double);
    const size_t smem = per_warp * nwarp;
    if (smem > 200 * 1024) return fail(GSG_ERR_UNSUPPORTED, "reconstruct tables exceed shared memory");
    if (smem > 48 * 1024)
        GSG_CUDA(cudaFuncSetAttribute(reconstruct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (npts + nwarp - 1) / nwarp;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)pl.sm_count * 8));
    reconstruct_kernel<<<grid, nwarp * 32, smem, pl.stream>>>(T, vcoeffs_dev, points_dev, npts, out_dev);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

int gsg_reconstruct(gsg_plan* plan, const double* vcoeffs, const double* points, int64_t npts, double* out) {
    GSG_TRY(check_plan(plan));
    if (!vcoeffs || !points || !out || npts < 0) return fail(GSG_ERR_ARG, "bad argument");
    if (npts == 0) return 0;
    gsg_plan& pl = *plan;
    // the reference indexes coeffs[key][cell] with cell = 1 + floor(2^(l-1) x): a point outside [0, 1] (or NaN)
    // is a BoundsError there (src/dg_methods.jl:158-159)
    for (int64_t i = 0; i < npts * pl.S.D; ++i)
        if (!(points[i] >= 0.0 && points[i] <= 1.0))
            return fail(GSG_ERR_ARG, "BoundsError: reconstruct point outside [0, 1]^D (point " + std::to_string(i / pl.S.D) + ")");
    GSG_TRY(pl.wx.resize((size_t)pl.S.Npad));
    GSG_TRY(pl.wpts.resize((size_t)npts * pl.S.D));
    GSG_TRY(pl.wout.resize((size_t)npts));
    GSG_TRY(copy_in(pl, pl.wx.p, vcoeffs, cudaMemcpyHostToDevice));
    GSG_CUDA(cudaMemcpyAsync(pl.wpts.p, points, (size_t)npts * pl.S.D * sizeof(double), cudaMemcpyHostToDevice, pl.stream));
    GSG_TRY(gsg_reconstruct_dev(plan, pl.wx.p, pl.wpts.p, npts, pl.wout.p));
    GSG_CUDA(cudaMemcpyAsync(out, pl.wout.p, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, pl.stream));
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    return 0;
}

// ---- generic SpMV ------------------------------------------------------------------------------------
int gsg_csr_create(int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                   int device, gsg_csr** out) {
    if (!colptr || !rowval || !nzval || !out || m < 0 || n < 0) return fail(GSG_ERR_ARG, "bad argument");
    if (n > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "n exceeds int32 columns");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(GSG_ERR_CUDA, "no CUDA device: libgsgb200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GSG_ERR_ARG, "bad device ordinal");
    GSG_CUDA(cudaSetDevice(device));
    const int64_t nnz = colptr[n] - 1;
    std::vector<long long> rowptr(m + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) {
        const int64_t i = rowval[p] - 1;
        if (i < 0 || i >= m) return fail(GSG_ERR_ARG, "bad rowval (expect 1-based)");
        rowptr[i + 1]++;
    }
    for (int64_t i = 0; i < m; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int> col(nnz);
    std::vector<double> val(nnz);
    std::vector<long long> fill(rowptr.begin(), rowptr.end() - 1);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
            const long long dst = fill[rowval[p] - 1]++;
            col[dst] = (int)j;
            val[dst] = nzval[p];
        }
    std::unique_ptr<gsg_csr> A(new gsg_csr());
    A->device = device;
    A->m = m; A->n = n; A->nnz = nnz;
    GSG_TRY(A->rowptr.upload(rowptr));
    GSG_TRY(A->col.upload(col));
    GSG_TRY(A->val.upload(val));
    *out = A.release();
    return 0;
}

int gsg_csr_destroy(gsg_csr* A) {
    if (A) { cudaSetDevice(A->device); delete A; }
    return 0;
}

int gsg_csr_apply_dev(gsg_csr* A, const double* x_dev, double* y_dev, void* stream) {
    if (!A || !x_dev || !y_dev) return fail(GSG_ERR_ARG, "null pointer");
    GSG_CUDA(cudaSetDevice(A->device));
    if (A->m == 0) return 0;
    constexpr int LANES = 8;
    const int64_t threads = A->m * LANES;
    const int grid = (int)((threads + 255) / 256);
    spmv_csr_kernel<LANES><<<grid, 256, 0, (cudaStream_t)stream>>>(A->m, A->rowptr.p, A->col.p, A->val.p, x_dev, y_dev);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

int gsg_csr_apply(gsg_csr* A, const double* x, double* y) {
    if (!A || !x || !y) return fail(GSG_ERR_ARG, "null pointer");
    GSG_CUDA(cudaSetDevice(A->device));
    GSG_TRY(A->wx.resize((size_t)A->n));
    GSG_TRY(A->wy.resize((size_t)A->m));
    GSG_CUDA(cudaMemcpy(A->wx.p, x, (size_t)A->n * sizeof(double), cudaMemcpyHostToDevice));
    GSG_TRY(gsg_csr_apply_dev(A, A->wx.p, A->wy.p, nullptr));
    GSG_CUDA(cudaMemcpy(y, A->wy.p, (size_t)A->m * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int gsg_spmv_csc(int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                 const double* x, double* y) {
    gsg_csr* A = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    GSG_TRY(gsg_csr_create(m, n, colptr, rowval, nzval, dev, &A));
    const int rc = gsg_csr_apply(A, x, y);
    gsg_csr_destroy(A);
    return rc;
}

}  // extern "C"

#include "multi_gpu.inl"
#include "vlasov.inl"
#include "ode.inl"
