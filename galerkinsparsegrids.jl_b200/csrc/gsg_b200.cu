// gsg_b200.cu -- plan construction, kernel dispatch and the C ABI of libgsgb200.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see __graft_entry__.build()).
#include "../../include/gsg_b200.h"
#include "host_setup.hpp"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
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

template <class T>
struct DevBuf {
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
        return 0;
    }
    int resize(size_t m) {
        if (m <= n && p) return 0;
        if (p) { cudaFree(p); p = nullptr; }
        n = m;
        GSG_CUDA(cudaMalloc(&p, n * sizeof(T)));
        return 0;
    }
};

constexpr int SHORT_MAX_NP = 32;      // register-resident poles up to this length
constexpr int SHORT_MAX_P = 3;
constexpr size_t GENERIC_SMEM_BUDGET = 96 * 1024;
constexpr int SHORT_TILE_DOUBLES = 4096;

struct SweepClass {          // all tiles of one direction with the same pole length
    int p = 0;
    bool is_short = false;
    int NPOLE = 0, Amin = 0; // generic kernel parameters
    size_t smem = 0;
    DevBuf<TileDev> tiles;
    int ntiles = 0;
};

struct Direction {
    int A = 1;               // K^(d-1)
    DevBuf<GroupDev> groups;
    std::vector<SweepClass> classes;
};

}  // namespace

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
    std::vector<std::unique_ptr<DevBuf<double>>> dense;   // index p

    std::vector<Direction> dirs;

    // reconstruct tables
    DevBuf<unsigned char> r_level;
    DevBuf<long long> r_offset;
    DevBuf<double> r_leg, r_dg;

    // workspaces
    DevBuf<double> wx, wy, wk, wacc, ww, wtmp, wred;
    DevBuf<double> wpts, wout;
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

// ------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------
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
    std::vector<int> rowptr(NQ + 1, 0), col;
    std::vector<double> val;
    for (int q = 0; q < NQ; ++q) {
        for (int qc = 0; qc < NQ; ++qc) {
            if (!blk[(size_t)q * NQ + qc]) continue;
            col.push_back(qc);
            const size_t o = val.size();
            val.resize(o + P.KK2, 0.0);
            for (int mo = 0; mo < K; ++mo)
                for (int mi = 0; mi < K; ++mi)
                    val[o + mo * K + mi] = Hd[(size_t)(q * K + mo) * N1 + (qc * K + mi)];
        }
        rowptr[q + 1] = (int)col.size();
    }
    GSG_TRY(P.b_rowptr.upload(rowptr));
    GSG_TRY(P.b_col.upload(col));
    GSG_TRY(P.b_val.upload(val));
    P.dense.resize(n + 1);
    for (int p = 0; p <= n; ++p) {
        const int NP = K << p;
        if (NP > SHORT_MAX_NP || p > SHORT_MAX_P) break;
        std::vector<double> d((size_t)NP * NP);
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j) d[(size_t)i * NP + j] = Hd[(size_t)i * N1 + j];
        P.dense[p].reset(new DevBuf<double>());
        GSG_TRY(P.dense[p]->upload(d));
    }
    return 0;
}

bool short_supported(int K, int p) {
    return K >= 1 && K <= 5 && p <= SHORT_MAX_P && (K << p) <= SHORT_MAX_NP;
}

int build_direction(gsg_plan& P, int d /*0-based*/) {
    const gsg::IndexSet& S = P.S;
    const int D = S.D, K = S.k, n = S.n;
    Direction& dir = P.dirs[d];
    dir.A = pow_int(K, d);
    const int KD = (int)S.kD;
    const int PI = KD / K;

    // groups keyed by the other dims' levels, in layout order of their level_d = 0 block
    std::vector<GroupDev> groups;
    std::vector<std::vector<TileDev>> tiles(n + 1);
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
            g.base[ld] = S.blocks[it->second].offset;
        }
        long long Slo = 1, Shi = 1;
        for (int i = 0; i < d; ++i) Slo *= b0.cells[i];
        for (int i = d + 1; i < D; ++i) Shi *= b0.cells[i];
        if (Slo * Shi > 0x7fffffffLL) return fail(GSG_ERR_UNSUPPORTED, "too many items in a pole group");
        g.S = (int)Slo;
        g.nitems = (int)(Slo * Shi);
        groups.push_back(g);
    }
    GSG_TRY(dir.groups.upload(groups));

    for (int p = 0; p <= n; ++p) {
        SweepClass c;
        c.p = p;
        const int NQ = 1 << p, NP = K * NQ;
        c.is_short = short_supported(K, p);
        std::vector<TileDev> tl;
        if (c.is_short) {
            int nr_max = std::max(1, SHORT_TILE_DOUBLES / (NQ * KD));
            c.NPOLE = PI;
            for (size_t gi = 0; gi < groups.size(); ++gi) {
                if (groups[gi].p != p) continue;
                for (int r0 = 0; r0 < groups[gi].nitems; r0 += nr_max)
                    tl.push_back(TileDev{(int)gi, r0, std::min(nr_max, groups[gi].nitems - r0), 0});
            }
            c.smem = ((size_t)((NP * NP + 1) & ~1) + (size_t)NQ * nr_max * KD) * sizeof(double);
        } else {
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
    // heavy classes first so their long CTAs overlap the streaming ones
    std::sort(dir.classes.begin(), dir.classes.end(), [](const SweepClass& a, const SweepClass& b) { return a.p > b.p; });
    return 0;
}

int launch_check(const char* what, int K, const SweepClass& c) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && getenv("GSG_DEBUG_SYNC")) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s<K=%d> p=%d ntiles=%d smem=%zu NPOLE=%d Amin=%d: %s", what, K, c.p, c.ntiles,
                 c.smem, c.NPOLE, c.Amin, cudaGetErrorString(e));
        return fail(GSG_ERR_CUDA, buf);
    }
    return 0;
}

template <int K, int P>
int launch_short_kp(gsg_plan& pl, const Direction& dir, const SweepClass& c, const double* x, double* y,
                    double alpha, double beta) {
    auto kern = sweep_short_kernel<K, P>;
    static thread_local size_t configured = 0;
    if (c.smem > 48 * 1024 && c.smem > configured) {
        GSG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        configured = c.smem;
    }
    kern<<<c.ntiles, 256, c.smem, pl.stream>>>(x, y, alpha, beta, dir.groups.p, c.tiles.p, pl.dense[P]->p,
                                                (int)pl.S.kD, dir.A);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_check("sweep_short", K, c);
}

template <int K>
int launch_short_k(gsg_plan& pl, const Direction& dir, const SweepClass& c, const double* x, double* y,
                   double alpha, double beta) {
    switch (c.p) {
        case 0: return launch_short_kp<K, 0>(pl, dir, c, x, y, alpha, beta);
        case 1: if constexpr ((K << 1) <= SHORT_MAX_NP) return launch_short_kp<K, 1>(pl, dir, c, x, y, alpha, beta); break;
        case 2: if constexpr ((K << 2) <= SHORT_MAX_NP) return launch_short_kp<K, 2>(pl, dir, c, x, y, alpha, beta); break;
        case 3: if constexpr ((K << 3) <= SHORT_MAX_NP) return launch_short_kp<K, 3>(pl, dir, c, x, y, alpha, beta); break;
    }
    return fail(GSG_ERR_UNSUPPORTED, "internal: short class not instantiated");
}

template <int K>
int launch_generic_k(gsg_plan& pl, const Direction& dir, const SweepClass& c, const double* x, double* y,
                     double alpha, double beta) {
    auto kern = sweep_generic_kernel<K>;
    static thread_local size_t configured = 0;
    if (c.smem > 48 * 1024 && c.smem > configured) {
        GSG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        configured = c.smem;
    }
    Bcsr M{pl.b_rowptr.p, pl.b_col.p, pl.b_val.p, pl.KK2};
    kern<<<c.ntiles, 256, c.smem, pl.stream>>>(x, y, alpha, beta, dir.groups.p, c.tiles.p, M, pl.S.k, c.p,
                                                (int)pl.S.kD, dir.A, c.NPOLE, c.Amin);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_check("sweep_generic", K, c);
}

// y = alpha * D_d x + beta * y   (d 0-based); x and y must not alias
int sweep(gsg_plan& pl, int d, double alpha, const double* x, double beta, double* y) {
    const Direction& dir = pl.dirs[d];
    const int K = pl.S.k;
    for (const SweepClass& c : dir.classes) {
        int rc;
        if (c.is_short) {
            switch (K) {
                case 1: rc = launch_short_k<1>(pl, dir, c, x, y, alpha, beta); break;
                case 2: rc = launch_short_k<2>(pl, dir, c, x, y, alpha, beta); break;
                case 3: rc = launch_short_k<3>(pl, dir, c, x, y, alpha, beta); break;
                case 4: rc = launch_short_k<4>(pl, dir, c, x, y, alpha, beta); break;
                case 5: rc = launch_short_k<5>(pl, dir, c, x, y, alpha, beta); break;
                default: rc = fail(GSG_ERR_UNSUPPORTED, "short kernel: k > 5");
            }
        } else {
            switch (K) {
                case 1: rc = launch_generic_k<1>(pl, dir, c, x, y, alpha, beta); break;
                case 2: rc = launch_generic_k<2>(pl, dir, c, x, y, alpha, beta); break;
                case 3: rc = launch_generic_k<3>(pl, dir, c, x, y, alpha, beta); break;
                case 4: rc = launch_generic_k<4>(pl, dir, c, x, y, alpha, beta); break;
                case 5: rc = launch_generic_k<5>(pl, dir, c, x, y, alpha, beta); break;
                default: rc = launch_generic_k<0>(pl, dir, c, x, y, alpha, beta); break;
            }
        }
        if (rc) return rc;
    }
    return 0;
}

int elementwise_grid(const gsg_plan& pl, int64_t N) {
    const int64_t want = (N + 255) / 256;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)pl.sm_count * 16));
}

// k = -sum_d a_d D_d w
int advect_rhs(gsg_plan& pl, const double* a, const double* w, double* k) {
    bool first = true;
    for (int d = 0; d < pl.S.D; ++d) {
        if (a[d] == 0.0 && !first) continue;
        GSG_TRY(sweep(pl, d, -a[d], w, first ? 0.0 : 1.0, k));
        first = false;
    }
    return 0;
}

// k = sum_d D_d (D_d u)
int laplacian(gsg_plan& pl, const double* u, double* k, double* tmp) {
    for (int d = 0; d < pl.S.D; ++d) {
        GSG_TRY(sweep(pl, d, 1.0, u, 0.0, tmp));
        GSG_TRY(sweep(pl, d, 1.0, tmp, d == 0 ? 0.0 : 1.0, k));
    }
    return 0;
}

template <class Rhs>
int rk4_loop(gsg_plan& pl, int64_t len, double* y, double dt, int64_t nsteps, Rhs rhs) {
    GSG_TRY(pl.wk.resize(len));
    GSG_TRY(pl.wacc.resize(len));
    GSG_TRY(pl.ww.resize(len));
    double* k = pl.wk.p;
    double* acc = pl.wacc.p;
    double* w = pl.ww.p;
    const int grid = elementwise_grid(pl, len);
    for (int64_t s = 0; s < nsteps; ++s) {
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
    }
    return 0;
}

int check_plan(const gsg_plan* p) {
    if (!p) return fail(GSG_ERR_ARG, "null plan");
    GSG_CUDA(cudaSetDevice(p->device));
    return 0;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int gsg_version(void) { return 100; }

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
    *size_out = gsg::get_size(D, k, n, scheme);
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
    S.build(D, k, n, scheme);
    gsg::tensor_construct(S, vcoeffs_1d, out);
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
    P->S.build(D, k, n, scheme);
    if (P->S.kD > (1 << 20)) return fail(GSG_ERR_UNSUPPORTED, "k^D too large");
    cudaDeviceProp prop;
    GSG_CUDA(cudaGetDeviceProperties(&prop, device));
    P->sm_count = prop.multiProcessorCount;
    GSG_CUDA(cudaStreamCreateWithFlags(&P->own_stream, cudaStreamNonBlocking));
    P->stream = P->own_stream;
    GSG_TRY(build_matrix(*P, H_n, H_colptr, H_rowval, H_nzval));
    P->dirs.resize(D);
    for (int d = 0; d < D; ++d) GSG_TRY(build_direction(*P, d));

    // reconstruct tables
    std::vector<unsigned char> lv;
    std::vector<long long> off;
    for (const gsg::Block& b : P->S.blocks) {
        for (int i = 0; i < D; ++i) lv.push_back((unsigned char)b.level[i]);
        off.push_back(b.offset);
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
    if (plan->own_stream) {
        cudaStreamSynchronize(plan->own_stream);
        cudaStreamDestroy(plan->own_stream);
    }
    delete plan;
    return 0;
}

int gsg_plan_size(const gsg_plan* plan, int64_t* size_out) {
    if (!plan || !size_out) return fail(GSG_ERR_ARG, "null pointer");
    *size_out = plan->S.N;
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

// ---- device-pointer operator apply ----------------------------------------------------------------
int gsg_apply_D_dev(gsg_plan* plan, int d, double alpha, const double* x_dev, double beta, double* y_dev) {
    GSG_TRY(check_plan(plan));
    if (d < 1 || d > plan->S.D) return fail(GSG_ERR_ARG, "axis d out of range [1,D]");
    if (!x_dev || !y_dev || x_dev == y_dev) return fail(GSG_ERR_ARG, "x and y must be distinct non-null device vectors");
    return sweep(*plan, d - 1, alpha, x_dev, beta, y_dev);
}

int gsg_apply_grad_dev(gsg_plan* plan, const double* a, const double* x_dev, double* y_dev) {
    GSG_TRY(check_plan(plan));
    if (!a || !x_dev || !y_dev || x_dev == y_dev) return fail(GSG_ERR_ARG, "bad pointers");
    for (int d = 0; d < plan->S.D; ++d) GSG_TRY(sweep(*plan, d, a[d], x_dev, d == 0 ? 0.0 : 1.0, y_dev));
    return 0;
}

int gsg_apply_laplacian_dev(gsg_plan* plan, const double* x_dev, double* y_dev, double* tmp_dev) {
    GSG_TRY(check_plan(plan));
    if (!x_dev || !y_dev || !tmp_dev) return fail(GSG_ERR_ARG, "bad pointers");
    return laplacian(*plan, x_dev, y_dev, tmp_dev);
}

// ---- host-pointer operator apply ------------------------------------------------------------------
static int stage_in(gsg_plan* plan, const double* x) {
    const size_t N = (size_t)plan->S.N;
    GSG_TRY(plan->wx.resize(N));
    GSG_TRY(plan->wy.resize(N));
    GSG_CUDA(cudaMemcpyAsync(plan->wx.p, x, N * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    return 0;
}

static int stage_out(gsg_plan* plan, double* y) {
    const size_t N = (size_t)plan->S.N;
    GSG_CUDA(cudaMemcpyAsync(y, plan->wy.p, N * sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
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
    GSG_TRY(plan->wtmp.resize((size_t)plan->S.N));
    GSG_TRY(laplacian(*plan, plan->wx.p, plan->wy.p, plan->wtmp.p));
    return stage_out(plan, y);
}

// ---- RK4 ------------------------------------------------------------------------------------------
int gsg_rk4_advect_dev(gsg_plan* plan, const double* a, double* y_dev, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!a || !y_dev || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    std::vector<double> av(a, a + plan->S.D);
    gsg_plan& pl = *plan;
    return rk4_loop(pl, plan->S.N, y_dev, dt, nsteps,
                    [&](const double* w, double* k) { return advect_rhs(pl, av.data(), w, k); });
}

int gsg_rk4_advect(gsg_plan* plan, const double* a, double* y, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!a || !y || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    const size_t N = (size_t)plan->S.N;
    GSG_TRY(plan->wx.resize(N));
    GSG_CUDA(cudaMemcpyAsync(plan->wx.p, y, N * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    GSG_TRY(gsg_rk4_advect_dev(plan, a, plan->wx.p, dt, nsteps));
    GSG_CUDA(cudaMemcpyAsync(y, plan->wx.p, N * sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    return 0;
}

int gsg_rk4_wave_dev(gsg_plan* plan, double* u_dev, double* v_dev, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!u_dev || !v_dev || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    gsg_plan& pl = *plan;
    const int64_t N = pl.S.N;
    // state y = [u; v] kept contiguous in a workspace so the stage kernels see one vector
    GSG_TRY(pl.wy.resize(2 * (size_t)N));
    GSG_TRY(pl.wtmp.resize((size_t)N));
    double* y = pl.wy.p;
    GSG_CUDA(cudaMemcpyAsync(y, u_dev, N * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    GSG_CUDA(cudaMemcpyAsync(y + N, v_dev, N * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    int rc = rk4_loop(pl, 2 * N, y, dt, nsteps, [&](const double* w, double* k) {
        // [u; v]' = [v; L u]   (src/pdes.jl:22-49)
        GSG_CUDA(cudaMemcpyAsync(k, w + N, N * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
        return laplacian(pl, w, k + N, pl.wtmp.p);
    });
    if (rc) return rc;
    GSG_CUDA(cudaMemcpyAsync(u_dev, y, N * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    GSG_CUDA(cudaMemcpyAsync(v_dev, y + N, N * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
    return 0;
}

int gsg_rk4_wave(gsg_plan* plan, double* u, double* v, double dt, int64_t nsteps) {
    GSG_TRY(check_plan(plan));
    if (!u || !v || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    const size_t N = (size_t)plan->S.N;
    GSG_TRY(plan->wx.resize(2 * N));
    double* du = plan->wx.p;
    double* dv = plan->wx.p + N;
    GSG_CUDA(cudaMemcpyAsync(du, u, N * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    GSG_CUDA(cudaMemcpyAsync(dv, v, N * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    GSG_TRY(gsg_rk4_wave_dev(plan, du, dv, dt, nsteps));
    GSG_CUDA(cudaMemcpyAsync(u, du, N * sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    GSG_CUDA(cudaMemcpyAsync(v, dv, N * sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    return 0;
}

int gsg_energy(gsg_plan* plan, const double* u, const double* udot, double* energy_out) {
    GSG_TRY(check_plan(plan));
    if (!u || !udot || !energy_out) return fail(GSG_ERR_ARG, "null pointer");
    gsg_plan& pl = *plan;
    const size_t N = (size_t)pl.S.N;
    GSG_TRY(pl.wx.resize(N));
    GSG_TRY(pl.wy.resize(N));
    GSG_TRY(pl.wred.resize(1));
    GSG_CUDA(cudaMemsetAsync(pl.wred.p, 0, sizeof(double), pl.stream));
    const int grid = elementwise_grid(pl, (int64_t)N);
    GSG_CUDA(cudaMemcpyAsync(pl.wx.p, udot, N * sizeof(double), cudaMemcpyHostToDevice, pl.stream));
    sumsq_kernel<<<grid, 256, 0, pl.stream>>>((long long)N, pl.wx.p, pl.wred.p);
    GSG_CUDA(cudaMemcpyAsync(pl.wx.p, u, N * sizeof(double), cudaMemcpyHostToDevice, pl.stream));
    for (int d = 0; d < pl.S.D; ++d) {
        GSG_TRY(sweep(pl, d, 1.0, pl.wx.p, 0.0, pl.wy.p));
        sumsq_kernel<<<grid, 256, 0, pl.stream>>>((long long)N, pl.wy.p, pl.wred.p);
    }
    g_launches.fetch_add(pl.S.D + 1, std::memory_order_relaxed);
    GSG_CUDA(cudaMemcpyAsync(energy_out, pl.wred.p, sizeof(double), cudaMemcpyDeviceToHost, pl.stream));
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
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
    T.legw = 2 * (gsg::K_MAX + 1);
    const int nwarp = 8;
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
    const size_t N = (size_t)pl.S.N;
    GSG_TRY(pl.wx.resize(N));
    GSG_TRY(pl.wpts.resize((size_t)npts * pl.S.D));
    GSG_TRY(pl.wout.resize((size_t)npts));
    GSG_CUDA(cudaMemcpyAsync(pl.wx.p, vcoeffs, N * sizeof(double), cudaMemcpyHostToDevice, pl.stream));
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
