/* csc_kernels.c -- C helper of the CPU oracle: TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Restates, for sizes the pure-Python oracle cannot reach, the two pieces of the reference's
 * hot path that cost CPU time:
 *   (1) the column loop of D_matrix(Val d, k, n, VD, DV, scheme)
 *       (src/multidim_derivative.jl:32-55): for column j = (l, c, m) take column (l_d, c_d, m_d)
 *       of the 1-D matrix H, replace entry d by each stored row (l', c', m'), skip if cut off
 *       (src/schemes.jl:21-23), look the row up in the vector layout (src/dg_vmethods.jl:102-142;
 *       done arithmetically here instead of through the Dict);
 *   (2) `A * x` for SparseMatrixCSC{Float64,Int64} (Julia 1.0 SparseArrays mul!, called from
 *       the RHS closures src/pdes.jl:63,65,179-180): per column j, y[rowval[p]] += nzval[p]*x[j],
 *       plain multiply-then-add, 64-bit indices.
 * plus an OpenMP row-parallel CSR product standing in for the optional MKLSparse path
 * (src/GalerkinSparseGrids.jl:5-7).  Pinned against gsg_oracle.D_matrix_literal in
 * tests/test_oracle_pins.py.  Nothing in the product links or loads this file.
 *
 * Build: gcc -O3 -march=native -ffp-contract=off -fopenmp -fPIC -shared (see __graft_entry__.build)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 12

typedef struct {
    int D, k, n, scheme;
    int64_t kD, N;
    int nblocks;
    int* level;        /* nblocks * D */
    int64_t* offset;   /* nblocks + 1 */
    int* table;        /* (n+1)^D -> block index or -1 */
} index_set;

static int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }
static int cells_of(int l) { return l <= 1 ? 1 : 1 << (l - 1); }

static int build_index_set(index_set* S, int D, int k, int n, int scheme) {
    S->D = D; S->k = k; S->n = n; S->scheme = scheme;
    S->kD = ipow(k, D);
    int64_t nt = ipow(n + 1, D);
    S->table = (int*)malloc(sizeof(int) * nt);
    S->level = (int*)malloc(sizeof(int) * nt * D);
    S->offset = (int64_t*)malloc(sizeof(int64_t) * (nt + 1));
    if (!S->table || !S->level || !S->offset) return -1;
    int lv[MAXD] = {0};
    int nb = 0;
    int64_t off = 0;
    for (int64_t t = 0; t < nt; ++t) {              /* first dimension fastest */
        int sum = 0;
        for (int i = 0; i < D; ++i) sum += lv[i];
        if (scheme == 1 || sum <= n) {
            int64_t nc = 1;
            for (int i = 0; i < D; ++i) { S->level[nb * D + i] = lv[i]; nc *= cells_of(lv[i]); }
            S->offset[nb] = off;
            off += nc * S->kD;
            S->table[t] = nb++;
        } else {
            S->table[t] = -1;
        }
        int i = 0;
        while (i < D && ++lv[i] > n) lv[i++] = 0;
    }
    S->offset[nb] = off;
    S->nblocks = nb;
    S->N = off;
    return 0;
}

static void free_index_set(index_set* S) { free(S->table); free(S->level); free(S->offset); }

/* one cached index set: slab assembly is called thousands of times with the same (D,k,n,scheme) */
static index_set g_cache;
static int g_cache_valid = 0;

static index_set* cached_index_set(int D, int k, int n, int scheme) {
    if (g_cache_valid && g_cache.D == D && g_cache.k == k && g_cache.n == n && g_cache.scheme == scheme)
        return &g_cache;
    if (g_cache_valid) { free_index_set(&g_cache); g_cache_valid = 0; }
    if (build_index_set(&g_cache, D, k, n, scheme)) return 0;
    g_cache_valid = 1;
    return &g_cache;
}

int64_t gsgo_get_size(int D, int k, int n, int scheme) {
    index_set S;
    if (build_index_set(&S, D, k, n, scheme)) return -1;
    int64_t N = S.N;
    free_index_set(&S);
    return N;
}

/* Columns [col_begin, col_end) of D_d (d 1-based) in CSC with 0-based int64 indices.
 * Returns nnz of the slab; if it exceeds cap nothing beyond colptr_out is written (call once
 * with cap = 0 to size the buffers).  H is CSC, 0-based. */
int64_t gsgo_assemble_cols(int D, int k, int n, int scheme, int d, const int64_t* Hcolptr,
                           const int64_t* Hrowval, const double* Hnzval, int64_t col_begin,
                           int64_t col_end, int64_t* colptr_out, int64_t* rowval_out,
                           double* nzval_out, int64_t cap) {
    if (D > MAXD) return -1;
    index_set* Sp = cached_index_set(D, k, n, scheme);
    if (!Sp) return -1;
    const index_set S = *Sp;
    const int dd = d - 1;
    int64_t nnz = 0;
    int blk = 0;
    int64_t radix[MAXD];
    radix[0] = 1;
    for (int i = 1; i < D; ++i) radix[i] = radix[i - 1] * (n + 1);
    for (int64_t j = col_begin; j < col_end; ++j) {
        while (S.offset[blk + 1] <= j) ++blk;
        while (S.offset[blk] > j) --blk;
        const int* lv = S.level + (int64_t)blk * D;
        int64_t loc = j - S.offset[blk];
        int64_t celllin = loc / S.kD, modelin = loc % S.kD;
        int c[MAXD], m[MAXD], C[MAXD];
        int64_t t = celllin;
        for (int i = 0; i < D; ++i) { C[i] = cells_of(lv[i]); c[i] = (int)(t % C[i]); t /= C[i]; }
        t = modelin;
        for (int i = 0; i < D; ++i) { m[i] = (int)(t % k); t /= k; }
        const int ld = lv[dd];
        const int64_t j1 = (int64_t)k * ((ld == 0 ? 0 : (1 << (ld - 1))) + c[dd]) + m[dd];
        int sum_other = 0;
        int64_t tkey = 0;
        for (int i = 0; i < D; ++i) if (i != dd) { sum_other += lv[i]; tkey += radix[i] * lv[i]; }
        /* strides of dim d inside the target block depend on its own cell count */
        int64_t cell_lo = 0, stride_lo = 1;      /* cells of dims < d */
        for (int i = 0; i < dd; ++i) { cell_lo += c[i] * stride_lo; stride_lo *= C[i]; }
        int64_t cell_hi = 0, stride_hi = 1;      /* cells of dims > d */
        for (int i = dd + 1; i < D; ++i) { cell_hi += c[i] * stride_hi; stride_hi *= C[i]; }
        int64_t mode_other = 0;
        { int64_t ms = 1; for (int i = 0; i < D; ++i) { if (i != dd) mode_other += m[i] * ms; ms *= k; } }
        const int64_t mstride_d = ipow(k, dd);
        colptr_out[j - col_begin] = nnz;
        /* rows of H are sorted and the 1-D index grows with the level, so under the sparse
         * cutoff every entry from the first cut-off one onwards is cut off too */
        const int64_t i1_end = (scheme == 0) ? ((int64_t)k << (n - sum_other)) : ((int64_t)k << n);
        for (int64_t p = Hcolptr[j1]; p < Hcolptr[j1 + 1]; ++p) {
            const int64_t i1 = Hrowval[p];
            if (i1 >= i1_end) break;
            const int q = (int)(i1 / k), m2 = (int)(i1 % k);
            int l2 = 0, c2 = 0;
            if (q > 0) { l2 = 1; while ((1 << l2) <= q) ++l2; c2 = q - (1 << (l2 - 1)); }
            if (scheme == 0 && sum_other + l2 > n) continue;            /* cutoff */
            if (l2 > n) continue;
            const int b2 = S.table[tkey + radix[dd] * l2];
            const int64_t C2 = cells_of(l2);
            const int64_t cl = cell_lo + stride_lo * (c2 + C2 * cell_hi);
            const int64_t row = S.offset[b2] + cl * S.kD + mode_other + m2 * mstride_d;
            if (nnz < cap) { rowval_out[nnz] = row; nzval_out[nnz] = Hnzval[p]; }
            ++nnz;
        }
    }
    colptr_out[col_end - col_begin] = nnz;
    return nnz;
}

/* y[rowval[p]] += nzval[p] * x[j]   -- Julia's SparseMatrixCSC * Vector inner loops */
void gsgo_spmv_csc(int64_t ncols, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                   const double* x, double* y) {
    for (int64_t j = 0; j < ncols; ++j) {
        const double xj = x[j];
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) y[rowval[p]] += nzval[p] * xj;
    }
}

/* y = D_d x through the per-pole principal-sub-block identity (verified against the
 * assembled matrix in tests): every pole (1-D fibre along axis d) is multiplied by H[0:N', 0:N'],
 * N' = k 2^(n - s).  Inside a pole the product runs column by column, y[row] += val * x[col], i.e. every
 * output row is accumulated in ascending global column order, multiply-then-add: the same per-row
 * summation order as Julia's CSC column scatter on the assembled D_d (src/pdes.jl:63 `RHS*x`).
 * Poles are independent (they write disjoint rows), so items run in parallel (OpenMP). */
int gsgo_apply_D_poles(int D, int k, int n, int scheme, int d, const int64_t* Hcolptr, const int64_t* Hrowval,
                       const double* Hnzval, const double* x, double* y) {
    if (D > MAXD) return -1;
    index_set* Sp = cached_index_set(D, k, n, scheme);
    if (!Sp) return -1;
    const index_set S = *Sp;
    const int dd = d - 1;
    int64_t radix[MAXD];
    radix[0] = 1;
    for (int i = 1; i < D; ++i) radix[i] = radix[i - 1] * (n + 1);
    const int64_t mstride_d = ipow(k, dd);
    const int64_t nmode_other = S.kD / k;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b0 = 0; b0 < S.nblocks; ++b0) {
        const int* lv = S.level + (int64_t)b0 * D;
        if (lv[dd] != 0) continue;
        int sum_other = 0;
        int64_t tkey = 0;
        for (int i = 0; i < D; ++i) if (i != dd) { sum_other += lv[i]; tkey += radix[i] * lv[i]; }
        const int p = scheme == 1 ? n : n - sum_other;
        const int64_t Np = (int64_t)k << p;
        int C[MAXD];
        int64_t stride_lo = 1, stride_hi = 1;
        for (int i = 0; i < D; ++i) C[i] = cells_of(lv[i]);
        for (int i = 0; i < dd; ++i) stride_lo *= C[i];
        for (int i = dd + 1; i < D; ++i) stride_hi *= C[i];
        int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * Np);
        double* xp = (double*)malloc(sizeof(double) * Np);
        double* yp = (double*)malloc(sizeof(double) * Np);
        if (!idx || !xp || !yp) { rc = -1; free(idx); free(xp); free(yp); continue; }
        for (int64_t hi = 0; hi < stride_hi; ++hi)
            for (int64_t lo = 0; lo < stride_lo; ++lo) {
                /* base index of every 1-D entry (l2, c2, m2) of the pole with other modes = 0 */
                for (int64_t i1 = 0; i1 < Np; ++i1) {
                    const int q = (int)(i1 / k), m2 = (int)(i1 % k);
                    int l2 = 0, c2 = 0;
                    if (q > 0) { l2 = 1; while ((1 << l2) <= q) ++l2; c2 = q - (1 << (l2 - 1)); }
                    const int b2 = S.table[tkey + radix[dd] * l2];
                    const int64_t C2 = cells_of(l2);
                    idx[i1] = S.offset[b2] + (lo + stride_lo * (c2 + C2 * hi)) * S.kD + m2 * mstride_d;
                }
                for (int64_t mo = 0; mo < nmode_other; ++mo) {
                    /* other-mode offset: digits of mo spread around axis d */
                    const int64_t mofs = (mo % mstride_d) + (mo / mstride_d) * mstride_d * k;
                    for (int64_t i1 = 0; i1 < Np; ++i1) { xp[i1] = x[idx[i1] + mofs]; yp[i1] = 0.0; }
                    for (int64_t j1 = 0; j1 < Np; ++j1) {
                        const double xj = xp[j1];
                        for (int64_t pp = Hcolptr[j1]; pp < Hcolptr[j1 + 1]; ++pp) {
                            const int64_t i1 = Hrowval[pp];
                            if (i1 >= Np) break;                 /* rows sorted: outside the principal sub-block */
                            yp[i1] += Hnzval[pp] * xj;
                        }
                    }
                    for (int64_t i1 = 0; i1 < Np; ++i1) y[idx[i1] + mofs] = yp[i1];
                }
            }
        free(idx); free(xp); free(yp);
    }
    return rc;
}

/* y += D_d x with D_d ASSEMBLED column slab by column slab exactly as the reference assembles it
 * (gsgo_assemble_cols = the column loop of src/multidim_derivative.jl:32-55) and applied by CSC column
 * scatter (gsgo_spmv_csc): the full-size stand-in for `D_matrix(D, d, k, n) * x` without holding the
 * 7.5 GB matrix.  y must be zeroed by the caller.  Returns the total nnz, or -1. */
int64_t gsgo_apply_D_assembled(int D, int k, int n, int scheme, int d, const int64_t* Hcolptr,
                               const int64_t* Hrowval, const double* Hnzval, const double* x, double* y,
                               int64_t slab_cols) {
    index_set* Sp = cached_index_set(D, k, n, scheme);
    if (!Sp) return -1;
    const int64_t N = Sp->N;
    int64_t maxcol = 0;
    {   /* widest column of H bounds the slab's nnz */
        const int64_t N1 = (int64_t)k << n;
        for (int64_t j = 0; j < N1; ++j) if (Hcolptr[j + 1] - Hcolptr[j] > maxcol) maxcol = Hcolptr[j + 1] - Hcolptr[j];
    }
    const int64_t cap = slab_cols * maxcol;
    int64_t* colptr = (int64_t*)malloc(sizeof(int64_t) * (slab_cols + 1));
    int64_t* rowval = (int64_t*)malloc(sizeof(int64_t) * cap);
    double* nzval = (double*)malloc(sizeof(double) * cap);
    if (!colptr || !rowval || !nzval) { free(colptr); free(rowval); free(nzval); return -1; }
    int64_t total = 0;
    for (int64_t b = 0; b < N; b += slab_cols) {
        const int64_t e = b + slab_cols < N ? b + slab_cols : N;
        const int64_t nnz = gsgo_assemble_cols(D, k, n, scheme, d, Hcolptr, Hrowval, Hnzval, b, e, colptr, rowval,
                                               nzval, cap);
        if (nnz < 0 || nnz > cap) { total = -1; break; }
        gsgo_spmv_csc(e - b, colptr, rowval, nzval, x + b, y);
        total += nnz;
    }
    free(colptr); free(rowval); free(nzval);
    return total;
}

/* row-parallel CSR product (MKLSparse analogue): y[i] = sum_p val[p] * x[col[p]] */
void gsgo_spmv_csr_omp(int64_t nrows, const int64_t* rowptr, const int64_t* col, const double* val,
                       const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nrows; ++i) {
        double s = 0.0;
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) s += val[p] * x[col[p]];
        y[i] = s;
    }
}

/* y += a * x on n entries (the allocating vector arithmetic of the integrator, serial) */
void gsgo_axpy(int64_t n, double a, const double* x, double* y) {
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}

void gsgo_set_threads(int nthreads) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

int gsgo_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
