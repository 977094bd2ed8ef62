// host_setup.cpp -- see host_setup.hpp.  MUST be compiled with -ffp-contract=off.
#include "host_setup.hpp"

#include <algorithm>
#include <cmath>
#include <mutex>

namespace gsg {
namespace {

using Vec = std::vector<double>;
using Basis = std::vector<Vec>;

// <x^i, x^j> on [-1,1] in the {x^i, sgn(x) x^i} coordinate convention
// (src/dg_basis.jl:21-35); n = vector length, first half plain, second half signed.
double product_matrix(int i, int j, int n) {
    const int k = n / 2;
    auto pm1 = [](int e) { return (e & 1) ? -1 : 1; };
    if (i < k && j < k) return double(1 + pm1(i + j)) / double(1 + i + j);
    if (i >= k && j < k) return double(1 - pm1((i - k) + j)) / double(1 + (i - k) + j);
    if (i < k && j >= k) return product_matrix(j, i, n);
    return double(1 + pm1((i - k) + (j - k))) / double(1 + (i - k) + (j - k));
}

// src/dg_basis.jl:37-50
double dot_vv(const Vec& a, const Vec& b) {
    double value = 0.0;
    const int n = (int)a.size();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (a[i] == 0 || b[j] == 0) continue;
            value += product_matrix(i, j, n) * a[i] * b[j];
        }
    return value;
}

// src/dg_basis.jl:52-60 : against the monomial x^(j1-1), j1 1-based
double dot_vmono(const Vec& a, int j1) {
    double value = 0.0;
    const int n = (int)a.size();
    for (int i = 0; i < n; ++i) value += product_matrix(i, j1 - 1, n) * a[i];
    return value;
}

void sub_scaled(Vec& y, double a, const Vec& x) {  // y -= a * x (two roundings per entry)
    for (size_t i = 0; i < y.size(); ++i) {
        const double t = a * x[i];
        y[i] = y[i] - t;
    }
}

void normalise(Vec& y) {
    const double nrm = std::sqrt(dot_vv(y, y));
    for (double& e : y) e = e / nrm;
}

// src/dg_basis.jl:68-85 (only the first n/2 vectors are processed)
Basis gram_schmidt(const Basis& Q0) {
    const int k = (int)Q0[0].size() / 2;
    Basis Q = Q0;
    for (int i = 0; i < k; ++i) {
        for (int j = 0; j < i; ++j) {
            const double proj = dot_vv(Q0[i], Q[j]) / dot_vv(Q[j], Q[j]);
            sub_scaled(Q[i], proj, Q[j]);
        }
        normalise(Q[i]);
    }
    return Q;
}

// src/dg_basis.jl:88-94
Basis legendre(int k) {
    Basis Q(k + 1, Vec(2 * (k + 1), 0.0));
    for (int j = 0; j <= k; ++j) Q[j][j] = 1.0;
    return gram_schmidt(Q);
}

// src/dg_basis.jl:104-120
Basis orthogonalize_1(const Basis& Q0) {
    const int k = (int)Q0[0].size() / 2;
    Basis Q = Q0;
    const Basis P = legendre(k - 1);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const double proj = dot_vv(Q0[i], P[j]) / dot_vv(P[j], P[j]);
            sub_scaled(Q[i], proj, P[j]);
        }
    return Q;
}

// src/dg_basis.jl:129-146 (fi comes from the INPUT set, as in the reference)
Basis orthogonalize_2(const Basis& Q0) {
    const int k = (int)Q0[0].size() / 2;
    Basis Q = Q0;
    for (int i = 1; i <= k - 1; ++i) {
        const Vec fi = Q0[i - 1];
        for (int j = i + 1; j <= k; ++j) {
            const double a = dot_vmono(Q[j - 1], k + i) / dot_vmono(fi, k + i);
            sub_scaled(Q[j - 1], a, fi);
        }
    }
    return Q;
}

// src/dg_basis.jl:154-169
Basis gram_schmidt_rev(const Basis& Q0) {
    const int n = (int)Q0[0].size(), k = n / 2;
    Basis Q(k, Vec(n, 0.0));
    for (int i = k; i >= 1; --i) {
        const Vec fi = Q0[i - 1];
        Q[i - 1] = fi;
        for (int j = i + 1; j <= k; ++j) {
            const double proj = dot_vv(fi, Q[j - 1]) / dot_vv(Q[j - 1], Q[j - 1]);
            sub_scaled(Q[i - 1], proj, Q[j - 1]);
        }
        normalise(Q[i - 1]);
    }
    return Q;
}

// src/dg_basis.jl:185-191
Basis dg_basis(int k) {
    Basis Q(k, Vec(2 * k, 0.0));
    for (int j = 0; j < k; ++j) Q[j][k + j] = 1.0;
    return gram_schmidt_rev(orthogonalize_2(orthogonalize_1(Q)));
}

// src/derivative_matrix_elements.jl:80-94
Vec symbolic_diff(const Vec& v) {
    const int n = (int)v.size(), k = n / 2;
    Vec ans(n, 0.0);
    for (int i = 1; i <= n; ++i) {
        if (i < k) ans[i - 1] = i * v[i];
        else if (i > k && i < 2 * k) ans[i - 1] = (i - k) * v[i];
    }
    return ans;
}

// Gauss-Legendre rule on [-1,1]; stands in for hquadrature (G7K15) on the polynomial
// integrands of hier2pos (degree <= 2k-2 per cell, so any rule with >= k points is exact).
struct GaussRule {
    std::vector<double> x, w;
    explicit GaussRule(int npts) : x(npts), w(npts) {
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < npts; ++i) {
            double z = std::cos(pi * (i + 0.75) / (npts + 0.5));
            double pp = 0.0;
            for (int it = 0; it < 100; ++it) {
                double p1 = 1.0, p2 = 0.0;
                for (int j = 1; j <= npts; ++j) {
                    const double p3 = p2;
                    p2 = p1;
                    p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
                }
                pp = npts * (z * p1 - p2) / (z * z - 1.0);
                const double z1 = z;
                z = z1 - p1 / pp;
                if (std::fabs(z - z1) < 1e-16) break;
            }
            x[npts - 1 - i] = z;
            w[npts - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
        }
    }
};

const GaussRule& gauss12() {
    static const GaussRule g(12);
    return g;
}

Csc from_columns(int64_t m, int64_t n, std::vector<std::map<int64_t, double>>& cols, bool drop_zeros) {
    Csc C;
    C.m = m;
    C.n = n;
    C.colptr.assign(n + 1, 0);
    for (int64_t j = 0; j < n; ++j) {
        for (auto& kv : cols[j]) {
            if (drop_zeros && kv.second == 0.0) continue;
            C.rowval.push_back(kv.first);
            C.nzval.push_back(kv.second);
        }
        C.colptr[j + 1] = (int64_t)C.rowval.size();
    }
    return C;
}

// src/derivative_matrix_elements.jl:101-112 and src/1d_derivative.jl:21-44:
// block-diagonal volume term; off-diagonal cell pairs are exact zeros in the reference.
Csc volume_matrix(int k, int level) {
    const auto& L = leg_coeffs();
    std::vector<double> blk(k * k);
    for (int m1 = 0; m1 < k; ++m1)
        for (int m2 = 0; m2 < k; ++m2)
            blk[m1 * k + m2] = double(int64_t(1) << (level + 1)) * dot_vv(L[m1], symbolic_diff(L[m2]));
    const int64_t nc = int64_t(1) << level, N = nc * k;
    std::vector<std::map<int64_t, double>> cols(N);
    for (int64_t c = 0; c < nc; ++c)
        for (int m1 = 0; m1 < k; ++m1)
            for (int m2 = 0; m2 < k; ++m2) {
                const double val = blk[m1 * k + m2];
                if (std::fabs(val) > 1.0e-15) cols[c * k + m2][c * k + m1] = val;
            }
    return from_columns(N, N, cols, false);
}

// src/1d_derivative.jl:52-74, alpha = 0; cells 1-based.
double lf_element(int level, int cell1, int mode1, int cell2, int mode2) {
    const double scale = double(int64_t(1) << level);
    const double point1 = double(cell2 - 1) / scale;
    const double point2 = double(cell2) / scale;
    const double tiny = 5.0e-16;
    double left1 = basis_pos(level, cell1, mode1, point1 - tiny);
    const double right1 = basis_pos(level, cell1, mode1, point1 + tiny);
    const double left2 = basis_pos(level, cell1, mode1, point2 - tiny);
    double right2 = basis_pos(level, cell1, mode1, point2 + tiny);
    if (cell2 == (1 << level)) right2 = basis_pos(level, cell1, mode1, 0.0 + tiny);
    if (cell2 == 1) left1 = basis_pos(level, cell1, mode1, 1.0 - tiny);
    const double alpha = 0.0;
    const double LF1 = 0.5 * (left1 + right1) + alpha * (right1 - left1);
    const double LF2 = 0.5 * (left2 + right2) + alpha * (right2 - left2);
    const double val1 = basis_pos(level, cell2, mode2, point1 + tiny);
    const double val2 = basis_pos(level, cell2, mode2, point2 - tiny);
    return LF2 * val2 - LF1 * val1;
}

// src/1d_derivative.jl:76-100.  Only periodic-neighbour cell pairs are visited: for any other
// pair every one-sided evaluation of `basis` falls outside its support (array2poly returns
// 0.0), the element is an exact 0.0 and the reference's `abs(val) > 1.0e-15` drops it.
Csc lf_matrix(int k, int level) {
    const int nc = 1 << level;
    const int64_t N = int64_t(nc) * k;
    std::vector<std::map<int64_t, double>> cols(N);
    for (int cell1 = 1; cell1 <= nc; ++cell1) {
        int cand[3] = {(cell1 - 2 + nc) % nc + 1, cell1, cell1 % nc + 1};
        std::sort(cand, cand + 3);
        const int ncand = int(std::unique(cand, cand + 3) - cand);
        for (int mode1 = 1; mode1 <= k; ++mode1)
            for (int ci = 0; ci < ncand; ++ci)
                for (int mode2 = 1; mode2 <= k; ++mode2) {
                    const double val = lf_element(level, cell1, mode1, cand[ci], mode2);
                    if (std::fabs(val) > 1.0e-15)
                        cols[int64_t(cand[ci] - 1) * k + (mode2 - 1)][int64_t(cell1 - 1) * k + (mode1 - 1)] = val;
                }
    }
    return from_columns(N, N, cols, false);
}

}  // namespace

// ------------------------------------------------------------------------------------------
const std::vector<std::vector<double>>& leg_coeffs() {
    static const Basis L = legendre(K_MAX);
    return L;
}

const std::vector<std::vector<double>>& dg_coeffs(int k) {
    static Basis tab[K_MAX + 1];
    static std::once_flag once[K_MAX + 1];
    std::call_once(once[k], [k] { tab[k] = dg_basis(k); });
    return tab[k];
}

double array2poly(const double* v, int n, double x) {
    if (std::fabs(x) > 1) return 0.0;
    const int k = n / 2;
    const bool neg = std::signbit(x);
    double s = 0.0;
    for (int i = k - 1; i >= 0; --i) {
        s = s * x;
        const double t = neg ? -v[i + k] : v[i + k];
        s = s + (v[i] + t);
    }
    return s;
}

double LegendreP(int kk, double x) {
    const auto& L = leg_coeffs();
    return array2poly(L[kk].data(), (int)L[kk].size(), x);
}

double h_fn(int k, int mode, double x) {
    const auto& T = dg_coeffs(k);
    return array2poly(T[mode - 1].data(), 2 * k, x);
}

double basis_pos(int level, int cell, int mode, double x) {
    const double t = double(int64_t(1) << level) * x - double(cell - 1);
    const double lg = std::sqrt(2.0) * LegendreP(mode - 1, 2 * t - 1);
    return lg * std::pow(2.0, level / 2.0);
}

double v_fn(int k, int level, int cell, int mode, double x) {
    if (level == 0) return LegendreP(mode - 1, 2 * x - 1) * std::sqrt(2.0);
    const double s = double(int64_t(1) << level);
    return h_fn(k, mode, s * x - double(2 * cell - 1)) * std::sqrt(1.0 * s);
}

int64_t cell_index(double x, int l) {
    if (l <= 1) return 1;
    if (x >= 1) return int64_t(1) << (l - 1);
    return 1 + (int64_t)std::floor(double(int64_t(1) << (l - 1)) * x);
}

Csc transpose(const Csc& A) {
    std::vector<std::map<int64_t, double>> cols(A.m);
    for (int64_t j = 0; j < A.n; ++j)
        for (int64_t p = A.colptr[j]; p < A.colptr[j + 1]; ++p) cols[A.rowval[p]][j] = A.nzval[p];
    return from_columns(A.n, A.m, cols, false);
}

// Gustavson product in the summation order of Julia 1.0's SparseArrays.spmatmul: for column i
// of B, stored B[j,i] ascending in j, scatter A[:,j]*B[j,i]; structural fill is kept.
Csc spmatmul(const Csc& A, const Csc& B) {
    Csc C;
    C.m = A.m;
    C.n = B.n;
    C.colptr.assign(B.n + 1, 0);
    std::vector<double> acc(A.m, 0.0);
    std::vector<int64_t> mark(A.m, -1);
    std::vector<int64_t> touched;
    for (int64_t i = 0; i < B.n; ++i) {
        touched.clear();
        for (int64_t jp = B.colptr[i]; jp < B.colptr[i + 1]; ++jp) {
            const int64_t j = B.rowval[jp];
            const double b = B.nzval[jp];
            for (int64_t kp = A.colptr[j]; kp < A.colptr[j + 1]; ++kp) {
                const int64_t r = A.rowval[kp];
                const double t = A.nzval[kp] * b;
                if (mark[r] != i) {
                    mark[r] = i;
                    acc[r] = t;
                    touched.push_back(r);
                } else {
                    acc[r] = acc[r] + t;
                }
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int64_t r : touched) {
            C.rowval.push_back(r);
            C.nzval.push_back(acc[r]);
        }
        C.colptr[i + 1] = (int64_t)C.rowval.size();
    }
    return C;
}

Csc hier2pos(int k, int max_level) {
    const int nc = 1 << max_level;
    const int64_t N = int64_t(nc) * k;
    const GaussRule& G = gauss12();
    std::vector<std::map<int64_t, double>> cols(N);
    int64_t j = 0;
    for (int level = 0; level <= max_level; ++level) {
        const int ncell_l = 1 << std::max(0, level - 1);
        const int width = nc / ncell_l;  // finest cells under one support
        for (int cell = 1; cell <= ncell_l; ++cell)
            for (int mode = 1; mode <= k; ++mode, ++j) {
                // pos_vcoeffs_DG(k, max_level, v(k, level, cell, mode)) restricted to the support
                for (int pc = (cell - 1) * width + 1; pc <= cell * width; ++pc)
                    for (int pm = 1; pm <= k; ++pm) {
                        const double a = double(pc - 1) / nc, b = double(pc) / nc;
                        const double half = 0.5 * (b - a), mid = 0.5 * (b + a);
                        double s = 0.0;
                        for (size_t q = 0; q < G.x.size(); ++q) {
                            const double x = mid + half * G.x[q];
                            const double f = basis_pos(max_level, pc, pm, x) * v_fn(k, level, cell, mode, x);
                            s += G.w[q] * f;
                        }
                        s = s * half;
                        if (std::fabs(s) > 1.0e-12) cols[j][int64_t(pc - 1) * k + (pm - 1)] = s;
                    }
            }
    }
    return from_columns(N, N, cols, false);
}

Csc periodic_pos_DLF_matrix(int k, int max_level) {
    const Csc Dm = volume_matrix(k, max_level);
    const Csc LF = lf_matrix(k, max_level);
    // -D + LF ; Julia's sparse map drops exact zeros of the result
    std::vector<std::map<int64_t, double>> cols(Dm.n);
    for (int64_t j = 0; j < Dm.n; ++j) {
        for (int64_t p = Dm.colptr[j]; p < Dm.colptr[j + 1]; ++p) cols[j][Dm.rowval[p]] = -Dm.nzval[p];
        for (int64_t p = LF.colptr[j]; p < LF.colptr[j + 1]; ++p) {
            auto it = cols[j].find(LF.rowval[p]);
            if (it == cols[j].end()) cols[j][LF.rowval[p]] = LF.nzval[p];
            else it->second = it->second + LF.nzval[p];
        }
    }
    return transpose(from_columns(Dm.m, Dm.n, cols, true));
}

Csc periodic_hier_DLF_matrix(int k, int max_level) {
    const Csc Q = hier2pos(k, max_level);
    const Csc A = periodic_pos_DLF_matrix(k, max_level);
    return spmatmul(transpose(Q), spmatmul(A, Q));
}

void dense_square(const double* A, int n, double* C) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) {
                const double prod = A[(size_t)i * n + k] * A[(size_t)k * n + j];
                acc = acc + prod;
            }
            C[(size_t)i * n + j] = acc;
        }
}

// ------------------------------------------------------------------------------------------
bool IndexSet::build(int D_, int k_, int n_, int scheme_) {
    D = D_; k = k_; n = n_; scheme = scheme_;
    blocks.clear();
    by_level.clear();
    kD = 1;
    for (int i = 0; i < D; ++i) kD *= k;
    kDp = (kD + 1) & ~int64_t(1);
    std::vector<int> lv(D, 0);
    int64_t off = 0, poff = 0;
    ncells_total = 0;
    N = 0;
    Npad = 0;
    int sum = 0;
    while (true) {
        {
            Block b;
            b.level = lv;
            b.cells.resize(D);
            b.ncells = 1;
            for (int i = 0; i < D; ++i) {
                b.cells[i] = 1 << std::max(0, lv[i] - 1);
                b.ncells *= b.cells[i];
            }
            b.offset = off;
            b.poffset = poff;
            off += b.ncells * kD;
            poff += b.ncells * kDp;
            ncells_total += b.ncells;
            by_level[lv] = (int)blocks.size();
            blocks.push_back(std::move(b));
            if ((int64_t)blocks.size() > MAX_BLOCKS || off > MAX_DOFS) return false;   // refuse absurd sizes early
        }
        // first dimension fastest (Julia CartesianIndices order).  Sparse scheme: only tuples with sum <= n
        // are visited -- once ++lv[i] (all lower dimensions already reset to 0) breaks the cutoff, every larger
        // lv[i] does too, so the odometer carries; the visiting order of the kept tuples is unchanged.
        int i = 0;
        while (i < D) {
            ++lv[i];
            ++sum;
            if (lv[i] <= n && (scheme == 1 || sum <= n)) break;
            sum -= lv[i];
            lv[i] = 0;
            ++i;
        }
        if (i == D) break;
    }
    N = off;
    Npad = poff;
    return true;
}

int64_t get_size(int D, int k, int n, int scheme) {
    IndexSet S;
    if (!S.build(D, k, n, scheme)) return -1;
    return S.N;
}

void tensor_construct(const IndexSet& S, const double* const* v1d, double* out) {
    const int D = S.D, k = S.k;
    std::vector<int> m(D), c(D);
    for (const Block& b : S.blocks) {
        double* dst = out + b.offset;
        std::fill(c.begin(), c.end(), 0);
        for (int64_t ci = 0; ci < b.ncells; ++ci) {
            std::fill(m.begin(), m.end(), 0);
            for (int64_t e = 0; e < S.kD; ++e) {
                double val = 1.0;
                for (int d = 0; d < D; ++d) {
                    const int l = b.level[d];
                    const int64_t base = l == 0 ? 0 : (int64_t(1) << (l - 1));
                    val *= v1d[d][k * (base + c[d]) + m[d]];
                }
                *dst++ = val;
                int i = 0;
                while (i < D && ++m[i] == k) m[i++] = 0;
            }
            int i = 0;
            while (i < D && ++c[i] == b.cells[i]) c[i++] = 0;
        }
    }
}

}  // namespace gsg
