// host_setup.hpp -- CPU-side setup of the sparse-grid DG operator: index set / vector layout,
// Alpert multiwavelet basis tables, and the 1-D hierarchical Lax-Friedrichs derivative matrix
// H = periodic_DLF_matrix(k, n).  These are the quantities the Julia host computes itself and
// hands over the C ABI; they are provided here so the library is self-contained for non-Julia
// hosts.  Compiled with -ffp-contract=off: the `tiny = 5.0e-16` one-sided-limit arithmetic of
// the reference (src/1d_derivative.jl:52-74) must be reproduced without FMA contraction.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

namespace gsg {

constexpr int K_MAX = 10;        // src/1d_dg_functions.jl:7
constexpr int N_MAX_LEVEL = 16;  // library limit on n
constexpr int64_t MAX_BLOCKS = 4000000;            // library limits on the index set (IndexSet::build fails beyond)
constexpr int64_t MAX_DOFS = int64_t(1) << 40;

struct Csc {                      // 0-based compressed sparse column
    int64_t m = 0, n = 0;
    std::vector<int64_t> colptr, rowval;
    std::vector<double> nzval;
    int64_t nnz() const { return colptr.empty() ? 0 : colptr.back(); }
};

// ---- basis ------------------------------------------------------------------------------
const std::vector<std::vector<double>>& leg_coeffs();        // src/1d_dg_functions.jl:35
const std::vector<std::vector<double>>& dg_coeffs(int k);    // src/1d_dg_functions.jl:52-62
double array2poly(const double* v, int n, double x);         // src/1d_dg_functions.jl:15-28
double LegendreP(int kk, double x);                          // :38-41
double h_fn(int k, int mode, double x);                      // :66-69
double basis_pos(int level, int cell, int mode, double x);   // :91-93
double v_fn(int k, int level, int cell, int mode, double x); // src/dg_methods.jl:27-36
int64_t cell_index(double x, int l);                         // src/dg_methods.jl:70-79

// ---- 1-D operator --------------------------------------------------------------------------
Csc hier2pos(int k, int max_level);                          // src/1d_dg_functions.jl:242-263
Csc periodic_pos_DLF_matrix(int k, int max_level);           // src/1d_derivative.jl:108-111
Csc periodic_hier_DLF_matrix(int k, int max_level);          // src/1d_derivative.jl:113-117
Csc spmatmul(const Csc& A, const Csc& B);                    // SparseArrays.spmatmul (Gustavson)
Csc transpose(const Csc& A);
// C = A * A for a dense row-major n x n block, inner index ascending, separate multiply and add (the order and
// rounding of SparseArrays' A*B on the stored entries; zeros add nothing)   // src/multidim_derivative.jl:76
void dense_square(const double* A, int n, double* C);

// ---- index set / layout -------------------------------------------------------------------
struct Block {
    std::vector<int> level;     // 0-based levels, size D
    std::vector<int> cells;     // cells per dim: 1 << max(0, level-1)
    int64_t offset = 0;         // offset of the block in the reference vector layout
    int64_t poffset = 0;        // offset in the device-internal layout (cells padded to kDp)
    int64_t ncells = 0;
};

struct IndexSet {               // src/dg_vmethods.jl:35-142 (the D2V / V2Dref loop order)
    int D = 0, k = 0, n = 0, scheme = 0;
    int64_t N = 0, kD = 0;
    int64_t kDp = 0, Npad = 0;  // device layout: every k^D multi-cell padded to an even length
    int64_t ncells_total = 0;   // (16-byte alignment for TMA bulk copies and 128-bit accesses)
    std::vector<Block> blocks;
    std::map<std::vector<int>, int> by_level;
    bool build(int D, int k, int n, int scheme);
};

int64_t get_size(int D, int k, int n, int scheme);           // src/dg_vmethods.jl:35-45

// tensor_construct on vector-layout inputs                  // src/tensor_construct.jl:19-63
void tensor_construct(const IndexSet& S, const double* const* v1d, double* out);

}  // namespace gsg
