/* gsg_b200.h -- C ABI of libgsgb200.so: the B200-native drop-in for the hot path of
 * GalerkinSparseGrids.jl (apply the sparse-grid DG derivative / gradient / Laplacian
 * operator inside every Runge-Kutta right-hand side, plus batched reconstruct_DG).
 *
 * The reference has no FFI of its own (it is pure Julia); the swap points are the Julia
 * dispatch sites listed per function below (paths relative to the reference root).
 * INTEGRATION.md shows the `ccall` shim a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative gsg_status on failure;
 *     gsg_last_error() returns a thread-local, NUL-terminated message.
 *   - no exceptions cross the boundary; host pointers are never retained past the call.
 *   - vectors use the reference's vector-hierarchical layout (src/dg_vmethods.jl:48-73).
 *   - sparse matrices cross as Julia's SparseMatrixCSC{Float64,Int64} fields
 *     {m, n, colptr[n+1], rowval[nnz], nzval[nnz]} with 1-BASED Int64 indices
 *     (the five fields the reference itself serialises, src/pdes.jl:149-155).
 *   - `scheme`: 0 = "sparse", 1 = "full" (src/schemes.jl:21-27).
 *   - `d` (sweep axis) is 1-BASED as in the reference (src/multidim_derivative.jl:61).
 *   - *_dev entry points take device pointers and a cudaStream_t (as void*) and are
 *     asynchronous on that stream; the host-pointer entry points synchronise before
 *     returning.  There is NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with GSG_ERR_CUDA.
 */
#ifndef GSG_B200_H
#define GSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GSG_OK = 0,
    GSG_ERR_ARG = -1,      /* invalid argument (mirrors Julia's ArgumentError / DomainError) */
    GSG_ERR_CUDA = -2,     /* CUDA runtime failure or no device                                */
    GSG_ERR_ALLOC = -3,    /* host or device allocation failure                               */
    GSG_ERR_UNSUPPORTED = -4
} gsg_status;

typedef struct gsg_plan gsg_plan;      /* operator plan: index set + device tables + 1-D matrices */
typedef struct gsg_csr gsg_csr;        /* resident generic sparse matrix (cross-check SpMV)        */

/* ---- diagnostics -------------------------------------------------------------------- */
int gsg_version(void);
const char* gsg_last_error(void);
/* writes "name;sm_count;cc_major.cc_minor;global_mem_bytes" */
int gsg_device_info(int device, char* buf, size_t buflen);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t gsg_launch_count(void);

/* ---- host-side setup mirrors (CPU only; these are what the Julia host computes itself) -- */
/* get_size(Val(D), k, n, Val(scheme))                        src/dg_vmethods.jl:35-45 */
int gsg_get_size(int D, int k, int n, int scheme, int64_t* size_out);
/* v(k, level, cell, mode, x) for an array of x; level 0-based, cell/mode 1-based
 *                                                             src/dg_methods.jl:27-36 */
int gsg_basis_v(int k, int level, int cell, int mode, const double* x, int64_t npts, double* out);
/* cell_index(x, l)                                            src/dg_methods.jl:70-79 */
int gsg_cell_index(double x, int level, int64_t* cell_out);
/* leg_coeffs (K_max+1 rows of 2(K_max+1)) and dg_coeffs[k] (k rows of 2k)
 *                                                             src/1d_dg_functions.jl:35,52-62 */
int gsg_basis_tables(int k, double* leg_out, double* dg_out);
/* periodic_DLF_matrix(k, n; basis="hier"|"pos")               src/1d_derivative.jl:136-148
 * two-call pattern: pass nzval == NULL to query nnz.  basis: 0 = hier, 1 = pos. */
int gsg_dlf_matrix(int k, int n, int basis, int64_t* nnz_inout,
                   int64_t* colptr, int64_t* rowval, double* nzval);
/* Structural block pattern of periodic_DLF_matrix(k, n) (src/1d_derivative.jl:113-117; k x k blocks indexed by
 * the 1-D cells q = 0 (level 0), 2^(l-1) + c (level l >= 1)): out[q * 2^n + r] = 1 iff the closed supports of
 * cells q and r intersect or touch periodically.  out has 4^n bytes. */
int gsg_block_pattern(int n, unsigned char* out);
/* tensor_construct(D, k, n, [v_1..v_D]; scheme) on 1-D coefficient vectors of length k*2^n,
 * result in vector layout                                     src/tensor_construct.jl:19-63 */
int gsg_tensor_construct(int D, int k, int n, int scheme, const double* const* vcoeffs_1d,
                         double* out);

/* ---- plan ----------------------------------------------------------------------------- */
/* (declared further down: gsg_tensor_construct_dev builds tensor_construct's result directly in a device vector) */
/* Replaces the construction of grad_matrix / laplacian_matrix (src/pdes.jl:59,141;
 * src/multidim_derivative.jl:19-79).  H = periodic_DLF_matrix(k, n) is handed over verbatim
 * (SURVEY.md headline fact 4: values as stored, noise entries included). */
int gsg_plan_create(int D, int k, int n, int scheme,
                    int64_t H_n, const int64_t* H_colptr, const int64_t* H_rowval,
                    const double* H_nzval, int device, gsg_plan** plan_out);
int gsg_plan_destroy(gsg_plan* plan);
int gsg_plan_size(const gsg_plan* plan, int64_t* size_out);
/* Device vectors live in the DEVICE LAYOUT: the reference layout with every k^D multi-cell
 * padded to an even number of doubles (16-byte aligned cells for TMA bulk copies).  dev_size
 * is that padded length; pack/unpack convert device-resident vectors (asynchronously on the
 * plan's stream); the padding slots of a device vector must be zero-initialised. */
int gsg_plan_dev_size(const gsg_plan* plan, int64_t* size_out);
/* tensor_construct(D, k, n, [v_1..v_D]) expanded ON THE DEVICE into out_dev (device layout; padding slots are left
 * untouched: zero-initialise the vector once).  vcoeffs_1d: D HOST vectors of length k*2^n.  src/tensor_construct.jl:19-63 */
int gsg_tensor_construct_dev(gsg_plan* plan, const double* const* vcoeffs_1d, double* out_dev);
int gsg_pack_dev(gsg_plan* plan, const double* ref_layout_dev, double* dev_layout_dev);
int gsg_unpack_dev(gsg_plan* plan, const double* dev_layout_dev, double* ref_layout_dev);
/* run all subsequent work of this plan on `stream` (cudaStream_t); NULL = the plan's own */
int gsg_plan_set_stream(gsg_plan* plan, void* stream);
int gsg_plan_sync(gsg_plan* plan);
/* Execution path of the operator applies.  mode 0: the tiled class kernels (streaming / long-pole) always;
 * 1: the flat kernel (one launch per right-hand side, all directions) whenever it supports the plan (k <= 5,
 * no partition); 2 (default): automatic -- flat for small index sets (N * D <= 6e6: BASELINE configs 2 and 3,
 * where a step is launch-latency bound), tiled above.  The environment variable GSG_FLAT sets the initial mode. */
int gsg_plan_set_flat(gsg_plan* plan, int mode);
/* CPU-side check of the row-tile kernel's tile program (no device needed) for pole class p of the library's own
 * H = periodic_DLF_matrix(k, n) and multi-cells of k^D doubles; budget_bytes = shared memory of a CTA (barriers + 2 nrg
 * staging cells + x cells + records), nrg = row groups.  Two-call pattern: NULL outputs return counts_out = {tiles,
 * groups, blob bytes}.  tiles_out: 48 int32 per tile {nx, rec_ofs, rec_bytes, grp0, rg_end[4], xq[40]}; groups_out:
 * 8 int32 per row group {q[4], rofs, nrec, partial mask, 0}; blob_out: the records, {int32 x-cell byte offset, int32
 * row mask | arrival barrier << 8} followed by one k x k block (row-major doubles) per mask bit. */
int gsg_debug_rowtile_program(int D, int k, int n, int p, int64_t budget_bytes, int nrg, int32_t* tiles_out,
                              int32_t* groups_out, unsigned char* blob_out, int64_t* counts_out);
/* CPU-side check of the row-tile kernel's lane order: nslots (a multiple of 32, >= PI) in-cell pole offsets
 * a + k*A*b, dealt so that the 16 lanes of a half-warp fall into different 8-byte shared-memory banks; padding
 * lanes are stored as ~offset. */
int gsg_debug_rowtile_pole_order(int k, int A, int PI, int nslots, int32_t* out);
/* human-readable summary: which kernel serves which pole classes in every direction ("kind(p, tiles)") */
int gsg_plan_describe(const gsg_plan* plan, char* buf, size_t buflen);
/* 1 if the operator applies of this plan currently take the flat kernel */
int gsg_plan_flat_active(const gsg_plan* plan, int* active_out);
/* CPU-side check of the flat path's tables (no device needed): pole groups of direction d (1-based) and, for
 * every multi-cell of the device layout, its {group, item, 1-D cell} along d.  groups_out: ngroups rows of
 * 20 int64 {base[0..16], p, S, nitems}; cells_out: ncells rows of 3 int32.  Pass NULL outputs to query the
 * counts (ngroups_out, ncells_out). */
int gsg_debug_flat_tables(int D, int k, int n, int scheme, int d, int64_t* groups_out, int32_t* cells_out,
                          int64_t* ngroups_out, int64_t* ncells_out);

/* ---- operator apply, host vectors (drop-in for `A*x`) ------------------------------------ */
/* y = D_d * x             `Ds[d] * f_modal`  src/pdes.jl:179-180, `D_ops[j]*u` :268 */
int gsg_apply_D(gsg_plan* plan, int d, const double* x, double* y);
/* y = sum_d a[d] * D_d x  (gradient combination; a has D entries) */
int gsg_apply_grad(gsg_plan* plan, const double* a, const double* x, double* y);
/* y = (sum_d D_d * D_d) x   `laplacian_matrix(D,k,n) * x`  src/multidim_derivative.jl:71-79: the short pole
 * classes apply the explicit square of their block (the reference's form), the long ones D_d (D_d x) */
int gsg_apply_laplacian(gsg_plan* plan, const double* x, double* y);

/* ---- operator apply, device vectors (DEVICE LAYOUT, length gsg_plan_dev_size) ---------------- */
/* y = alpha * D_d x + beta * y   (beta == 0: y is not read); x and y must not alias */
int gsg_apply_D_dev(gsg_plan* plan, int d, double alpha, const double* x_dev, double beta,
                    double* y_dev);
/* y = sum_d a[d] D_d x */
int gsg_apply_grad_dev(gsg_plan* plan, const double* a, const double* x_dev, double* y_dev);
int gsg_apply_laplacian_dev(gsg_plan* plan, const double* x_dev, double* y_dev, double* tmp_dev);
/* y = beta * y + sum of c[d-1] * D_d x over the directions d whose bit (d-1) is set in dmask (beta 0 or 1);
 * direction pairs inside the mask are swept fused.  Used by the multi-GPU driver (local / partition groups). */
int gsg_apply_dirs_dev(gsg_plan* plan, const double* c, unsigned dmask, double beta, const double* x_dev,
                       double* y_dev);

/* RK4 driver variant: 0 (default) = automatic -- the linear right-hand sides below are advanced in the
 * Taylor form u += dt L u + dt^2/2 L^2 u + dt^3/6 L^3 u + dt^4/24 L^4 u (same four operator applies,
 * one combine pass; identical to the staged form up to rounding); 1 = always the staged form. */
int gsg_plan_set_rk4_mode(gsg_plan* plan, int mode);

/* ---- fused fixed-step RK4 evolutions, state resident on device ------------------------- */
/* u' = -sum_d a[d] D_d u  (the operator vlasov_evolve applies, src/pdes.jl:179-180;
 * BASELINE config 4).  Classical RK4, stage order in DESIGN.md.  y is updated in place. */
int gsg_rk4_advect(gsg_plan* plan, const double* a, double* y, double dt, int64_t nsteps);
int gsg_rk4_advect_dev(gsg_plan* plan, const double* a, double* y_dev, double dt, int64_t nsteps);
/* [u; v]' = [v; L u]   (`wave_data`, src/pdes.jl:22-49; RHS closure :63) */
int gsg_rk4_wave(gsg_plan* plan, double* u, double* v, double dt, int64_t nsteps);
int gsg_rk4_wave_dev(gsg_plan* plan, double* u_dev, double* v_dev, double dt, int64_t nsteps);
/* E = sum_d |D_d u|^2 + |udot|^2     energy_func, src/pdes.jl:258-273 */
int gsg_energy(gsg_plan* plan, const double* u, const double* udot, double* energy_out);

/* ---- building blocks for host-driven / multi-GPU time stepping (device layout) ------------ */
/* ---- multi-GPU block partition -------------------------------------------
 * nranks = 2^b ranks; dimension D-j (j < b) splits the multi-level blocks into {level == 0} (rank bit j
 * = 1) and {level >= 1} (bit 0).  After this call the plan's sweeps cover only the pole groups this rank
 * computes: directions 1..D-b are local; along a partition dimension the poles that straddle a rank pair
 * are swept by the bit-0 rank once it has received the pair's level-0 blocks (list kind 1), and that
 * rank returns its contribution for those blocks.  Vectors stay full-length (device layout). */
int gsg_plan_set_partition(gsg_plan* plan, int rank, int nranks);
/* kind 0: blocks owned by this rank; kind 1: level_d == 0 blocks exchanged with *partner_out along the
 * partition dimension d (1-based; count 0 if d is not one).  Offsets/sizes in doubles of the device layout;
 * call with offsets == NULL for the count. */
int gsg_plan_partition_blocks(gsg_plan* plan, int kind, int d, int64_t* offsets, int64_t* sizes, int64_t* count,
                              int* partner_out);
/* u += c1 v1 + c2 v2 + c3 v3 + c4 v4 on the listed multi-cells (indices in units of the padded cell) */
int gsg_rk4_taylor_cells_dev(gsg_plan* plan, const int* cells_dev, int64_t ncells, double* u, const double* v1,
                             const double* v2, const double* v3, const double* v4, double c1, double c2, double c3,
                             double c4);

/* ---- adaptive integrators: drop-in for ODE.jl's ode45 / ode78 --------------------------------------------
 * `ode45((t,x)->RHS*x, y0, [t0,t1])` / `ode78(...)`  src/pdes.jl:62-68, 113-119, 206-213 (ODE.jl 2.4.0: Dormand-
 * Prince 5(4) `bt_dopri5`, Fehlberg 7(8) `bt_feh78`, `oderk_adapt` step control; defaults reltol 1e-5, abstol 1e-8,
 * 2-norm of the scaled error, maxstep |t1-t0|/2.5).  The state, the stages and all vector arithmetic stay on the
 * device; the host receives one (error, nan) pair per attempted step.  The integrator is a stepper: every call of
 * gsg_ode_step advances to the next ACCEPTED step, so the host rebuilds ODE.jl's (tout, yout) with points=:all by
 * reading the state after each call, or points=:specified with gsg_ode_interp (3rd-order Hermite, as ODE.jl). */
typedef struct gsg_ode gsg_ode;
#define GSG_RHS_ADVECT 0      /* y' = -sum_d a[d] D_d y          state length N   (the operator of src/pdes.jl:179-180) */
#define GSG_RHS_WAVE 1        /* [u; v]' = [v; L u]              state length 2N  (wave_data, src/pdes.jl:22-49)        */
#define GSG_RHS_CSR 2         /* y' = A y, A a resident gsg_csr  state length A.m (wave_evolve_1D's RHS, src/pdes.jl:109-114) */
#define GSG_RHS_VLASOV 3      /* f' = steprule(t, f)             state length N   (src/pdes.jl:174-192; gsg_ode_create_vlasov) */
/* method: 45 or 78; reltol / abstol <= 0 select ODE.jl's defaults; `a` only for GSG_RHS_ADVECT, `A` only for GSG_RHS_CSR */
int gsg_ode_create(gsg_plan* plan, int rhs_kind, const double* a, gsg_csr* A, int method, double reltol, double abstol,
                   const double* y0_host, double t0, double t1, gsg_ode** out);
int gsg_ode_destroy(gsg_ode* ode);
int gsg_ode_step(gsg_ode* ode, double* t_out, double* dt_out, int* done_out);
int gsg_ode_state(gsg_ode* ode, double* y_host);
int gsg_ode_interp(gsg_ode* ode, double tquery, double* y_host);
int gsg_ode_stats(gsg_ode* ode, int64_t* accepted, int64_t* rejected, int64_t* rhs_evals);

/* ---- Vlasov right-hand side (src/pdes.jl:165-192) -----------------------------------------------------------------
 * steprule(t, f) = n2m * (p2n * (-sum_d v_point[d] .* (n2p*(m2n*(Ds[d]*f))) + sum_d F_point[d] .* (n2p*(m2n*(Ds[D+d]*f)))))
 * on the device: Ds[d]*f are the plan's matrix-free sweeps (plan = the (2D, k, n) operator), the four transform
 * matrices are the ones the host passes to vlasov_evolve (resident gsg_csr handles), the products with v_point /
 * F_point are fused pointwise kernels.  F_point: D host vectors of length N (point basis). */
typedef struct gsg_vlasov gsg_vlasov;
int gsg_vlasov_create(gsg_plan* plan, gsg_csr* m2n, gsg_csr* n2p, gsg_csr* p2n, gsg_csr* n2m,
                      const double* const* F_point, gsg_vlasov** out);
int gsg_vlasov_destroy(gsg_vlasov* v);
int gsg_vlasov_rhs(gsg_vlasov* v, const double* f_modal, double* out);
int gsg_vlasov_v_point(gsg_vlasov* v, int i, double* out);
/* `solver(steprule, f0_modal, range(t0, t1, length = nout))`  src/pdes.jl:206-213 */
int gsg_ode_create_vlasov(gsg_vlasov* v, int method, double reltol, double abstol, const double* f0_modal, double t0,
                          double t1, gsg_ode** out);

/* ---- multi-GPU RK4 inside the library: peer-mapped state slabs, no NCCL on the data path ---------------------
 * One gsg_mg per rank (one process per GPU, or several ranks inside one process).  gsg_mg_create partitions the
 * plan (gsg_plan_set_partition) and allocates the rank's slab of 5 full-length vectors (u, v1..v4) plus a flag
 * area; the slabs are mapped into the partner ranks' address spaces -- between processes through CUDA IPC handles
 * (gsg_mg_ipc_handle / gsg_mg_connect_ipc; the host exchanges the 64-byte handles any way it likes), inside one
 * process directly (gsg_mg_connect_local).  Per right-hand side and partition dimension the sweeping rank pulls
 * the level-0 cells of the stage input out of its partner's slab over NVLink and the owner pull-adds the
 * contribution; ranks synchronise through counters in each other's flag area, and a whole RK4 step replays from
 * one CUDA graph.  u' = -sum_d a_d D_d u (BASELINE config 4; src/pdes.jl:179-180 operator). */
typedef struct gsg_mg gsg_mg;
#define GSG_MG_HANDLE_BYTES 64
int gsg_mg_create(gsg_plan* plan, int rank, int nranks, gsg_mg** out);
int gsg_mg_destroy(gsg_mg* mg);
int gsg_mg_ipc_handle(gsg_mg* mg, void* handle64);
/* handles: nranks consecutive 64-byte handles in rank order */
int gsg_mg_connect_ipc(gsg_mg* mg, const void* handles);
int gsg_mg_connect_local(gsg_mg* const* all, int nranks);
/* COLLECTIVE (every rank, after a host-level barrier): load this rank's owned blocks from a full reference-layout
 * host vector / write them back into one (entries of other ranks' blocks are left untouched) */
int gsg_mg_set_state(gsg_mg* mg, const double* u_host);
int gsg_mg_get_state(gsg_mg* mg, double* u_host);
/* fraction of the state this rank owns; bytes it pulls over NVLink per right-hand side */
int gsg_mg_owned_fraction(gsg_mg* mg, double* frac_out, int64_t* exchange_bytes_per_rhs_out);
/* COLLECTIVE: nsteps classical RK4 steps (Taylor form), asynchronous on the plan's stream */
int gsg_mg_rk4_advect(gsg_mg* mg, const double* a, double dt, int64_t nsteps);
/* one host thread drives all ranks of the partition (phases enqueued in lockstep) */
int gsg_mg_rk4_advect_all(gsg_mg* const* all, int nranks, const double* a, double dt, int64_t nsteps);
/* wait for this rank's streams; reports a flag-wait timeout (20 s) as GSG_ERR_CUDA */
int gsg_mg_sync(gsg_mg* mg);

/* w = u + cw*k ; acc = (first ? u : acc) + ca*k   on `len` entries (any sub-range) */
int gsg_rk_stage_dev(gsg_plan* plan, int64_t len, const double* u, const double* k, double* acc,
                     double* w, double cw, double ca, int first);
/* u = acc + ca*k */
int gsg_rk_final_dev(gsg_plan* plan, int64_t len, double* u, const double* k, const double* acc,
                     double ca);
/* Time the dominant (streaming TMA) sweep kernel with CUDA events on the stream it is launched
 * on (on = 1: every launch; on > 1: only the first `on` launches after this call -- timing all
 * ~24 launches of every step costs ~8 % of the step); read returns the number of timed launches,
 * their summed duration and the DOFs they processed (bench.py's roofline figure). */
int gsg_profile_enable(gsg_plan* plan, int on);
int gsg_profile_read(gsg_plan* plan, int64_t* launches_out, double* total_ms_out, double* dofs_out);

/* development aids (tools/stamps.py, tools/placement.py): enable (first call) / read back per-phase clock
 * stamps of the sweep kernels; launch a spinner with a chosen resource footprint */
int gsg_debug_stamps(gsg_plan* plan, long long* out, int n);
int gsg_debug_spin(gsg_plan* plan, int which, int grid, int threads, int smem, int ns, int slot);

/* ---- batched reconstruct_DG -------------------------------------------------------------------- */
/* out[i] = reconstruct_DG(V2D(vcoeffs), points[:, i])   src/dg_methods.jl:150-165;
 * points is a column-major D x npts matrix (Julia Matrix{Float64}).  Called once per point
 * by mcerr's loop in the reference (src/error_measure.jl:12-19,39-41). */
int gsg_reconstruct(gsg_plan* plan, const double* vcoeffs, const double* points, int64_t npts,
                    double* out);
/* device variant: points are not validated on the host; a point outside [0, 1]^D (a BoundsError in the
 * reference) yields NaN in out_dev instead of an out-of-bounds read */
int gsg_reconstruct_dev(gsg_plan* plan, const double* vcoeffs_dev, const double* points_dev,
                        int64_t npts, double* out_dev);

/* ---- generic SpMV on the reference's assembled matrix (correctness / perf cross-check) ------ */
/* y = A * x with A a Julia SparseMatrixCSC (1-based Int64)   `*(RHS, x)` src/pdes.jl:63 */
int gsg_spmv_csc(int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                 const double* nzval, const double* x, double* y);
/* resident variant: upload once (converted to CSR int32 on the way), apply many times */
int gsg_csr_create(int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                   const double* nzval, int device, gsg_csr** out);
int gsg_csr_destroy(gsg_csr* A);
int gsg_csr_apply(gsg_csr* A, const double* x, double* y);
int gsg_csr_apply_dev(gsg_csr* A, const double* x_dev, double* y_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSG_B200_H */
