// vlasov.inl -- the Vlasov right-hand side `steprule` of vlasov_evolve on the device (included by gsg_b200.cu).
//
// src/pdes.jl:165-192.  Phase space has 2D dimensions (D positions, D momenta); the plan is the (2D, k, n) operator.
//   dfdxs_modal = [Ds[d] * f for d in 1:D]            -> matrix-free sweeps (this library's operator apply)
//   dfdps_modal = [Ds[d] * f for d in D+1:2D]
//   *_point     = [n2p * (m2n * x) ...]                -> resident CSR SpMVs on the host-built transform matrices
//   contrib1 = sum(v_point[d] .* dfdxs_point[d]); contrib2 = sum(F_point[d] .* dfdps_point[d])   -> fused pointwise kernels
//   return n2m * (p2n * (-contrib1 + contrib2))
// The four transform matrices are arguments of vlasov_evolve in the reference (built once by the host with
// make_modal2point_matrices / make_point2modal_matrices, src/multidim_nodal_basis.jl:127-141) and cross the C ABI as
// resident gsg_csr handles.  v_point is computed here as the reference does (src/pdes.jl:165-172):
// v_modal[i] = tensor_construct(2D, k, n, [j - D == i ? x_1D : one_1D]), v_point = n2p * (m2n * v_modal).

namespace {

// acc = (first ? 0 : acc) + a .* b     (multiply, then add: the reference's `sum(v .* p for d)`)
__global__ void vl_mul_acc_kernel(long long N, double* __restrict__ acc, const double* __restrict__ a,
                                  const double* __restrict__ b, int first) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const double p = __dmul_rn(a[i], b[i]);
        acc[i] = first ? p : __dadd_rn(acc[i], p);
    }
}

// out = -c1 + c2
__global__ void vl_combine_kernel(long long N, const double* __restrict__ c1, const double* __restrict__ c2, double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) out[i] = __dadd_rn(-c1[i], c2[i]);
}

}  // namespace

struct gsg_vlasov {
    gsg_plan* plan = nullptr;
    gsg_csr *m2n = nullptr, *n2p = nullptr, *p2n = nullptr, *n2m = nullptr;
    int Dspace = 0;
    int64_t N = 0;
    std::vector<DevBuf<double>> v_point, F_point;      // Dspace vectors of length N each (reference layout)
    DevBuf<double> tmp_pad, r1, r2, c1, c2, fin, fout;
};

namespace {

// k_pad (device layout) = steprule(f_pad)
int vlasov_rhs_dev(gsg_vlasov& V, const double* f_pad, double* k_pad) {
    gsg_plan& pl = *V.plan;
    nvtx_range r("vlasov steprule");
    const int D = V.Dspace;
    const int grid = elementwise_grid(pl, V.N);
    for (int d = 0; d < 2 * D; ++d) {
        GSG_TRY(sweep(pl, d, 1.0, f_pad, 0.0, V.tmp_pad.p));                                  // Ds[d] * f_modal
        GSG_TRY(copy_out(pl, V.r1.p, V.tmp_pad.p, cudaMemcpyDeviceToDevice));                 // device -> reference layout
        GSG_TRY(gsg_csr_apply_dev(V.m2n, V.r1.p, V.r2.p, pl.stream));
        GSG_TRY(gsg_csr_apply_dev(V.n2p, V.r2.p, V.r1.p, pl.stream));
        const bool xpart = d < D;
        const double* w = xpart ? V.v_point[d].p : V.F_point[d - D].p;
        vl_mul_acc_kernel<<<grid, 256, 0, pl.stream>>>(V.N, xpart ? V.c1.p : V.c2.p, w, V.r1.p, (d == 0 || d == D) ? 1 : 0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    vl_combine_kernel<<<grid, 256, 0, pl.stream>>>(V.N, V.c1.p, V.c2.p, V.r1.p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    GSG_TRY(gsg_csr_apply_dev(V.p2n, V.r1.p, V.r2.p, pl.stream));
    GSG_TRY(gsg_csr_apply_dev(V.n2m, V.r2.p, V.r1.p, pl.stream));
    return copy_in(pl, k_pad, V.r1.p, cudaMemcpyDeviceToDevice);                              // reference -> device layout
}

}  // namespace

extern "C" {

int gsg_vlasov_create(gsg_plan* plan, gsg_csr* m2n, gsg_csr* n2p, gsg_csr* p2n, gsg_csr* n2m,
                      const double* const* F_point, gsg_vlasov** out) {
    GSG_TRY(check_plan(plan));
    if (!m2n || !n2p || !p2n || !n2m || !F_point || !out) return fail(GSG_ERR_ARG, "null pointer");
    const gsg::IndexSet& S = plan->S;
    if (S.D % 2 != 0) return fail(GSG_ERR_ARG, "vlasov: the plan must have an even number of dimensions (D positions + D momenta)");
    if (S.k < 2) return fail(GSG_ERR_ARG, "vlasov: k >= 2 required (the coefficients of v need the linear mode)");
    for (gsg_csr* A : {m2n, n2p, p2n, n2m})
        if (A->m != S.N || A->n != S.N || A->device != plan->device) return fail(GSG_ERR_ARG, "DimensionMismatch: transform matrices must be N x N on the plan's device");
    std::unique_ptr<gsg_vlasov> V(new gsg_vlasov());
    V->plan = plan;
    V->m2n = m2n; V->n2p = n2p; V->p2n = p2n; V->n2m = n2m;
    V->Dspace = S.D / 2;
    V->N = S.N;
    const size_t N = (size_t)S.N;
    GSG_TRY(V->tmp_pad.resize((size_t)S.Npad));
    GSG_TRY(V->r1.resize(N)); GSG_TRY(V->r2.resize(N)); GSG_TRY(V->c1.resize(N)); GSG_TRY(V->c2.resize(N));
    GSG_TRY(V->fin.resize((size_t)S.Npad)); GSG_TRY(V->fout.resize((size_t)S.Npad));
    V->v_point.resize(V->Dspace);
    V->F_point.resize(V->Dspace);
    // v_modal[i] = tensor_construct(2D, k, n, [j - D == i ? x_1D : one_1D])   (src/pdes.jl:165-171)
    const int n1d = S.k << S.n;
    std::vector<double> one1(n1d, 0.0), x1(n1d, 0.0), vm(N);
    one1[0] = 1.0;                      // get_one_modal(1, k, n)
    x1[1] = 1.0 / std::sqrt(3.0);       // get_xi_modal(1, 1, k, n)   (src/basic_function_exact_coeffs.jl:12-19)
    for (int i = 0; i < V->Dspace; ++i) {
        std::vector<const double*> arr(S.D);
        for (int j = 0; j < S.D; ++j) arr[j] = (j - V->Dspace == i) ? x1.data() : one1.data();
        gsg::tensor_construct(S, arr.data(), vm.data());
        GSG_TRY(V->v_point[i].resize(N));
        GSG_TRY(V->F_point[i].resize(N));
        // everything on the plan's (non-blocking) stream, and waited for before vm / r1 are reused by the next dimension
        GSG_CUDA(cudaMemcpyAsync(V->r1.p, vm.data(), N * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
        GSG_TRY(gsg_csr_apply_dev(m2n, V->r1.p, V->r2.p, plan->stream));
        GSG_TRY(gsg_csr_apply_dev(n2p, V->r2.p, V->v_point[i].p, plan->stream));               // v_point = n2p * (m2n * v)
        GSG_CUDA(cudaStreamSynchronize(plan->stream));
        if (!F_point[i]) return fail(GSG_ERR_ARG, "null F_point vector");
        GSG_CUDA(cudaMemcpyAsync(V->F_point[i].p, F_point[i], N * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    }
    GSG_CUDA(cudaStreamSynchronize(plan->stream));
    *out = V.release();
    return 0;
}

int gsg_vlasov_destroy(gsg_vlasov* v) {
    if (!v) return 0;
    if (v->plan) { cudaSetDevice(v->plan->device); cudaStreamSynchronize(v->plan->stream); }
    delete v;
    return 0;
}

// out = steprule(t, f_modal) on host vectors (reference layout)          src/pdes.jl:174-192
int gsg_vlasov_rhs(gsg_vlasov* v, const double* f_modal, double* out) {
    if (!v || !f_modal || !out) return fail(GSG_ERR_ARG, "null pointer");
    GSG_TRY(check_plan(v->plan));
    gsg_plan& pl = *v->plan;
    GSG_TRY(copy_in(pl, v->fin.p, f_modal, cudaMemcpyHostToDevice));
    GSG_TRY(vlasov_rhs_dev(*v, v->fin.p, v->fout.p));
    GSG_TRY(copy_out(pl, out, v->fout.p, cudaMemcpyDeviceToHost));
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    return 0;
}

// v_point[i] (i 0-based) -> host, for hosts that want to inspect it
int gsg_vlasov_v_point(gsg_vlasov* v, int i, double* out) {
    if (!v || !out || i < 0 || i >= v->Dspace) return fail(GSG_ERR_ARG, "bad argument");
    GSG_TRY(check_plan(v->plan));
    GSG_CUDA(cudaMemcpy(out, v->v_point[i].p, sizeof(double) * (size_t)v->N, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
