// ode.inl -- adaptive explicit Runge-Kutta time integration resident on the device (included by gsg_b200.cu).
//
// The reference hands its right-hand-side closures to ODE.jl's `ode45` / `ode78` (src/pdes.jl:62-68, 113-119,
// 206-213; ODE.jl 2.4.0 is a third-party dependency, Manifest.toml:184-188, not under /root/reference).  This is
// the drop-in for that call: the same algorithm (ODE.jl src/runge_kutta.jl `oderk_adapt` with `bt_dopri5` /
// `bt_feh78`, `hinit`, `stepsize_hw92!`, `hermite_interp!`; restated for the tests in oracle/ode_oracle.py) with the
// state, the stages and every vector operation on the GPU.  Only the scalar step-size control runs on the host: one
// 16-byte read-back per attempted step.  Arithmetic order follows ODE.jl's loops (no FMA contraction) so that the
// oracle and the device take the same accept / reject decisions.
//
//   calc_next_k!      ytmp = y; for ss < s: ytmp += dt * ks[ss] * a[s, ss]
//   rk_embedded_step! ytrial = sum_s b[1, s] ks[s]; yerr = sum_s b[2, s] ks[s]; yerr = dt (ytrial - yerr); ytrial = y + dt ytrial
//   stepsize_hw92!    err = || yerr ./ (abstol + max(|y|, |ytrial|) reltol) ||_2 ; newdt = dt max(1/5, 0.8 err^(-1/(order+1)))

namespace {

constexpr int ODE_MAXS = 13;

struct OdeTableau {
    int S = 0, order = 0;
    bool fsal = false;
    double a[ODE_MAXS][ODE_MAXS];
    double b[2][ODE_MAXS];
    double c[ODE_MAXS];
};

struct OdeStageArgs {             // ytmp = y + sum_{ss < n} (dt * k[ss]) * coef[ss]
    const double* k[ODE_MAXS];
    double coef[ODE_MAXS];
    int n;
};

__global__ void ode_stage_kernel(long long N, const double* __restrict__ y, double* __restrict__ ytmp, double dt,
                                 const __grid_constant__ OdeStageArgs A) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        double v = y[i];
        for (int s = 0; s < A.n; ++s) v = __dadd_rn(v, __dmul_rn(__dmul_rn(dt, A.k[s][i]), A.coef[s]));
        ytmp[i] = v;
    }
}

struct OdeFinalArgs {
    const double* k[ODE_MAXS];
    double b1[ODE_MAXS], b2[ODE_MAXS];
    int S;
};

// ytrial and the scaled error; per-block partial sums of squares (deterministic two-level reduction) + NaN flag
__global__ void __launch_bounds__(256)
ode_final_kernel(long long N, const double* __restrict__ y, double* __restrict__ ytrial, double dt, double abstol,
                 double reltol, const __grid_constant__ OdeFinalArgs A, double* __restrict__ partial,
                 int* __restrict__ nanflag) {
    double sq = 0.0;
    bool bad = false;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        double yt = __dmul_rn(A.b1[0], A.k[0][i]);
        double ye = __dmul_rn(A.b2[0], A.k[0][i]);
        for (int s = 1; s < A.S; ++s) {
            const double ks = A.k[s][i];
            yt = __dadd_rn(yt, __dmul_rn(A.b1[s], ks));
            ye = __dadd_rn(ye, __dmul_rn(A.b2[s], ks));
        }
        ye = __dmul_rn(dt, __dsub_rn(yt, ye));
        const double y0 = y[i];
        yt = __dadd_rn(y0, __dmul_rn(dt, yt));
        ytrial[i] = yt;
        bad = bad || isnan(yt);
        const double sc = __dadd_rn(abstol, __dmul_rn(fmax(fabs(y0), fabs(yt)), reltol));
        const double e = ye / sc;
        sq = __dadd_rn(sq, __dmul_rn(e, e));
    }
    __shared__ double ws[256];
    ws[threadIdx.x] = sq;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) ws[threadIdx.x] += ws[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = ws[0];
    if (bad) atomicExch(nanflag, 1);
}

// out[0] = sum of partial[0..n) (fixed order), out[1] = nan flag (as a double); resets the flag
__global__ void ode_reduce_kernel(const double* __restrict__ partial, int n, int* __restrict__ nanflag, double* __restrict__ out) {
    __shared__ double ws[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    ws[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) ws[threadIdx.x] += ws[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = ws[0];
        out[1] = (double)*nanflag;
        *nanflag = 0;
    }
}

// partial[b] = max_i |x[i] - (z ? z[i] : 0)| over the block's share
__global__ void __launch_bounds__(256)
ode_maxabs_kernel(long long N, const double* __restrict__ x, const double* __restrict__ z, double* __restrict__ partial) {
    double m = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const double v = fabs(z ? __dsub_rn(x[i], z[i]) : x[i]);
        m = (v > m || isnan(v)) ? v : m;
    }
    __shared__ double ws[256];
    ws[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double v = ws[threadIdx.x + o];
            if (v > ws[threadIdx.x] || isnan(v)) ws[threadIdx.x] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = ws[0];
}

__global__ void ode_axpy_kernel(long long N, const double* __restrict__ x, double c, const double* __restrict__ f, double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
        out[i] = __dadd_rn(x[i], __dmul_rn(c, f[i]));
}

// hermite_interp!: y = (1-th) y0 + th y1 + th (th-1) ((1-2th)(y1-y0) + (th-1) dt f0 + th dt f1)
__global__ void ode_hermite_kernel(long long N, double theta, double dt, const double* __restrict__ y0, const double* __restrict__ y1,
                                   const double* __restrict__ f0, const double* __restrict__ f1, double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const double a = __dadd_rn(__dmul_rn(1.0 - theta, y0[i]), __dmul_rn(theta, y1[i]));
        const double inner = __dadd_rn(__dadd_rn(__dmul_rn(1.0 - 2.0 * theta, __dsub_rn(y1[i], y0[i])),
                                                 __dmul_rn(__dmul_rn(theta - 1.0, dt), f0[i])),
                                       __dmul_rn(__dmul_rn(theta, dt), f1[i]));
        out[i] = __dadd_rn(a, __dmul_rn(theta * (theta - 1.0), inner));
    }
}

void ode_fill_tableau(OdeTableau& T, int method) {
    std::memset(&T, 0, sizeof(T));
    auto R = [](double num, double den) { return num / den; };
    if (method == 45) {                 // ODE.jl bt_dopri5, order (5, 4)
        T.S = 7; T.order = 4; T.fsal = true;
        const double c[7] = {0, R(1, 5), R(3, 10), R(4, 5), R(8, 9), 1, 1};
        std::memcpy(T.c, c, sizeof(c));
        T.a[1][0] = R(1, 5);
        T.a[2][0] = R(3, 40); T.a[2][1] = R(9, 40);
        T.a[3][0] = R(44, 45); T.a[3][1] = R(-56, 15); T.a[3][2] = R(32, 9);
        T.a[4][0] = R(19372, 6561); T.a[4][1] = R(-25360, 2187); T.a[4][2] = R(64448, 6561); T.a[4][3] = R(-212, 729);
        T.a[5][0] = R(9017, 3168); T.a[5][1] = R(-355, 33); T.a[5][2] = R(46732, 5247); T.a[5][3] = R(49, 176); T.a[5][4] = R(-5103, 18656);
        T.a[6][0] = R(35, 384); T.a[6][2] = R(500, 1113); T.a[6][3] = R(125, 192); T.a[6][4] = R(-2187, 6784); T.a[6][5] = R(11, 84);
        const double b1[7] = {R(35, 384), 0, R(500, 1113), R(125, 192), R(-2187, 6784), R(11, 84), 0};
        const double b2[7] = {R(5179, 57600), 0, R(7571, 16695), R(393, 640), R(-92097, 339200), R(187, 2100), R(1, 40)};
        std::memcpy(T.b[0], b1, sizeof(b1));
        std::memcpy(T.b[1], b2, sizeof(b2));
    } else {                            // ODE.jl bt_feh78, order (7, 8)
        T.S = 13; T.order = 7; T.fsal = false;
        const double c[13] = {0, R(2, 27), R(1, 9), R(1, 6), R(5, 12), R(1, 2), R(5, 6), R(1, 6), R(2, 3), R(1, 3), 1, 0, 1};
        std::memcpy(T.c, c, sizeof(c));
        T.a[1][0] = R(2, 27);
        T.a[2][0] = R(1, 36); T.a[2][1] = R(1, 12);
        T.a[3][0] = R(1, 24); T.a[3][2] = R(1, 8);
        T.a[4][0] = R(5, 12); T.a[4][2] = R(-25, 16); T.a[4][3] = R(25, 16);
        T.a[5][0] = R(1, 20); T.a[5][3] = R(1, 4); T.a[5][4] = R(1, 5);
        T.a[6][0] = R(-25, 108); T.a[6][3] = R(125, 108); T.a[6][4] = R(-65, 27); T.a[6][5] = R(125, 54);
        T.a[7][0] = R(31, 300); T.a[7][4] = R(61, 225); T.a[7][5] = R(-2, 9); T.a[7][6] = R(13, 900);
        T.a[8][0] = 2; T.a[8][3] = R(-53, 6); T.a[8][4] = R(704, 45); T.a[8][5] = R(-107, 9); T.a[8][6] = R(67, 90); T.a[8][7] = 3;
        T.a[9][0] = R(-91, 108); T.a[9][3] = R(23, 108); T.a[9][4] = R(-976, 135); T.a[9][5] = R(311, 54); T.a[9][6] = R(-19, 60);
        T.a[9][7] = R(17, 6); T.a[9][8] = R(-1, 12);
        T.a[10][0] = R(2383, 4100); T.a[10][3] = R(-341, 164); T.a[10][4] = R(4496, 1025); T.a[10][5] = R(-301, 82);
        T.a[10][6] = R(2133, 4100); T.a[10][7] = R(45, 82); T.a[10][8] = R(45, 164); T.a[10][9] = R(18, 41);
        T.a[11][0] = R(3, 205); T.a[11][5] = R(-6, 41); T.a[11][6] = R(-3, 205); T.a[11][7] = R(-3, 41); T.a[11][8] = R(3, 41); T.a[11][9] = R(6, 41);
        T.a[12][0] = R(-1777, 4100); T.a[12][3] = R(-341, 164); T.a[12][4] = R(4496, 1025); T.a[12][5] = R(-289, 82);
        T.a[12][6] = R(2193, 4100); T.a[12][7] = R(51, 82); T.a[12][8] = R(33, 164); T.a[12][9] = R(12, 41); T.a[12][11] = 1;
        const double b1[13] = {R(41, 840), 0, 0, 0, 0, R(34, 105), R(9, 35), R(9, 35), R(9, 280), R(9, 280), R(41, 840), 0, 0};
        const double b2[13] = {0, 0, 0, 0, 0, R(34, 105), R(9, 35), R(9, 35), R(9, 280), R(9, 280), 0, R(41, 840), R(41, 840)};
        std::memcpy(T.b[0], b1, sizeof(b1));
        std::memcpy(T.b[1], b2, sizeof(b2));
    }
}

}  // namespace

struct gsg_ode {
    gsg_plan* plan = nullptr;
    gsg_csr* A = nullptr;
    gsg_vlasov* V = nullptr;
    int kind = 0;
    std::vector<double> a;                  // advection coefficients
    OdeTableau T;
    int64_t len = 0, len_ref = 0;           // device length (padded) / host length
    int nvec_state = 1;                     // 1, or 2 for the wave system [u; v]
    double reltol = 1e-5, abstol = 1e-8, maxstep = 0, minstep = 0;
    double t = 0, tend = 0, dt = 0, tdir = 1, last_dt = 0;
    bool laststep = false, done = false, failed = false;
    int timeout = 0;
    int64_t nacc = 0, nrej = 0, nrhs = 0;
    std::vector<DevBuf<double>> ks;         // S + 1 stage vectors (one spare for the non-FSAL f1)
    std::vector<double*> kp;                // current binding of ks[0..S-1], kp[S] = spare
    DevBuf<double> y, ytrial, ytmp, partial, red;
    DevBuf<int> nanflag;
    double* yp = nullptr;                   // state at time t
    double* ytp = nullptr;                  // trial state (after an accepted step: the state at t, yp holds the step's start)
    double* f0_last = nullptr;              // f at the start / end of the last accepted step (for hermite_interp)
    double* f1_last = nullptr;
    int grid = 1;
};

namespace {

int ode_rhs(gsg_ode& O, const double* w, double* k) {
    gsg_plan& pl = *O.plan;
    ++O.nrhs;
    switch (O.kind) {
        case GSG_RHS_ADVECT: return advect_rhs(pl, O.a.data(), w, k);
        case GSG_RHS_WAVE: {
            const int64_t Np = pl.S.Npad;
            GSG_CUDA(cudaMemcpyAsync(k, w + Np, Np * sizeof(double), cudaMemcpyDeviceToDevice, pl.stream));
            return laplacian(pl, w, k + Np, pl.wtmp.p);
        }
        case GSG_RHS_CSR: return gsg_csr_apply_dev(O.A, w, k, pl.stream);
        case GSG_RHS_VLASOV: return vlasov_rhs_dev(*O.V, w, k);
    }
    return fail(GSG_ERR_ARG, "bad right-hand-side kind");
}

int ode_maxabs(gsg_ode& O, const double* x, const double* z, double* out) {
    gsg_plan& pl = *O.plan;
    ode_maxabs_kernel<<<O.grid, 256, 0, pl.stream>>>(O.len, x, z, O.partial.p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    std::vector<double> h(O.grid);
    GSG_CUDA(cudaMemcpyAsync(h.data(), O.partial.p, sizeof(double) * O.grid, cudaMemcpyDeviceToHost, pl.stream));
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    double m = 0.0;
    for (double v : h) m = (v > m || std::isnan(v)) ? v : m;
    *out = m;
    return 0;
}

int ode_copy_in(gsg_ode& O, double* dev, const double* host) {
    gsg_plan& pl = *O.plan;
    if (O.kind == GSG_RHS_CSR) {
        GSG_CUDA(cudaMemcpyAsync(dev, host, sizeof(double) * O.len, cudaMemcpyHostToDevice, pl.stream));
        return 0;
    }
    for (int v = 0; v < O.nvec_state; ++v)
        GSG_TRY(copy_in(pl, dev + (size_t)v * pl.S.Npad, host + (size_t)v * pl.S.N, cudaMemcpyHostToDevice));
    return 0;
}

int ode_copy_out(gsg_ode& O, double* host, const double* dev) {
    gsg_plan& pl = *O.plan;
    if (O.kind == GSG_RHS_CSR) {
        GSG_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * O.len, cudaMemcpyDeviceToHost, pl.stream));
    } else {
        for (int v = 0; v < O.nvec_state; ++v)
            GSG_TRY(copy_out(pl, host + (size_t)v * pl.S.N, dev + (size_t)v * pl.S.Npad, cudaMemcpyDeviceToHost));
    }
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    return 0;
}

// ODE.jl hinit: first step size, direction, ks[0] = F(t0, y0)
int ode_hinit(gsg_ode& O) {
    gsg_plan& pl = *O.plan;
    const double t0 = O.t, tend = O.tend;
    O.tdir = tend > t0 ? 1.0 : -1.0;
    double n0 = 0, n1 = 0, n2 = 0;
    GSG_TRY(ode_maxabs(O, O.yp, nullptr, &n0));
    const double tau = std::max(O.reltol * n0, O.abstol);
    const double d0 = n0 / tau;
    GSG_TRY(ode_rhs(O, O.yp, O.kp[0]));
    GSG_TRY(ode_maxabs(O, O.kp[0], nullptr, &n1));
    const double d1 = n1 / tau;
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    ode_axpy_kernel<<<O.grid, 256, 0, pl.stream>>>(O.len, O.yp, O.tdir * h0, O.kp[0], O.ytmp.p);      // Euler step
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_TRY(ode_rhs(O, O.ytmp.p, O.kp[1]));
    GSG_TRY(ode_maxabs(O, O.kp[1], O.kp[0], &n2));
    const double d2 = n2 / (tau * h0);
    double h1;
    if (std::max(d1, d2) <= 1e-15) h1 = std::max(1e-6, 1e-3 * h0);
    else h1 = std::pow(10.0, -(2.0 + std::log10(std::max(d1, d2))) / (O.T.order + 1.0));
    O.dt = O.tdir * std::min(std::min(100 * h0, h1), O.tdir * (tend - t0));
    return 0;
}

// one ATTEMPTED step: stages, trial state, error norm -> (err, nan)
int ode_attempt(gsg_ode& O, double* err_out, bool* nan_out) {
    gsg_plan& pl = *O.plan;
    const OdeTableau& T = O.T;
    for (int s = 1; s < T.S; ++s) {
        OdeStageArgs A;
        A.n = s;
        for (int ss = 0; ss < s; ++ss) { A.k[ss] = O.kp[ss]; A.coef[ss] = T.a[s][ss]; }
        ode_stage_kernel<<<O.grid, 256, 0, pl.stream>>>(O.len, O.yp, O.ytmp.p, O.dt, A);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GSG_TRY(ode_rhs(O, O.ytmp.p, O.kp[s]));
    }
    OdeFinalArgs F;
    F.S = T.S;
    for (int s = 0; s < T.S; ++s) { F.k[s] = O.kp[s]; F.b1[s] = T.b[0][s]; F.b2[s] = T.b[1][s]; }
    ode_final_kernel<<<O.grid, 256, 0, pl.stream>>>(O.len, O.yp, O.ytp, O.dt, O.abstol, O.reltol, F, O.partial.p, O.nanflag.p);
    ode_reduce_kernel<<<1, 256, 0, pl.stream>>>(O.partial.p, O.grid, O.nanflag.p, O.red.p);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    double h[2];
    GSG_CUDA(cudaMemcpyAsync(h, O.red.p, sizeof(h), cudaMemcpyDeviceToHost, pl.stream));
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    *err_out = std::sqrt(h[0]);
    *nan_out = h[1] != 0.0 || std::isnan(h[0]);
    return 0;
}

}  // namespace

extern "C" {

static int ode_create_impl(gsg_plan* plan, int rhs_kind, const double* a, gsg_csr* A, gsg_vlasov* V, int method, double reltol,
                           double abstol, const double* y0_host, double t0, double t1, gsg_ode** out) {
    GSG_TRY(check_plan(plan));
    if (!out || !y0_host) return fail(GSG_ERR_ARG, "null pointer");
    if (method != 45 && method != 78) return fail(GSG_ERR_ARG, "ArgumentError(:order): method must be 45 or 78");
    if (!(t1 != t0)) return fail(GSG_ERR_ARG, "Zero time span");
    if (rhs_kind == GSG_RHS_ADVECT && !a) return fail(GSG_ERR_ARG, "advection coefficients required");
    if (rhs_kind == GSG_RHS_CSR && (!A || A->m != A->n)) return fail(GSG_ERR_ARG, "square resident matrix required");
    if (rhs_kind == GSG_RHS_VLASOV && !V) return fail(GSG_ERR_ARG, "vlasov handle required");
    if (rhs_kind != GSG_RHS_ADVECT && rhs_kind != GSG_RHS_WAVE && rhs_kind != GSG_RHS_CSR && rhs_kind != GSG_RHS_VLASOV)
        return fail(GSG_ERR_ARG, "bad right-hand-side kind");
    std::unique_ptr<gsg_ode> O(new gsg_ode());
    O->plan = plan;
    O->A = A;
    O->V = V;
    O->kind = rhs_kind;
    if (a) O->a.assign(a, a + plan->S.D);
    ode_fill_tableau(O->T, method);
    if (rhs_kind == GSG_RHS_CSR) {
        if (A->device != plan->device) return fail(GSG_ERR_ARG, "matrix and plan live on different devices");
        O->len = O->len_ref = A->m;
    } else {
        O->nvec_state = rhs_kind == GSG_RHS_WAVE ? 2 : 1;
        O->len = plan->S.Npad * O->nvec_state;
        O->len_ref = plan->S.N * O->nvec_state;
        if (rhs_kind == GSG_RHS_WAVE) GSG_TRY(plan->wtmp.resize((size_t)plan->S.Npad));
    }
    O->reltol = reltol > 0 ? reltol : 1.0e-5;           // ODE.jl defaults
    O->abstol = abstol > 0 ? abstol : 1.0e-8;
    O->maxstep = std::fabs(t1 - t0) / 2.5;
    O->minstep = std::fabs(t1 - t0) / 1e18;
    O->t = t0;
    O->tend = t1;
    const int S = O->T.S;
    O->ks.resize(S + 1);
    O->kp.resize(S + 1);
    for (int s = 0; s <= S; ++s) { GSG_TRY(O->ks[s].resize((size_t)O->len)); O->kp[s] = O->ks[s].p; }
    GSG_TRY(O->y.resize((size_t)O->len));
    GSG_TRY(O->ytrial.resize((size_t)O->len));
    GSG_TRY(O->ytmp.resize((size_t)O->len));
    O->grid = elementwise_grid(*plan, O->len);
    GSG_TRY(O->partial.resize((size_t)O->grid));
    GSG_TRY(O->red.resize(2));
    GSG_TRY(O->nanflag.resize(1));
    O->yp = O->y.p;
    O->ytp = O->ytrial.p;
    GSG_TRY(ode_copy_in(*O, O->yp, y0_host));
    GSG_TRY(ode_hinit(*O));
    *out = O.release();
    return 0;
}

int gsg_ode_create(gsg_plan* plan, int rhs_kind, const double* a, gsg_csr* A, int method, double reltol, double abstol,
                   const double* y0_host, double t0, double t1, gsg_ode** out) {
    if (rhs_kind == GSG_RHS_VLASOV) return fail(GSG_ERR_ARG, "use gsg_ode_create_vlasov");
    return ode_create_impl(plan, rhs_kind, a, A, nullptr, method, reltol, abstol, y0_host, t0, t1, out);
}

// f' = steprule(t, f): the integrator call of vlasov_evolve (src/pdes.jl:206-213)
int gsg_ode_create_vlasov(gsg_vlasov* v, int method, double reltol, double abstol, const double* f0_modal, double t0,
                          double t1, gsg_ode** out) {
    if (!v) return fail(GSG_ERR_ARG, "null vlasov handle");
    return ode_create_impl(v->plan, GSG_RHS_VLASOV, nullptr, nullptr, v, method, reltol, abstol, f0_modal, t0, t1, out);
}

int gsg_ode_destroy(gsg_ode* ode) {
    if (!ode) return 0;
    if (ode->plan) { cudaSetDevice(ode->plan->device); cudaStreamSynchronize(ode->plan->stream); }
    delete ode;
    return 0;
}

// Advance to the next ACCEPTED step (rejected attempts are retried with the smaller step inside).  After the call
// the state is at *t_out; *done_out != 0 once tend has been reached (or the step size fell below minstep).
int gsg_ode_step(gsg_ode* ode, double* t_out, double* dt_out, int* done_out) {
    if (!ode) return fail(GSG_ERR_ARG, "null integrator");
    gsg_ode& O = *ode;
    GSG_TRY(check_plan(O.plan));
    if (O.done) {
        if (t_out) *t_out = O.t;
        if (dt_out) *dt_out = 0.0;
        if (done_out) *done_out = 1;
        return 0;
    }
    nvtx_range r("ode step");
    const int S = O.T.S;
    const double order = O.T.order;
    for (;;) {
        double err = 0;
        bool isnan_ = false;
        GSG_TRY(ode_attempt(O, &err, &isnan_));
        // stepsize_hw92!
        const double facmax = 5.0, facmin = 1.0 / facmax, fac = 0.8;
        double newdt;
        if (isnan_) {
            err = 10.0;
            newdt = O.dt * facmin;
            O.timeout = 5;
        } else {
            const double grow = err > 0 ? fac * std::pow(1.0 / err, 1.0 / (order + 1.0)) : INFINITY;
            newdt = std::min(O.maxstep, O.tdir * O.dt * std::max(facmin, grow));
            if (O.timeout > 0) {
                newdt = std::min(newdt, O.dt);
                --O.timeout;
            }
            newdt *= O.tdir;
        }
        if (err <= 1.0) {
            ++O.nacc;
            // f1 = F(t + dt, ytrial): the last stage for an FSAL tableau, else one more evaluation
            double* f0 = O.kp[0];
            double* f1;
            if (O.T.fsal) {
                f1 = O.kp[S - 1];
            } else {
                GSG_TRY(ode_rhs(O, O.ytp, O.kp[S]));
                f1 = O.kp[S];
            }
            // ks[1] = f1 for the next step; keep f0 / f1 of this step for hermite_interp (they stay untouched until
            // the next attempt overwrites stage vectors 1..S-1 -- f0 moves to the slot f1 came from)
            if (O.T.fsal) std::swap(O.kp[0], O.kp[S - 1]);
            else std::swap(O.kp[0], O.kp[S]);
            O.f0_last = f0;
            O.f1_last = f1;
            std::swap(O.yp, O.ytp);            // yp = new state, ytp = the step's start
            O.last_dt = O.dt;
            O.t += O.dt;
            if (O.laststep) {
                O.done = true;
            } else {
                O.dt = newdt;
                if (O.tdir * (O.t + O.dt * 1.01) >= O.tdir * O.tend) {       // hit the end point exactly
                    O.dt = O.tend - O.t;
                    O.laststep = true;
                }
            }
            break;
        } else if (std::fabs(newdt) < O.minstep) {
            O.done = O.failed = true;
            break;
        } else {
            O.laststep = false;
            ++O.nrej;
            O.dt = newdt;
            O.timeout = 5;
        }
    }
    if (t_out) *t_out = O.t;
    if (dt_out) *dt_out = O.last_dt;
    if (done_out) *done_out = O.done ? 1 : 0;
    if (O.failed) return fail(GSG_ERR_UNSUPPORTED, "Warning: dt < minstep.  Stopping.");
    return 0;
}

int gsg_ode_state(gsg_ode* ode, double* y_host) {
    if (!ode || !y_host) return fail(GSG_ERR_ARG, "null pointer");
    GSG_TRY(check_plan(ode->plan));
    return ode_copy_out(*ode, y_host, ode->yp);
}

// Hermite interpolation (3rd order, ODE.jl hermite_interp!) at a time inside the LAST accepted step
int gsg_ode_interp(gsg_ode* ode, double tquery, double* y_host) {
    if (!ode || !y_host) return fail(GSG_ERR_ARG, "null pointer");
    gsg_ode& O = *ode;
    GSG_TRY(check_plan(O.plan));
    if (O.nacc == 0 || !O.f0_last) return fail(GSG_ERR_ARG, "no accepted step yet");
    const double tstart = O.t - O.last_dt;
    if (O.tdir * (tquery - tstart) < 0 || O.tdir * (O.t - tquery) < 0) return fail(GSG_ERR_ARG, "query time outside the last accepted step");
    const double theta = (tquery - tstart) / O.last_dt;
    ode_hermite_kernel<<<O.grid, 256, 0, O.plan->stream>>>(O.len, theta, O.last_dt, O.ytp, O.yp, O.f0_last, O.f1_last, O.ytmp.p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return ode_copy_out(O, y_host, O.ytmp.p);
}

int gsg_ode_stats(gsg_ode* ode, int64_t* accepted, int64_t* rejected, int64_t* rhs_evals) {
    if (!ode) return fail(GSG_ERR_ARG, "null integrator");
    if (accepted) *accepted = ode->nacc;
    if (rejected) *rejected = ode->nrej;
    if (rhs_evals) *rhs_evals = ode->nrhs;
    return 0;
}

}  // extern "C"
