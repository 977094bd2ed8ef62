// multi_gpu.inl -- block-partitioned RK4 over N GPUs with peer-mapped state vectors (included by gsg_b200.cu).
//
// SURVEY.md 8(e) / DESIGN.md section 7.  Every rank holds one slab of NVEC full-length vectors (device layout)
// plus a small flag area; slabs have the same internal layout on every rank and are mapped into every other
// rank's address space (CUDA IPC between processes, plain pointers + cudaDeviceEnablePeerAccess inside one
// process).  Nothing travels through NCCL or the host: along a partition dimension e the sweeping rank (bit 0)
// PULLS the level_e == 0 cells of the stage input out of its partner's slab over NVLink, sweeps the straddling
// poles locally, and the owner PULL-ADDS the contribution to its level-0 cells out of the sweeper's slab.
// Ranks synchronise through 64-bit counters written into each other's flag area (st.release.sys / ld.acquire.sys
// in one-thread kernels), so a whole RK4 step is a fixed sequence of kernels on two streams and can be replayed
// from one CUDA graph; the counters are derived from a device-resident RHS counter, never from kernel arguments.
//
// Flag area of rank r (unsigned long long [nranks][MG_NFLAG]): row q is written by rank q only.
//   [q][0]      READY  : value v = "the input vector of rank q's RHS number v is final on q's owned cells"
//   [q][1 + j]  SWEPT_j: value v = "rank q has swept the straddling poles of partition bit j for RHS number v"

namespace {

constexpr int MG_NVEC = 5;              // u, v1 .. v4 (Taylor form of RK4)
constexpr int MG_NFLAG = 8;             // READY + up to 7 partition bits
constexpr unsigned long long MG_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

__global__ void mg_bump_kernel(unsigned long long* seq) { *seq += 1ull; }

__global__ void mg_signal_kernel(unsigned long long* target, const unsigned long long* seq, unsigned long long off) {
    const unsigned long long v = *seq + off;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(target), "l"(v) : "memory");
}

__global__ void mg_wait_kernel(const unsigned long long* flag, const unsigned long long* seq, unsigned long long off,
                               int* err) {
    const unsigned long long want = *seq + off;
    const long long t0 = gsgk::gtimer();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= want) break;
        if ((unsigned long long)(gsgk::gtimer() - t0) > MG_TIMEOUT_NS) {      // never hang the GPU: flag the error
            atomicExch(err, 1);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// w_local[cell] = w_peer[cell], k_local[cell] = 0 on the listed multi-cells (16-byte aligned, KDp even)
__global__ void __launch_bounds__(256)
mg_pull_copy_kernel(const int* __restrict__ cells, long long ncells, int KDp, double* __restrict__ w_local,
                    const double* __restrict__ w_peer, double* __restrict__ k_local) {
    const int n2 = KDp >> 1;
    for (long long ci = blockIdx.x; ci < ncells; ci += gridDim.x) {
        const long long base = (long long)cells[ci] * KDp;
        const double2* src = reinterpret_cast<const double2*>(w_peer + base);
        double2* dst = reinterpret_cast<double2*>(w_local + base);
        double2* kz = reinterpret_cast<double2*>(k_local + base);
        for (int e = threadIdx.x; e < n2; e += blockDim.x) {
            dst[e] = src[e];
            kz[e] = make_double2(0.0, 0.0);
        }
    }
}

// k_local[cell] += k_peer[cell]
__global__ void __launch_bounds__(256)
mg_pull_add_kernel(const int* __restrict__ cells, long long ncells, int KDp, double* __restrict__ k_local,
                   const double* __restrict__ k_peer) {
    const int n2 = KDp >> 1;
    for (long long ci = blockIdx.x; ci < ncells; ci += gridDim.x) {
        const long long base = (long long)cells[ci] * KDp;
        const double2* src = reinterpret_cast<const double2*>(k_peer + base);
        double2* dst = reinterpret_cast<double2*>(k_local + base);
        for (int e = threadIdx.x; e < n2; e += blockDim.x) {
            const double2 a = src[e];
            // reductions at the L2: the pull-adds of different partition dimensions overlap in cells and in time
            atomicAdd(reinterpret_cast<double*>(dst + e), a.x);
            atomicAdd(reinterpret_cast<double*>(dst + e) + 1, a.y);
        }
    }
}

// y[cell] = 0 on the listed multi-cells
__global__ void __launch_bounds__(256)
mg_zero_cells_kernel(const int* __restrict__ cells, long long ncells, int KDp, double* __restrict__ y) {
    const int n2 = KDp >> 1;
    for (long long ci = blockIdx.x; ci < ncells; ci += gridDim.x) {
        double2* dst = reinterpret_cast<double2*>(y + (long long)cells[ci] * KDp);
        for (int e = threadIdx.x; e < n2; e += blockDim.x) dst[e] = make_double2(0.0, 0.0);
    }
}

}  // namespace

struct gsg_mg {
    gsg_plan* plan = nullptr;
    int rank = 0, nranks = 1, bits = 0;
    int64_t Npad = 0;
    double* slab = nullptr;                       // MG_NVEC * Npad doubles, then the flag area
    size_t slab_bytes = 0;
    std::vector<double*> peer_slab;               // slab of every rank in THIS process's address space
    std::vector<char> opened;                     // peer_slab[q] came from cudaIpcOpenMemHandle
    bool connected = false;
    unsigned long long* dev_seq = nullptr;        // RHS counter (device resident: graph replays stay valid)
    int* dev_err = nullptr;
    DevBuf<unsigned long long> seqbuf;
    DevBuf<int> errbuf;
    DevBuf<int> owned_cells;
    int64_t n_owned = 0;
    struct Ex {                                   // one partition dimension
        int d = 0;                                // 0-based dimension
        int j = 0;                                // partition bit
        int partner = -1;
        int mybit = 0;
        DevBuf<int> cells;                        // level_d == 0 multi-cells of the straddling poles
        int64_t ncells = 0;
    };
    std::vector<Ex> ex;
    std::vector<std::pair<int64_t, int64_t>> owned_blocks;     // (reference offset, device offset) of owned blocks
    std::vector<int64_t> owned_block_cells;
    cudaStream_t comm = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> ev_x;
    cudaGraphExec_t step_exec = nullptr;
    int64_t graph_launches = 0;
    double graph_dt = 0.0;
    std::vector<double> graph_a;

    double* vec(int q, int i) const { return peer_slab[q] + (size_t)i * Npad; }
    unsigned long long* flags(int q) const { return reinterpret_cast<unsigned long long*>(peer_slab[q] + (size_t)MG_NVEC * Npad); }
};

namespace {

int mg_check(const gsg_mg* m) {
    if (!m || !m->plan) return fail(GSG_ERR_ARG, "null multi-GPU handle");
    GSG_CUDA(cudaSetDevice(m->plan->device));
    return 0;
}

int mg_signal(gsg_mg& M, cudaStream_t st, int to_rank, int slot, unsigned long long off) {
    unsigned long long* target = M.flags(to_rank) + (size_t)M.rank * MG_NFLAG + slot;
    mg_signal_kernel<<<1, 1, 0, st>>>(target, M.dev_seq, off);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

int mg_wait(gsg_mg& M, cudaStream_t st, int from_rank, int slot, unsigned long long off) {
    const unsigned long long* flag = M.flags(M.rank) + (size_t)from_rank * MG_NFLAG + slot;
    mg_wait_kernel<<<1, 1, 0, st>>>(flag, M.dev_seq, off, M.dev_err);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSG_CUDA(cudaGetLastError());
    return 0;
}

// ---- the three phases of one right-hand side k = sum_d c_d D_d w on the owned cells ---------------------
// phase A: start the RHS; on the comm stream pull the level-0 cells of the stage input from the partners
int mg_rhs_phase_a(gsg_mg& M, int wi, int ki, const double* c) {
    gsg_plan& pl = *M.plan;
    nvtx_range r("mg_rhs:pull_x");
    mg_bump_kernel<<<1, 1, 0, pl.stream>>>(M.dev_seq);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    bool any_pull = false;
    for (const gsg_mg::Ex& e : M.ex) any_pull = any_pull || (e.mybit == 0 && e.ncells > 0 && c[e.d] != 0.0);
    if (!any_pull) return 0;              // (a forked stream without work would stay unjoined in a graph capture)
    GSG_CUDA(cudaEventRecord(M.ev_begin, pl.stream));
    GSG_CUDA(cudaStreamWaitEvent(M.comm, M.ev_begin, 0));
    for (size_t i = 0; i < M.ex.size(); ++i) {
        gsg_mg::Ex& e = M.ex[i];
        if (e.mybit != 0 || e.ncells == 0 || c[e.d] == 0.0) continue;
        GSG_TRY(mg_wait(M, M.comm, e.partner, 0, 0));                    // partner's READY >= this RHS
        const int grid = (int)std::min<int64_t>(e.ncells, (int64_t)pl.sm_count * 8);
        mg_pull_copy_kernel<<<grid, 256, 0, M.comm>>>(e.cells.p, e.ncells, (int)pl.S.kDp, M.vec(M.rank, wi),
                                                       M.vec(e.partner, wi), M.vec(M.rank, ki));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GSG_CUDA(cudaGetLastError());
        GSG_CUDA(cudaEventRecord(M.ev_x[i], M.comm));
    }
    return 0;
}

// phase B: every sweep of the right-hand side through the concurrent scheduler -- the local directions start at
// once (the first initialises k), the partition dimensions as soon as their level-0 cells have been pulled
// (straddling poles on the bit-0 rank, p == 0 poles on the bit-1 rank); then the SWEPT signals
int mg_rhs_phase_b(gsg_mg& M, int wi, int ki, const double* c) {
    gsg_plan& pl = *M.plan;
    const int D = pl.S.D;
    const double* w = M.vec(M.rank, wi);
    double* k = M.vec(M.rank, ki);
    nvtx_range r("mg_rhs:sweeps");
    static const bool serial = getenv("GSG_RHS_SERIAL") != nullptr;
    if (!serial) {
        cudaEvent_t pre[16] = {nullptr};
        unsigned mask = 0;
        for (int d = 0; d < D; ++d)
            if (c[d] != 0.0) mask |= 1u << d;
        for (size_t i = 0; i < M.ex.size(); ++i) {
            gsg_mg::Ex& e = M.ex[i];
            if (c[e.d] != 0.0 && e.mybit == 0 && e.ncells > 0) pre[e.d] = M.ev_x[i];
        }
        // k = 0 on the owned cells (1/nranks of the state), then EVERY sweep accumulates: no piece has to finish
        // before the others may start (a single rank initialises k with its first sweep instead: no extra pass)
        static const bool zero_first = !getenv("GSG_MG_NO_ZERO");
        if (M.nranks > 1 && zero_first && M.n_owned > 0) {
            const int grid = (int)std::min<int64_t>(M.n_owned, (int64_t)pl.sm_count * 8);
            mg_zero_cells_kernel<<<grid, 256, 0, pl.stream>>>(M.owned_cells.p, M.n_owned, (int)pl.S.kDp, k);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            GSG_TRY(rhs_concurrent(pl, c, mask, w, k, 1.0, pre));
        } else {
            bool local_any = false;
            for (int d = 0; d < D - M.bits; ++d) local_any = local_any || c[d] != 0.0;
            if (!local_any) GSG_TRY(sweep(pl, 0, 0.0, w, 0.0, k));          // k starts from zero on the owned cells
            GSG_TRY(rhs_concurrent(pl, c, mask, w, k, local_any ? 0.0 : 1.0, pre));
        }
        for (gsg_mg::Ex& e : M.ex)
            if (c[e.d] != 0.0 && e.mybit == 0 && e.ncells > 0) GSG_TRY(mg_signal(M, pl.stream, e.partner, 1 + e.j, 0));
        return 0;
    }
    unsigned local_mask = 0;
    for (int d = 0; d < D - M.bits; ++d)
        if (c[d] != 0.0) local_mask |= 1u << d;
    if (local_mask == 0) {
        GSG_TRY(sweep(pl, 0, 0.0, w, 0.0, k));
    } else if (can_fuse(pl, c, local_mask)) {
        GSG_TRY(grad_fused(pl, c, w, k, local_mask, 0.0));
    } else {
        bool first = true;
        for (int d = 0; d < D; ++d) {
            if (!((local_mask >> d) & 1)) continue;
            GSG_TRY(sweep(pl, d, c[d], w, first ? 0.0 : 1.0, k));
            first = false;
        }
    }
    for (size_t i = 0; i < M.ex.size(); ++i) {
        gsg_mg::Ex& e = M.ex[i];
        if (c[e.d] == 0.0) continue;
        const bool pulls = e.mybit == 0 && e.ncells > 0;
        if (pulls) GSG_CUDA(cudaStreamWaitEvent(pl.stream, M.ev_x[i], 0));
        GSG_TRY(sweep(pl, e.d, c[e.d], w, 1.0, k));
        if (pulls) GSG_TRY(mg_signal(M, pl.stream, e.partner, 1 + e.j, 0));        // SWEPT_j = this RHS
    }
    return 0;
}

// phase C: pull-add the partners' contributions to my level-0 cells (each partition dimension on its own pool
// stream: wait for the partner's SWEPT, then reduce its scratch cells into mine); then k is final
int mg_rhs_phase_c(gsg_mg& M, int ki, const double* c, bool signal_ready) {
    gsg_plan& pl = *M.plan;
    nvtx_range r("mg_rhs:pull_add");
    GSG_CUDA(cudaEventRecord(pl.ev_p1, pl.stream));
    PoolCtx ctx(pl);
    for (size_t i = 0; i < M.ex.size(); ++i) {
        gsg_mg::Ex& e = M.ex[i];
        if (e.mybit != 1 || e.ncells == 0 || c[e.d] == 0.0) continue;
        cudaStream_t st;
        GSG_TRY(ctx.take(&st, pl.ev_p1));
        GSG_TRY(mg_wait(M, st, e.partner, 1 + e.j, 0));
        const int grid = (int)std::min<int64_t>(e.ncells, (int64_t)pl.sm_count * 8);
        mg_pull_add_kernel<<<grid, 256, 0, st>>>(e.cells.p, e.ncells, (int)pl.S.kDp, M.vec(M.rank, ki), M.vec(e.partner, ki));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GSG_CUDA(cudaGetLastError());
    }
    GSG_TRY(ctx.join());
    if (signal_ready) {
        for (gsg_mg::Ex& e : M.ex)
            if (e.mybit == 1 && e.ncells > 0) GSG_TRY(mg_signal(M, pl.stream, e.partner, 0, 1));    // READY = next RHS
    }
    return 0;
}

// u += dt v1 + dt^2/2 v2 + dt^3/6 v3 + dt^4/24 v4 on the owned cells; then u is the next step's input
int mg_combine(gsg_mg& M, double dt) {
    gsg_plan& pl = *M.plan;
    nvtx_range r("mg_combine");
    if (M.n_owned > 0) {
        const int grid = (int)std::min<int64_t>(M.n_owned, (int64_t)pl.sm_count * 16);
        rk4_taylor_cells_kernel<<<grid, 256, 0, pl.stream>>>(M.owned_cells.p, M.n_owned, (int)pl.S.kDp, M.vec(M.rank, 0),
                                                              M.vec(M.rank, 1), M.vec(M.rank, 2), M.vec(M.rank, 3),
                                                              M.vec(M.rank, 4), dt, dt * dt / 2.0, dt * dt * dt / 6.0,
                                                              dt * dt * dt * dt / 24.0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GSG_CUDA(cudaGetLastError());
    }
    for (gsg_mg::Ex& e : M.ex)
        if (e.mybit == 1 && e.ncells > 0) GSG_TRY(mg_signal(M, pl.stream, e.partner, 0, 1));
    return 0;
}

// one RK4 step of this rank, everything enqueued on the plan's stream / the comm stream
int mg_step(gsg_mg& M, const double* c, double dt) {
    for (int i = 0; i < 4; ++i) {
        GSG_TRY(mg_rhs_phase_a(M, i, i + 1, c));
        GSG_TRY(mg_rhs_phase_b(M, i, i + 1, c));
        GSG_TRY(mg_rhs_phase_c(M, i + 1, c, i < 3));
    }
    return mg_combine(M, dt);
}

int mg_check_err(gsg_mg& M) {
    int err = 0;
    GSG_CUDA(cudaMemcpy(&err, M.dev_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(GSG_ERR_CUDA, "multi-GPU: timed out waiting for a partner rank's flag");
    return 0;
}

}  // namespace

extern "C" {

int gsg_mg_create(gsg_plan* plan, int rank, int nranks, gsg_mg** out) {
    GSG_TRY(check_plan(plan));
    if (!out) return fail(GSG_ERR_ARG, "null output");
    if (nranks > 1 && nranks > (1 << (MG_NFLAG - 1))) return fail(GSG_ERR_UNSUPPORTED, "too many ranks");
    GSG_TRY(gsg_plan_set_partition(plan, rank, nranks));
    std::unique_ptr<gsg_mg> M(new gsg_mg());
    M->plan = plan;
    M->rank = rank;
    M->nranks = nranks;
    M->bits = plan->part_bits;
    M->Npad = plan->S.Npad;
    const size_t flag_bytes = ((size_t)nranks * MG_NFLAG * sizeof(unsigned long long) + 255) & ~(size_t)255;
    M->slab_bytes = (size_t)MG_NVEC * M->Npad * sizeof(double) + flag_bytes;
    GSG_CUDA(cudaMalloc(&M->slab, M->slab_bytes));
    GSG_CUDA(cudaMemset(M->slab, 0, M->slab_bytes));
    M->peer_slab.assign(nranks, nullptr);
    M->opened.assign(nranks, 0);
    M->peer_slab[rank] = M->slab;
    GSG_TRY(M->seqbuf.resize(1));
    GSG_TRY(M->errbuf.resize(1));
    M->dev_seq = M->seqbuf.p;
    M->dev_err = M->errbuf.p;
    GSG_CUDA(cudaStreamCreateWithFlags(&M->comm, cudaStreamNonBlocking));
    GSG_CUDA(cudaEventCreateWithFlags(&M->ev_begin, cudaEventDisableTiming));
    GSG_CUDA(cudaEventCreateWithFlags(&M->ev_join, cudaEventDisableTiming));
    // cell lists
    const gsg::IndexSet& S = plan->S;
    const int64_t KDp = S.kDp;
    {
        std::vector<int> cells;
        for (const gsg::Block& b : S.blocks) {
            if (block_owner(*plan, b, -1) != rank) continue;
            M->owned_blocks.emplace_back(b.offset, b.poffset);
            M->owned_block_cells.push_back(b.ncells);
            for (int64_t c = 0; c < b.ncells; ++c) cells.push_back((int)(b.poffset / KDp + c));
        }
        M->n_owned = (int64_t)cells.size();
        GSG_TRY(M->owned_cells.upload(cells));
    }
    for (int j = 0; j < M->bits; ++j) {
        const int d1 = S.D - j;                     // 1-based partition dimension carrying bit j
        int64_t cnt = 0;
        int partner = -1;
        GSG_TRY(gsg_plan_partition_blocks(plan, 1, d1, nullptr, nullptr, &cnt, &partner));
        std::vector<int64_t> offs(std::max<int64_t>(cnt, 1)), sizes(std::max<int64_t>(cnt, 1));
        GSG_TRY(gsg_plan_partition_blocks(plan, 1, d1, offs.data(), sizes.data(), &cnt, &partner));
        std::vector<int> cells;
        for (int64_t i = 0; i < cnt; ++i)
            for (int64_t c = 0; c < sizes[i] / KDp; ++c) cells.push_back((int)(offs[i] / KDp + c));
        M->ex.emplace_back();
        gsg_mg::Ex& e = M->ex.back();
        e.d = d1 - 1;
        e.j = j;
        e.partner = partner;
        e.mybit = (rank >> j) & 1;
        e.ncells = (int64_t)cells.size();
        GSG_TRY(e.cells.upload(cells));
        cudaEvent_t ev = nullptr;
        GSG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        M->ev_x.push_back(ev);
    }
    *out = M.release();
    return 0;
}

int gsg_mg_destroy(gsg_mg* mg) {
    if (!mg) return 0;
    if (mg->plan) cudaSetDevice(mg->plan->device);
    cudaDeviceSynchronize();
    if (mg->step_exec) cudaGraphExecDestroy(mg->step_exec);
    for (int q = 0; q < (int)mg->peer_slab.size(); ++q)
        if (mg->opened[q] && mg->peer_slab[q]) cudaIpcCloseMemHandle(mg->peer_slab[q]);
    for (cudaEvent_t ev : mg->ev_x) cudaEventDestroy(ev);
    if (mg->ev_begin) cudaEventDestroy(mg->ev_begin);
    if (mg->ev_join) cudaEventDestroy(mg->ev_join);
    if (mg->comm) cudaStreamDestroy(mg->comm);
    if (mg->slab) cudaFree(mg->slab);
    delete mg;
    return 0;
}

int gsg_mg_ipc_handle(gsg_mg* mg, void* handle64) {
    GSG_TRY(mg_check(mg));
    if (!handle64) return fail(GSG_ERR_ARG, "null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == GSG_MG_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    GSG_CUDA(cudaIpcGetMemHandle(&h, mg->slab));
    std::memcpy(handle64, &h, sizeof(h));
    return 0;
}

// multi-process: handles = nranks consecutive 64-byte IPC handles in rank order (this rank's own entry is ignored)
int gsg_mg_connect_ipc(gsg_mg* mg, const void* handles) {
    GSG_TRY(mg_check(mg));
    if (!handles) return fail(GSG_ERR_ARG, "null handles");
    for (int q = 0; q < mg->nranks; ++q) {
        if (q == mg->rank) continue;
        bool needed = false;
        for (const gsg_mg::Ex& e : mg->ex) needed = needed || e.partner == q;
        if (!needed) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)q * GSG_MG_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        GSG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        mg->peer_slab[q] = static_cast<double*>(p);
        mg->opened[q] = 1;
    }
    mg->connected = true;
    return 0;
}

// one process: all ranks' handles at once (devices may differ -> peer access is enabled both ways)
int gsg_mg_connect_local(gsg_mg* const* all, int nranks) {
    if (!all || nranks < 1) return fail(GSG_ERR_ARG, "bad argument");
    for (int r = 0; r < nranks; ++r)
        if (!all[r] || all[r]->rank != r || all[r]->nranks != nranks) return fail(GSG_ERR_ARG, "handles must be in rank order");
    for (int r = 0; r < nranks; ++r) {
        gsg_mg& M = *all[r];
        GSG_CUDA(cudaSetDevice(M.plan->device));
        for (int q = 0; q < nranks; ++q) {
            if (q == r) continue;
            const int dq = all[q]->plan->device;
            if (dq != M.plan->device) {
                int can = 0;
                GSG_CUDA(cudaDeviceCanAccessPeer(&can, M.plan->device, dq));
                if (!can) return fail(GSG_ERR_UNSUPPORTED, "devices cannot access each other's memory");
                const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(GSG_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            M.peer_slab[q] = all[q]->slab;
        }
        M.connected = true;
    }
    return 0;
}

// state in: the owned blocks of a full reference-layout HOST vector -> u; then READY for the first RHS
int gsg_mg_set_state(gsg_mg* mg, const double* u_host) {
    GSG_TRY(mg_check(mg));
    if (!u_host) return fail(GSG_ERR_ARG, "null state");
    if (!mg->connected && mg->nranks > 1) return fail(GSG_ERR_ARG, "connect the ranks first");
    gsg_plan& pl = *mg->plan;
    const gsg::IndexSet& S = pl.S;
    double* u = mg->vec(mg->rank, 0);
    for (size_t i = 0; i < mg->owned_blocks.size(); ++i) {
        const double* src = u_host + mg->owned_blocks[i].first;
        double* dst = u + mg->owned_blocks[i].second;
        if (S.kD == S.kDp) {
            GSG_CUDA(cudaMemcpyAsync(dst, src, (size_t)mg->owned_block_cells[i] * S.kD * sizeof(double), cudaMemcpyHostToDevice, pl.stream));
        } else {
            GSG_CUDA(cudaMemcpy2DAsync(dst, (size_t)S.kDp * sizeof(double), src, (size_t)S.kD * sizeof(double),
                                       (size_t)S.kD * sizeof(double), (size_t)mg->owned_block_cells[i], cudaMemcpyHostToDevice, pl.stream));
        }
    }
    // a new state invalidates every READY value signalled so far: all ranks skip one virtual step of RHS numbers
    for (int i = 0; i < 4; ++i) mg_bump_kernel<<<1, 1, 0, pl.stream>>>(mg->dev_seq);
    g_launches.fetch_add(4, std::memory_order_relaxed);
    for (gsg_mg::Ex& e : mg->ex)
        if (e.mybit == 1 && e.ncells > 0) GSG_TRY(mg_signal(*mg, pl.stream, e.partner, 0, 1));
    return 0;
}

// state out: the owned blocks of u are written into a full reference-layout HOST vector (other entries untouched)
int gsg_mg_get_state(gsg_mg* mg, double* u_host) {
    GSG_TRY(mg_check(mg));
    if (!u_host) return fail(GSG_ERR_ARG, "null state");
    gsg_plan& pl = *mg->plan;
    const gsg::IndexSet& S = pl.S;
    const double* u = mg->vec(mg->rank, 0);
    for (size_t i = 0; i < mg->owned_blocks.size(); ++i) {
        double* dst = u_host + mg->owned_blocks[i].first;
        const double* src = u + mg->owned_blocks[i].second;
        if (S.kD == S.kDp) {
            GSG_CUDA(cudaMemcpyAsync(dst, src, (size_t)mg->owned_block_cells[i] * S.kD * sizeof(double), cudaMemcpyDeviceToHost, pl.stream));
        } else {
            GSG_CUDA(cudaMemcpy2DAsync(dst, (size_t)S.kD * sizeof(double), src, (size_t)S.kDp * sizeof(double),
                                       (size_t)S.kD * sizeof(double), (size_t)mg->owned_block_cells[i], cudaMemcpyDeviceToHost, pl.stream));
        }
    }
    GSG_CUDA(cudaStreamSynchronize(pl.stream));
    return mg_check_err(*mg);
}

int gsg_mg_owned_fraction(gsg_mg* mg, double* frac_out, int64_t* exchange_bytes_per_rhs_out) {
    if (!mg || !mg->plan) return fail(GSG_ERR_ARG, "null multi-GPU handle");
    if (frac_out) *frac_out = (double)mg->n_owned * mg->plan->S.kDp / (double)mg->Npad;
    if (exchange_bytes_per_rhs_out) {
        int64_t b = 0;
        for (const gsg_mg::Ex& e : mg->ex) b += 8 * e.ncells * mg->plan->S.kDp;      // one pull per partition dimension
        *exchange_bytes_per_rhs_out = b;
    }
    return 0;
}

// nsteps RK4 steps of u' = -sum_d a_d D_d u on this rank's share; asynchronous on the plan's stream.  Every rank
// of the partition must make the same call (one process per GPU, or one host thread per rank).
int gsg_mg_rk4_advect(gsg_mg* mg, const double* a, double dt, int64_t nsteps) {
    GSG_TRY(mg_check(mg));
    if (!a || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    if (!mg->connected && mg->nranks > 1) return fail(GSG_ERR_ARG, "connect the ranks first");
    gsg_mg& M = *mg;
    gsg_plan& pl = *M.plan;
    double c[16];
    for (int d = 0; d < pl.S.D; ++d) c[d] = -a[d];
    static const bool no_graph = getenv("GSG_MG_NO_GRAPH") != nullptr;
    const bool use_graph = !no_graph && nsteps >= 3 && !pl.prof_on && !pl.dbg;
    int64_t s = 0;
    if (use_graph) {
        std::vector<double> av(a, a + pl.S.D);
        if (!M.step_exec || M.graph_dt != dt || M.graph_a != av) {
            if (M.step_exec) { cudaGraphExecDestroy(M.step_exec); M.step_exec = nullptr; }
            GSG_TRY(mg_step(M, c, dt));                       // eager first step (configures kernel attributes)
            ++s;
            const int64_t l0 = g_launches.load();
            cudaGraph_t graph = nullptr;
            GSG_CUDA(cudaStreamBeginCapture(pl.stream, cudaStreamCaptureModeThreadLocal));
            const int rc = mg_step(M, c, dt);
            const cudaError_t ce = cudaStreamEndCapture(pl.stream, &graph);
            if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(GSG_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            M.graph_launches = g_launches.load() - l0;
            g_launches.fetch_sub(M.graph_launches, std::memory_order_relaxed);      // the capture launched nothing
            const cudaError_t ie = cudaGraphInstantiate(&M.step_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return fail(GSG_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
            M.graph_dt = dt;
            M.graph_a = av;
        }
        for (; s < nsteps; ++s) {
            GSG_CUDA(cudaGraphLaunch(M.step_exec, pl.stream));
            g_launches.fetch_add(M.graph_launches, std::memory_order_relaxed);
        }
        return 0;
    }
    for (; s < nsteps; ++s) GSG_TRY(mg_step(M, c, dt));
    return 0;
}

// One host thread drives every rank of the partition (one process holding N devices, or N virtual ranks on one
// device): the phases of all ranks are enqueued in lockstep, so every wait kernel is enqueued after the signal
// it waits for (no deadlock even when the ranks' streams share hardware queues).
int gsg_mg_rk4_advect_all(gsg_mg* const* all, int nranks, const double* a, double dt, int64_t nsteps) {
    if (!all || nranks < 1 || !a || nsteps < 0) return fail(GSG_ERR_ARG, "bad argument");
    for (int r = 0; r < nranks; ++r) {
        GSG_TRY(mg_check(all[r]));
        if (!all[r]->connected && nranks > 1) return fail(GSG_ERR_ARG, "connect the ranks first");
    }
    double c[16];
    const int D = all[0]->plan->S.D;
    for (int d = 0; d < D; ++d) c[d] = -a[d];
    for (int64_t s = 0; s < nsteps; ++s) {
        for (int i = 0; i < 4; ++i) {
            for (int r = 0; r < nranks; ++r) { GSG_CUDA(cudaSetDevice(all[r]->plan->device)); GSG_TRY(mg_rhs_phase_a(*all[r], i, i + 1, c)); }
            for (int r = 0; r < nranks; ++r) { GSG_CUDA(cudaSetDevice(all[r]->plan->device)); GSG_TRY(mg_rhs_phase_b(*all[r], i, i + 1, c)); }
            for (int r = 0; r < nranks; ++r) { GSG_CUDA(cudaSetDevice(all[r]->plan->device)); GSG_TRY(mg_rhs_phase_c(*all[r], i + 1, c, i < 3)); }
        }
        for (int r = 0; r < nranks; ++r) { GSG_CUDA(cudaSetDevice(all[r]->plan->device)); GSG_TRY(mg_combine(*all[r], dt)); }
    }
    return 0;
}

int gsg_mg_sync(gsg_mg* mg) {
    GSG_TRY(mg_check(mg));
    GSG_CUDA(cudaStreamSynchronize(mg->plan->stream));
    GSG_CUDA(cudaStreamSynchronize(mg->comm));
    return mg_check_err(*mg);
}

}  // extern "C"
