// rowtile.inl -- host side of the row-tile kernel for the long-pole classes (kernels.cuh, sweep_rowtile_kernel):
// the TILE PROGRAM of a pole class and the per-direction work lists.  Included by gsg_b200.cu.
//
// A pole of class p is multiplied by the principal sub-block H_p = H[0:N', 0:N'] (N' = K 2^p), a K x K block matrix
// over the 2^p one-dimensional cells whose block (q, r) is stored iff the closed supports of the hierarchical cells
// q and r intersect or touch periodically.  For an ITEM (fixed cells of the other dimensions) every 1-D cell is one
// contiguous multi-cell of k^D doubles that holds that cell's entries of ALL k^(D-1) poles of the item -- so the unit
// of data movement is the whole multi-cell (one TMA bulk copy), and lanes = poles.  A TILE is a set of x cells that
// fits shared memory together with the block records that use only those cells:
//   * a class whose cells all fit is ONE tile (k = 3, D = 6: p = 4, 16 cells = 93 KB);
//   * otherwise the bottom t levels are cut into subtrees; the tile of a subtree holds the subtree's cells plus
//     every cell its rows touch (ancestors, periodic / boundary neighbours) and computes (a) the subtree's rows
//     completely and (b) the contributions of the subtree's cells to the rows ABOVE (partial sums, reduced into y
//     at the L2); the rows and columns above the subtrees form the pattern of class l0 - 1 and are tiled
//     recursively -- those rows only ever receive partial sums.
// Every stored block is used by exactly one tile (checked on the CPU: tests/test_rowtile_program.py replays the
// program in numpy against H_p x).
namespace {

struct RTProgram {
    std::vector<RTTile> tiles;
    std::vector<RTRow> rows;
    std::vector<unsigned char> recs;
    int rec_bytes = 0;
    // per class p: tile range and the rows that receive partial sums (must be zeroed before a beta = 0 sweep)
    std::vector<int> cls_first, cls_count;
    std::vector<std::vector<int>> cls_partial_q;
};

struct RTBuilder {
    const std::vector<int>& rowptr;
    const std::vector<int>& col;
    const std::vector<double>& val;
    int KK2, K, KDp, nrg;
    size_t budget;
    RTProgram& out;
    std::vector<int> blk;          // dense NQ x NQ: index of the stored block or -1
    int NQ = 0;
    std::vector<char> partial_q;

    int level_of(int q) const { return q == 0 ? 0 : 32 - __builtin_clz((unsigned)q); }

    size_t tile_bytes(size_t nx, size_t nrec) const { return 64 + nx * (size_t)KDp * 8 + nrec * (size_t)out.rec_bytes; }

    // one tile: complete rows `crow` (all their blocks with columns < nqt), partial rows = rows < nupper restricted to
    // the columns in `pcols`
    bool emit_tile(const std::vector<int>& crow, bool crow_partial, int nqt, int nupper, const std::vector<int>& pcols, bool dry,
                   size_t* bytes_out) {
        std::vector<int> slot(NQ, -1), xs;
        auto need = [&](int r) { if (slot[r] < 0) { slot[r] = (int)xs.size(); xs.push_back(r); } };
        struct Rec { int q, r, b; };
        std::vector<std::vector<Rec>> rws;       // records by row
        std::vector<int> rq;
        std::vector<char> rpart;
        for (int q : crow) {
            std::vector<Rec> v;
            for (int r = 0; r < nqt; ++r)
                if (blk[(size_t)q * NQ + r] >= 0) { need(r); v.push_back(Rec{q, r, blk[(size_t)q * NQ + r]}); }
            if (!v.empty()) { rws.push_back(std::move(v)); rq.push_back(q); rpart.push_back(crow_partial ? 1 : 0); }
        }
        for (int r : pcols) need(r);
        for (int u = 0; u < nupper; ++u) {
            std::vector<Rec> v;
            for (int r : pcols)
                if (blk[(size_t)u * NQ + r] >= 0) v.push_back(Rec{u, r, blk[(size_t)u * NQ + r]});
            if (!v.empty()) { rws.push_back(std::move(v)); rq.push_back(u); rpart.push_back(1); }
        }
        size_t nrec = 0;
        for (const auto& v : rws) nrec += v.size();
        if (bytes_out) *bytes_out = tile_bytes(xs.size(), nrec);
        if (xs.size() > (size_t)RT_MAXX || tile_bytes(xs.size(), nrec) > budget) return false;
        if (dry) return true;
        // rows dealt to the row groups, longest first onto the least loaded group
        std::vector<int> order(rws.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rws[a].size() > rws[b].size(); });
        std::vector<std::vector<int>> grp(nrg);
        std::vector<size_t> load(nrg, 0);
        for (int i : order) {
            int g = 0;
            for (int h = 1; h < nrg; ++h)
                if (load[h] < load[g]) g = h;
            grp[g].push_back(i);
            load[g] += rws[i].size() + 2;          // + the row's epilogue
        }
        RTTile T;
        std::memset(&T, 0, sizeof(T));
        T.nx = (int)xs.size();
        for (size_t i = 0; i < xs.size(); ++i) T.xq[i] = xs[i];
        T.rec0 = (int)(out.recs.size() / out.rec_bytes);
        T.row0 = (int)out.rows.size();
        int rel = 0, nrow = 0;
        for (int g = 0; g < nrg; ++g) {
            for (int i : grp[g]) {
                RTRow R;
                R.q = rq[i];
                R.rb = rel;
                R.partial = rpart[i];
                for (const Rec& rc : rws[i]) {
                    const size_t o = out.recs.size();
                    out.recs.resize(o + out.rec_bytes, 0);
                    std::memcpy(out.recs.data() + o, val.data() + (size_t)rc.b * KK2, (size_t)K * K * 8);
                    const int meta[2] = {slot[rc.r] * KDp * 8, 0};
                    std::memcpy(out.recs.data() + o + (size_t)K * K * 8, meta, 8);
                    ++rel;
                }
                R.re = rel;
                out.rows.push_back(R);
                if (rpart[i]) partial_q[rq[i]] = 1;
                ++nrow;
            }
            T.rg_end[g] = nrow;
        }
        for (int g = nrg; g < RT_MAXRG; ++g) T.rg_end[g] = nrow;
        T.nrec = rel;
        out.tiles.push_back(T);
        return true;
    }

    // pattern of the cells < 2^ptop (rows and columns); all_partial: these rows also receive sums from deeper tiles
    int emit(int ptop, bool all_partial) {
        const int nqt = 1 << ptop;
        std::vector<int> all(nqt);
        for (int q = 0; q < nqt; ++q) all[q] = q;
        if (emit_tile(all, all_partial, nqt, 0, {}, true, nullptr)) {
            emit_tile(all, all_partial, nqt, 0, {}, false, nullptr);
            return 0;
        }
        for (int t = std::min(ptop - 1, 6); t >= 1; --t) {
            const int l0 = ptop - t + 1;                    // subtree roots: the 2^(l0-1) cells of level l0
            const int nupper = 1 << (l0 - 1);               // cells of the levels < l0
            bool fits = true;
            std::vector<std::vector<int>> subs(nupper);
            for (int c0 = 0; c0 < nupper && fits; ++c0) {
                for (int l = l0; l <= ptop; ++l) {
                    const int ncl = 1 << (l - l0);
                    for (int i = 0; i < ncl; ++i) subs[c0].push_back((1 << (l - 1)) + c0 * ncl + i);
                }
                fits = emit_tile(subs[c0], all_partial, nqt, nupper, subs[c0], true, nullptr);
            }
            if (!fits) continue;
            for (int c0 = 0; c0 < nupper; ++c0) emit_tile(subs[c0], all_partial, nqt, nupper, subs[c0], false, nullptr);
            return emit(l0 - 1, true);
        }
        return fail(GSG_ERR_UNSUPPORTED, "row-tile program: no subtree tiling fits the shared-memory budget");
    }
};

// tile programs of the classes pmin..n (block CSR of the 1-D matrix over 2^n cells)
int rt_build_program(const std::vector<int>& rowptr, const std::vector<int>& col, const std::vector<double>& val, int KK2,
                     int K, int KDp, int n, int pmin, size_t budget, int nrg, RTProgram& out) {
    out = RTProgram();
    out.rec_bytes = (K * K * 8 + 8 + 15) & ~15;
    out.cls_first.assign(n + 1, 0);
    out.cls_count.assign(n + 1, 0);
    out.cls_partial_q.assign(n + 1, {});
    if (nrg < 1 || nrg > RT_MAXRG) return fail(GSG_ERR_ARG, "row-tile program: row groups out of range");
    for (int p = pmin; p <= n; ++p) {
        RTBuilder B{rowptr, col, val, KK2, K, KDp, nrg, budget, out, {}, 0, {}};
        B.NQ = 1 << p;
        B.blk.assign((size_t)B.NQ * B.NQ, -1);
        for (int q = 0; q < B.NQ; ++q)
            for (int b = rowptr[q]; b < rowptr[q + 1]; ++b)
                if (col[b] < B.NQ) B.blk[(size_t)q * B.NQ + col[b]] = b;
        B.partial_q.assign(B.NQ, 0);
        out.cls_first[p] = (int)out.tiles.size();
        GSG_TRY(B.emit(p, false));
        out.cls_count[p] = (int)out.tiles.size() - out.cls_first[p];
        for (int q = 0; q < B.NQ; ++q)
            if (B.partial_q[q]) out.cls_partial_q[p].push_back(q);
    }
    return 0;
}

}  // namespace
