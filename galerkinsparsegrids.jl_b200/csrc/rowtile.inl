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
// Inside a tile the rows are collected into ROW GROUPS of up to RT_R rows with overlapping column sets (siblings
// share their ancestors, ancestors share their descendants): one record per distinct x cell of the group carries the
// K x K blocks of all its rows, so the kernel reads x once per group instead of once per row.
// Every stored block is used by exactly one tile (checked on the CPU: tests/test_rowtile_program.py replays the
// program in numpy against H_p x).
namespace {

struct RTProgram {
    std::vector<RTTile> tiles;
    std::vector<RTGroup> groups;
    std::vector<unsigned char> blob;     // per tile: the records of its groups, {x cell offset, row mask} + one K x K block per mask bit
    // per class p: tile range and the rows that receive partial sums (must be zeroed before a beta = 0 sweep)
    std::vector<int> cls_first, cls_count;
    std::vector<std::vector<int>> cls_partial_q;
};

struct RTBuilder {
    const std::vector<int>& rowptr;
    const std::vector<int>& col;
    const std::vector<double>& val;
    int KK2, K, KDp, nrg;
    size_t budget;                 // shared memory of a CTA: barriers + 2 * nrg staging cells + x cells + records
    RTProgram& out;
    std::vector<int> blk;          // dense NQ x NQ: index of the stored block or -1
    int NQ = 0;
    std::vector<char> partial_q;

    static int level_of(int q) { return q == 0 ? 0 : 32 - __builtin_clz((unsigned)q); }

    size_t tile_bytes(size_t nx, size_t blob_bytes) const {
        return 64 + (size_t)2 * nrg * KDp * 8 + nx * (size_t)KDp * 8 + blob_bytes;
    }

    struct Rec { int q, r, b; };
    struct Row { int q; char partial; std::vector<Rec> recs; };

    // rows that share most of their columns are accumulated together (up to RT_R): greedy, seeded in (partial, level,
    // cell) order, each time adding the row with the largest column overlap
    std::vector<std::vector<int>> group_rows(const std::vector<Row>& rows, std::vector<int> left) const {
        std::stable_sort(left.begin(), left.end(), [&](int a, int b) {
            if (rows[a].partial != rows[b].partial) return rows[a].partial < rows[b].partial;
            if (level_of(rows[a].q) != level_of(rows[b].q)) return level_of(rows[a].q) < level_of(rows[b].q);
            return rows[a].q < rows[b].q;
        });
        std::vector<std::vector<int>> groups;
        std::vector<char> in(NQ);
        while (!left.empty()) {
            std::vector<int> g{left.front()};
            left.erase(left.begin());
            std::fill(in.begin(), in.end(), 0);
            for (const Rec& rc : rows[g[0]].recs) in[rc.r] = 1;
            while ((int)g.size() < RT_R && !left.empty()) {
                int best = -1, bs = 0;
                for (size_t i = 0; i < left.size(); ++i) {
                    if (rows[left[i]].partial != rows[g[0]].partial) continue;
                    int inter = 0;
                    for (const Rec& rc : rows[left[i]].recs) inter += in[rc.r];
                    if (inter > bs) { bs = inter; best = (int)i; }
                }
                if (best < 0) break;
                for (const Rec& rc : rows[left[best]].recs) in[rc.r] = 1;
                g.push_back(left[best]);
                left.erase(left.begin() + best);
            }
            groups.push_back(std::move(g));
        }
        return groups;
    }

    // one tile: complete rows `crow` (all their blocks with columns < nqt), partial rows = rows < nupper restricted to
    // the columns in `pcols`
    bool emit_tile(const std::vector<int>& crow, bool crow_partial, int nqt, int nupper, const std::vector<int>& pcols, bool dry) {
        std::vector<int> slot(NQ, -1), xs;
        auto need = [&](int r) { if (slot[r] < 0) { slot[r] = (int)xs.size(); xs.push_back(r); } };
        std::vector<Row> rows;
        for (int q : crow) {
            Row R{q, (char)(crow_partial ? 1 : 0), {}};
            for (int r = 0; r < nqt; ++r)
                if (blk[(size_t)q * NQ + r] >= 0) { need(r); R.recs.push_back(Rec{q, r, blk[(size_t)q * NQ + r]}); }
            if (!R.recs.empty()) rows.push_back(std::move(R));
        }
        for (int r : pcols) need(r);
        for (int u = 0; u < nupper; ++u) {
            Row R{u, 1, {}};
            for (int r : pcols)
                if (blk[(size_t)u * NQ + r] >= 0) R.recs.push_back(Rec{u, r, blk[(size_t)u * NQ + r]});
            if (!R.recs.empty()) rows.push_back(std::move(R));
        }
        if (xs.size() > (size_t)RT_MAXX) return false;
        // Rows are first dealt to the row groups (warps along the rows) so that every row group gets about the same
        // number of blocks (measured: a warp's time in a tile is proportional to its blocks and the tile ends with its
        // slowest warp; grouping similar rows FIRST left the coarse, long rows together in one warp: 8.2 vs 5.7 us
        // slowest vs mean).  To keep similar rows together all the same, the rows are ordered by the position of their
        // cell's centre (an in-order walk of the cell tree: a cell sits between its descendants) and that order is cut
        // into nrg contiguous pieces of equal cost.  Inside a row group the rows are then collected into groups of
        // <= RT_R rows by column overlap.
        std::vector<std::vector<int>> rows_of(nrg);
        {
            auto centre = [&](int q) {
                const int l = level_of(q);
                if (l == 0) return 0.5;
                const int nc = 1 << (l - 1);
                return ((q - nc) + 0.5) / nc;
            };
            std::vector<int> order(rows.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                const double ca = centre(rows[a].q), cb = centre(rows[b].q);
                if (ca != cb) return ca < cb;
                return level_of(rows[a].q) < level_of(rows[b].q);
            });
            size_t total = 0;
            for (const Row& R : rows) total += R.recs.size() + 3;          // + the row's epilogue
            size_t run = 0;
            for (int i : order) {
                const size_t c = rows[i].recs.size() + 3;
                int g = (int)(((run + c / 2) * nrg) / total);               // the piece the row's midpoint falls into
                g = std::min(g, nrg - 1);
                rows_of[g].push_back(i);
                run += c;
            }
        }
        std::vector<std::vector<int>> grp;
        std::vector<std::vector<int>> bin(nrg);              // groups of every row group
        for (int g = 0; g < nrg; ++g)
            for (auto& gg : group_rows(rows, rows_of[g])) {
                bin[g].push_back((int)grp.size());
                grp.push_back(std::move(gg));
            }
        std::vector<size_t> gbytes(grp.size()), gcost(grp.size());
        size_t blob_bytes = 0;
        std::vector<int> cnt(NQ);
        for (size_t g = 0; g < grp.size(); ++g) {
            std::fill(cnt.begin(), cnt.end(), 0);
            size_t nb = 0, nr = 0;
            for (int i : grp[g])
                for (const Rec& rc : rows[i].recs) { nr += cnt[rc.r]++ == 0; ++nb; }
            gbytes[g] = nr * 8 + nb * (size_t)K * K * 8;
            blob_bytes += gbytes[g];
        }
        blob_bytes = (blob_bytes + 15) & ~size_t(15);
        if (tile_bytes(xs.size(), blob_bytes) > budget) return false;
        if (dry) return true;
        RTTile T;
        std::memset(&T, 0, sizeof(T));
        T.nx = (int)xs.size();
        for (size_t i = 0; i < xs.size(); ++i) T.xq[i] = xs[i];
        T.rec_ofs = (int)out.blob.size();
        T.rec_bytes = (int)blob_bytes;
        T.grp0 = (int)out.groups.size();
        int ngrp = 0;
        for (int g = 0; g < nrg; ++g) {
            for (int gi : bin[g]) {
                RTGroup G;
                std::memset(&G, 0, sizeof(G));
                for (int r = 0; r < RT_R; ++r) G.q[r] = -1;
                G.rofs = (int)(out.blob.size() - (size_t)T.rec_ofs);
                // the group's columns in ascending x slot order; block (row r of the group, column) or none
                std::vector<std::array<int, RT_R>> at(xs.size());
                for (auto& a : at) a.fill(-1);
                for (size_t r = 0; r < grp[gi].size(); ++r) {
                    const Row& R = rows[grp[gi][r]];
                    G.q[r] = R.q;
                    if (R.partial) { G.partial |= 1 << r; partial_q[R.q] = 1; }
                    for (const Rec& rc : R.recs) at[slot[rc.r]][r] = rc.b;
                }
                for (size_t sl = 0; sl < xs.size(); ++sl) {
                    int mask = 0;
                    for (int r = 0; r < RT_R; ++r)
                        if (at[sl][r] >= 0) mask |= 1 << r;
                    if (!mask) continue;
                    // the barrier the x cell arrives on rides in bits 8.. of the mask word
                    const int hd[2] = {(int)sl * KDp * 8, mask | (int)((sl * RT_NBAR) / xs.size()) << 8};
                    size_t o = out.blob.size();
                    out.blob.resize(o + 8);
                    std::memcpy(out.blob.data() + o, hd, 8);
                    for (int r = 0; r < RT_R; ++r) {
                        if (at[sl][r] < 0) continue;
                        o = out.blob.size();
                        out.blob.resize(o + (size_t)K * K * 8);
                        std::memcpy(out.blob.data() + o, val.data() + (size_t)at[sl][r] * KK2, (size_t)K * K * 8);
                    }
                    ++G.nrec;
                }
                out.groups.push_back(G);
                ++ngrp;
            }
            T.rg_end[g] = ngrp;
        }
        for (int g = nrg; g < RT_MAXRG; ++g) T.rg_end[g] = ngrp;
        out.blob.resize((size_t)T.rec_ofs + blob_bytes, 0);
        out.tiles.push_back(T);
        return true;
    }

    // pattern of the cells < 2^ptop (rows and columns); all_partial: these rows also receive sums from deeper tiles
    int emit(int ptop, bool all_partial) {
        const int nqt = 1 << ptop;
        std::vector<int> all(nqt);
        for (int q = 0; q < nqt; ++q) all[q] = q;
        if (emit_tile(all, all_partial, nqt, 0, {}, true)) {
            emit_tile(all, all_partial, nqt, 0, {}, false);
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
                fits = emit_tile(subs[c0], all_partial, nqt, nupper, subs[c0], true);
            }
            if (!fits) continue;
            for (int c0 = 0; c0 < nupper; ++c0) emit_tile(subs[c0], all_partial, nqt, nupper, subs[c0], false);
            return emit(l0 - 1, true);
        }
        return fail(GSG_ERR_UNSUPPORTED, "row-tile program: no subtree tiling fits the shared-memory budget");
    }
};

// tile programs of the classes pmin..n (block CSR of the 1-D matrix over 2^n cells)
int rt_build_program(const std::vector<int>& rowptr, const std::vector<int>& col, const std::vector<double>& val, int KK2,
                     int K, int KDp, int n, int pmin, size_t budget, int nrg, RTProgram& out) {
    out = RTProgram();
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

// Lane order of the poles: entry j = (pole warp * C + c) * 32 + lane of the table holds the in-cell offset
// a + K*A*b of the pole that lane works on in slot c.  Shared memory serves a 64-bit warp access in two half-warp
// phases, 16 banks of 8 bytes each, so the poles are dealt such that the 16 lanes of every half-warp fall into 16
// different banks (offset mod 16) -- natural orders conflict whenever a wraps inside a half-warp (measured with
// tools/lds_bench.cu: 2-way for A = 3, 9, 27 at k = 3).  Padding lanes repeat a pole of their own half-warp (a
// broadcast, no extra wavefront) and are stored as ~offset.
std::vector<int> rt_pole_order(int K, int A, int PI, int nslots) {
    const int nhw = nslots / 16;
    std::vector<std::vector<int>> bank(16);
    const int B = PI / A;
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < A; ++a) bank[(a + K * A * b) & 15].push_back(a + K * A * b);
    std::vector<std::vector<int>> hw(nhw);
    std::vector<std::array<char, 16>> used(nhw);
    for (auto& u : used) u.fill(0);
    std::vector<int> left;
    for (int k16 = 0; k16 < 16; ++k16) {
        // the poles of one bank go to different half-warps; a bank with more poles than half-warps overflows
        for (size_t i = 0; i < bank[k16].size(); ++i) {
            if ((int)i < nhw) { hw[i].push_back(bank[k16][i]); used[i][k16] = 1; }
            else left.push_back(bank[k16][i]);
        }
    }
    for (int o : left) {                   // overflow: a half-warp that has room and not this bank yet, else the emptiest
        int best = -1;
        for (int h = 0; h < nhw; ++h)
            if (hw[h].size() < 16 && !used[h][o & 15] && (best < 0 || hw[h].size() < hw[best].size())) best = h;
        if (best < 0) {
            best = 0;
            for (int h = 1; h < nhw; ++h)
                if (hw[h].size() < hw[best].size()) best = h;
        }
        hw[best].push_back(o);
        used[best][o & 15] = 1;
    }
    // (every bank list starts at half-warp 0, so the first half-warps are full and the last ones short; padding lanes
    // repeat entry 0 of their half-warp)
    std::vector<int> tab(nslots);
    for (int h = 0; h < nhw; ++h) {
        for (int l = 0; l < 16; ++l) {
            if (l < (int)hw[h].size()) tab[h * 16 + l] = hw[h][l];
            else tab[h * 16 + l] = hw[h].empty() ? ~0 : ~hw[h][0];
        }
    }
    return tab;
}

}  // namespace
