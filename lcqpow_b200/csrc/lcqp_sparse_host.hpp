// lcqp_sparse_host.hpp -- host-side symbolic analysis of the OSQP-flavour path (plain C++17; used by lcqp_cabi.cu
// and by the CPU build of the device code in tests/emu).
//
// What it replaces.  The reference's OSQP run factors the quasi-definite KKT matrix
//     K = [ P + sigma I   A' ;  A   -diag(1/rho) ]          (/root/reference/external/osqp/src/kkt.c:6-177)
// with QDLDL after an AMD ordering (lin_sys/direct/qdldl/qdldl_interface.c:177-323, qdldl_sources/src/qdldl.c:11-233).
// Ordering, elimination tree and the pattern of L depend on the sparsity pattern only, and a batch shares its
// pattern: this file does that analysis ONCE per load, on the host, and emits flat index "programs" that every
// instance (one GPU thread each, lcqp_osqp.cuh) executes with its own values -- no integer workspace, no pattern
// logic and no divergence on the device.
//     ordering      multiple minimum degree on the elimination graph (bitset adjacency; independent sets per round)
//     etree + L     row-by-row reach through the elimination tree (the up-looking scheme of qdldl.c:72-233)
//     row program   for row k of L: the columns it touches in elimination order and the slot of L(k, c) in column c
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace lcqp {
namespace osq {

enum { KSRC_P = 0, KSRC_A = 1, KSRC_RHO = 2, KSRC_SIGMA = 3 };
inline int ksrc(int type, int idx) { return (type << 28) | idx; }

struct Symbolic {
    int n = 0, m = 0, N = 0;
    // P: upper triangle of Q in CSC (rows sorted); A: [A; L; R] in CSC.  *src: index into the caller's value array
    // (Qsrc: into the Q values; Asrc: (matrix << 28) | index, matrix 0 = A, 1 = L, 2 = R)
    std::vector<int> Pp, Pi, Psrc;
    std::vector<int> Ap, Ai, Asrc;
    // the full symmetric Q in CSC for the outer loop's products (strict lower entries are separate values of the
    // caller's array; the reference reads them too)
    std::vector<int> Qp, Qi, Qsrc;
    // permuted KKT (upper triangle, CSC), perm[new] = old, iperm[old] = new
    std::vector<int> perm, iperm, Kp, Ki, Ksrc;
    // L (unit lower, CSC, rows in increasing order) and the row programs
    std::vector<int> Lp, Li;
    std::vector<int> rp, rcol, rpos;   // row k: t in [rp[k], rp[k+1]): column rcol[t], slot rpos[t] of L
    std::vector<int> Lcol, Lrev;       // column of entry p; order of the backward sweep (columns descending, entries ascending)
    std::vector<int> Pcol, Acol, Qcol; // column of every entry of P, A, Q
    // the same entries listed by ROW (entry ids sorted by row, then column): a product accumulates every output element
    // in a register, in the reference's order of accumulation, and rows are independent of each other
    std::vector<int> ArP, ArE, PrP, PrE, QrP, QrE;
    // L by rows: row i holds slots [LrP[i], LrP[i+1]) of a second, row-major copy of the values; LrC = column of a slot;
    // rposr[t] = row-major slot of the entry that step t of the row programs produces
    std::vector<int> LrP, LrC, rposr;
    // level sets of the triangular solves: rows (forward) / columns (backward) of one level do not depend on each other
    std::vector<int> flP, flR, blP, blC;
    long long factor_flops = 0;        // multiply-subtracts of one numeric factorisation
    // STREAMED triangular solves (one warp per instance, lcqp_osqp_impl.inc stream_sweep): the entries of L in the order
    // in which a level-by-level sweep consumes them, cut into chunks of kStreamVals values that a warp brings into shared
    // memory ahead of use (cp.async.bulk + mbarrier), so that the only dependent loads of a sweep are shared-memory ones.
    //   index chunk c (kStreamIdx 16-bit words, shared by the batch): [nlev] then per level [r][E][r rows][E x r columns]
    //       a level here has r <= 32 independent chains (lane j owns chain j) of up to E entries each; entry e of lane j is
    //       value slot voff + e * r + j of value chunk c (voff runs over the chunk), its column is the matching index word;
    //       padding entries carry value 0 and column N + 2 (a scratch slot that holds 0)
    //   value chunk c (kStreamVals doubles, per instance): filled after every numeric factorisation from the row-major
    //       (forward sweep) / column-major (backward sweep) copy of L through fsSrc / bsSrc (-1: padding)
    // stream = 0 when a level does not fit a chunk (the sweeps then read L in place, level by level).
    // the numeric factorisation's updates ("pairs"), listed in the order the row programs perform them: step t of the row
    // programs subtracts L(i, c) * y_c for the entries p in [Lp[c], rpos[t]) of column c = rcol[t]; pair q in
    // [sOff[t], sOff[t+1]) is one of them: fpIdx[q] = p, fpLi[q] = Li[p], fpStep[q] = t - rp[row].  rowPair[k] = sOff[rp[k]].
    // (the warp build prefetches a row's steps and pairs in two rounds of loads instead of three per step)
    std::vector<int> sOff, fpIdx, fpLi, fpStep, rowPair;
    std::vector<int> fsI, bsI;         // index chunks, two 16-bit words per int
    std::vector<int> fsSrc, bsSrc;     // kStreamVals per chunk
    int fsChunks = 0, bsChunks = 0, stream = 0;
};
constexpr int kStreamVals = 512;       // doubles per value chunk (4 KB)
constexpr int kStreamIdx = 1280;       // 16-bit words per index chunk (2.5 KB)

// rows[k] = (row id, entries as (source slot, column)) of one sweep, grouped by level through lvP
struct StreamRow { int id; std::vector<std::pair<int, int>> ent; };
inline bool build_stream(int N, const std::vector<std::vector<StreamRow>>& levels, std::vector<int>& Ipack, std::vector<int>& Src, int& nchunks)
{
    std::vector<unsigned short> I;
    Src.clear();
    nchunks = 0;
    if (N + 3 > 65535) return false;
    size_t ibase = 0, vbase = 0;
    int iused = 0, vused = 0, nlev = 0;
    auto open_chunk = [&]() {
        ibase = I.size(); vbase = Src.size();
        I.resize(ibase + kStreamIdx, 0); Src.resize(vbase + kStreamVals, -1);
        iused = 1; vused = 0; nlev = 0; nchunks++;
    };
    auto close_chunk = [&]() { I[ibase] = (unsigned short)nlev; };
    bool open = false;
    for (const auto& lv : levels) {
        std::vector<const StreamRow*> rows;
        for (const auto& r : lv) if (!r.ent.empty()) rows.push_back(&r);
        std::stable_sort(rows.begin(), rows.end(), [](const StreamRow* a, const StreamRow* b) { return a->ent.size() > b->ent.size(); });
        for (size_t g = 0; g < rows.size();) {
            // the longest of the group (rows are sorted by length); shorter rows are padded with no-op entries
            const int E = (int)rows[g]->ent.size();
            if (E > 4) return false;   // (the device loop keeps a level's entries in registers: four per chain)
            // at most 32 rows (one per lane), fewer when the rows are long: the group must fit one chunk
            int r = (int)std::min<size_t>(32, rows.size() - g);
            while (r > 1 && (r * E > kStreamVals || 1 + 2 + r + r * E > kStreamIdx)) r--;
            const int ineed = 2 + r + r * E, vneed = r * E;
            if (1 + ineed > kStreamIdx || vneed > kStreamVals) return false;
            if (!open || iused + ineed > kStreamIdx || vused + vneed > kStreamVals) { if (open) close_chunk(); open_chunk(); open = true; }
            unsigned short* ip = I.data() + ibase + iused;
            ip[0] = (unsigned short)r; ip[1] = (unsigned short)E;
            for (int j = 0; j < r; j++) ip[2 + j] = (unsigned short)rows[g + j]->id;
            for (int e = 0; e < E; e++)
                for (int j = 0; j < r; j++) {
                    const auto& en = rows[g + j]->ent;
                    const bool real = e < (int)en.size();
                    ip[2 + r + e * r + j] = (unsigned short)(real ? en[e].second : N + 2);
                    Src[vbase + vused + e * r + j] = real ? en[e].first : -1;
                }
            iused += ineed; vused += vneed; nlev++;
            g += r;
        }
    }
    if (open) close_chunk();
    Ipack.assign((I.size() + 1) / 2, 0);
    if (!I.empty()) std::memcpy(Ipack.data(), I.data(), I.size() * sizeof(unsigned short));
    return true;
}

// minimum-degree ordering of a symmetric pattern given as adjacency lists (no self loops)
inline std::vector<int> min_degree_order(int N, const std::vector<std::vector<int>>& adj)
{
    const int W = (N + 63) / 64;
    std::vector<uint64_t> bits((size_t)N * W, 0);
    auto row = [&](int i) { return bits.data() + (size_t)i * W; };
    for (int i = 0; i < N; i++) for (int j : adj[i]) if (j != i) row(i)[j >> 6] |= 1ull << (j & 63);
    std::vector<int> deg(N), order;
    std::vector<char> gone(N, 0);
    auto popcnt = [&](int i) { int c = 0; for (int w = 0; w < W; w++) c += __builtin_popcountll(row(i)[w]); return c; };
    for (int i = 0; i < N; i++) deg[i] = popcnt(i);
    order.reserve(N);
    std::vector<int> nb;
    // MULTIPLE elimination: a round eliminates an independent set of nodes of minimum degree (a node is skipped when a
    // neighbour went in the same round).  On chain-like graphs (banded KKT systems) this halves the graph per round and
    // the elimination tree gets logarithmic height instead of linear -- the height is the number of levels, i.e. of
    // warp-wide synchronisations, of every triangular solve on the device.
    std::vector<int> blocked(N, -1);
    int round = 0;
    while ((int)order.size() < N) {
        int mind = 1 << 30;
        for (int i = 0; i < N; i++) if (!gone[i] && deg[i] < mind) mind = deg[i];
        for (int best = 0; best < N; best++) {
            if (gone[best] || deg[best] != mind || blocked[best] == round) continue;
            order.push_back(best);
            gone[best] = 1;
            nb.clear();
            for (int w = 0; w < W; w++) { uint64_t b = row(best)[w]; while (b) { const int t = __builtin_ctzll(b); nb.push_back(w * 64 + t); b &= b - 1; } }
            // the neighbours of the eliminated node become a clique; the node leaves their lists
            for (int a : nb) {
                uint64_t* ra = row(a);
                const uint64_t* rb = row(best);
                for (int w = 0; w < W; w++) ra[w] |= rb[w];
                ra[a >> 6] &= ~(1ull << (a & 63));
                ra[best >> 6] &= ~(1ull << (best & 63));
                deg[a] = popcnt(a);
                blocked[a] = round;
            }
        }
        round++;
    }
    return order;
}

// Nested dissection by breadth-first level structures: a connected piece is split at the middle level of the BFS from
// a pseudo-peripheral node, the two sides are ordered first (recursively), the separator last.  For chain-like graphs
// (banded KKT systems) the elimination tree gets height O(separator x log N) where minimum degree gives O(N).
inline std::vector<int> nested_dissection_order(int N, const std::vector<std::vector<int>>& adj, int leaf = 24)
{
    std::vector<int> order, tag(N, 0), dist(N, -1), queue;
    order.reserve(N);
    int next_tag = 1;
    // work items: (tag of the node set, emit separator afterwards?) -- an explicit stack of "ordered output" pieces
    struct Item { std::vector<int> nodes; bool emit_only; };
    std::vector<Item> stack;
    { Item all; all.nodes.resize(N); for (int i = 0; i < N; i++) all.nodes[i] = i; all.emit_only = false; stack.push_back(std::move(all)); }
    auto bfs = [&](int start, int t, std::vector<int>& out) {   // BFS inside the set tagged t; returns the visit order, dist[] filled
        out.clear();
        out.push_back(start); dist[start] = 0;
        for (size_t h = 0; h < out.size(); h++) { const int u = out[h]; for (int v : adj[u]) if (tag[v] == t && dist[v] < 0) { dist[v] = dist[u] + 1; out.push_back(v); } }
    };
    while (!stack.empty()) {
        Item it = std::move(stack.back());
        stack.pop_back();
        if (it.emit_only || (int)it.nodes.size() <= leaf) { for (int u : it.nodes) order.push_back(u); continue; }
        const int t = next_tag++;
        for (int u : it.nodes) { tag[u] = t; dist[u] = -1; }
        // one connected component at a time
        std::vector<int> comp, rest;
        bfs(it.nodes[0], t, comp);
        if (comp.size() < it.nodes.size()) {
            for (int u : it.nodes) if (dist[u] < 0) rest.push_back(u);
            for (int u : comp) dist[u] = -1;
            Item a; a.nodes = rest; a.emit_only = false;
            Item b; b.nodes = comp; b.emit_only = false;
            stack.push_back(std::move(a));
            stack.push_back(std::move(b));
            continue;
        }
        // pseudo-peripheral start: the last node of a BFS, twice
        int start = comp.back();
        for (int rep = 0; rep < 2; rep++) { for (int u : comp) dist[u] = -1; bfs(start, t, comp); start = comp.back(); }
        for (int u : comp) dist[u] = -1;
        bfs(start, t, comp);
        const int depth = dist[comp.back()];
        if (depth < 2) { for (int u : comp) order.push_back(u); continue; }
        const int mid = depth / 2;
        Item A, B, Sep;
        for (int u : comp) { if (dist[u] < mid) A.nodes.push_back(u); else if (dist[u] > mid) B.nodes.push_back(u); else Sep.nodes.push_back(u); }
        A.emit_only = false; B.emit_only = false; Sep.emit_only = true;
        // output order: A, B, separator -> push in reverse
        stack.push_back(std::move(Sep));
        stack.push_back(std::move(B));
        stack.push_back(std::move(A));
    }
    return order;
}

// Build everything from the patterns.  Qpat: full symmetric n x n pattern as (row, col, source index) triplets;
// Apat: m x n pattern of [A; L; R] as (row, col, source) triplets.
struct Trip { int r, c, src; };

inline void analyse_with(int n, int m, const std::vector<Trip>& Qpat, const std::vector<Trip>& Apat, Symbolic& S, int strategy)
{
    S.n = n; S.m = m; S.N = n + m;
    const int N = S.N;
    auto to_csc = [](int ncols, std::vector<Trip> t, std::vector<int>& p, std::vector<int>& i, std::vector<int>& s) {
        std::sort(t.begin(), t.end(), [](const Trip& a, const Trip& b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
        p.assign(ncols + 1, 0); i.clear(); s.clear();
        for (const Trip& e : t) { p[e.c + 1]++; i.push_back(e.r); s.push_back(e.src); }
        for (int c = 0; c < ncols; c++) p[c + 1] += p[c];
    };
    std::vector<Trip> up;
    for (const Trip& e : Qpat) if (e.r <= e.c) up.push_back(e);
    to_csc(n, up, S.Pp, S.Pi, S.Psrc);
    to_csc(n, Qpat, S.Qp, S.Qi, S.Qsrc);
    to_csc(n, Apat, S.Ap, S.Ai, S.Asrc);

    // KKT pattern (unpermuted, upper triangle): P + sigma I ; A' ; -1/rho
    struct KE { int r, c, src; };
    std::vector<KE> ke;
    std::vector<char> hasdiag(n, 0);
    for (int c = 0; c < n; c++) for (int p = S.Pp[c]; p < S.Pp[c + 1]; p++) { ke.push_back({S.Pi[p], c, ksrc(KSRC_P, p)}); if (S.Pi[p] == c) hasdiag[c] = 1; }
    for (int c = 0; c < n; c++) if (!hasdiag[c]) ke.push_back({c, c, ksrc(KSRC_SIGMA, c)});
    for (int c = 0; c < n; c++) for (int p = S.Ap[c]; p < S.Ap[c + 1]; p++) ke.push_back({c, n + S.Ai[p], ksrc(KSRC_A, p)});   // A'(c, i) sits at (c, n + i)
    for (int i = 0; i < m; i++) ke.push_back({n + i, n + i, ksrc(KSRC_RHO, i)});
    std::vector<std::vector<int>> adj(N);
    for (const KE& e : ke) if (e.r != e.c) { adj[e.r].push_back(e.c); adj[e.c].push_back(e.r); }
    S.perm = (strategy == 1) ? nested_dissection_order(N, adj) : min_degree_order(N, adj);
    S.iperm.assign(N, 0);
    for (int k = 0; k < N; k++) S.iperm[S.perm[k]] = k;
    // permuted upper triangle
    {
        std::vector<Trip> t;
        for (const KE& e : ke) {
            int a = S.iperm[e.r], b = S.iperm[e.c];
            if (a > b) std::swap(a, b);
            t.push_back({a, b, e.src});
        }
        to_csc(N, t, S.Kp, S.Ki, S.Ksrc);
    }
    // elimination tree and the pattern of every row of L (reach of the row's entries through the tree), qdldl.c:11-70
    std::vector<int> parent(N, -1), mark(N, -1);
    std::vector<std::vector<int>> rowpat(N);   // columns c < k with L(k, c) != 0, in elimination order
    std::vector<int> colcount(N, 0), stack, path;
    for (int k = 0; k < N; k++) {
        mark[k] = k;
        std::vector<int>& pat = rowpat[k];
        // QDLDL walks the entries of column k from first to last and pushes every new path in REVERSE on a buffer that
        // is then read back to front: the resulting order is a topological order of the reach
        std::vector<int> buf;
        for (int p = S.Kp[k]; p < S.Kp[k + 1]; p++) {
            int i = S.Ki[p];
            if (i == k) continue;
            path.clear();
            while (mark[i] != k) {
                if (parent[i] < 0) parent[i] = k;
                path.push_back(i);
                mark[i] = k;
                i = parent[i];
            }
            for (int q = (int)path.size() - 1; q >= 0; q--) buf.push_back(path[q]);
        }
        for (int q = (int)buf.size() - 1; q >= 0; q--) pat.push_back(buf[q]);
        for (int c : pat) colcount[c]++;
    }
    S.Lp.assign(N + 1, 0);
    for (int c = 0; c < N; c++) S.Lp[c + 1] = S.Lp[c] + colcount[c];
    S.Li.assign(S.Lp[N], 0);
    std::vector<int> fill(N, 0);
    S.rp.assign(N + 1, 0);
    S.rcol.clear(); S.rpos.clear();
    S.factor_flops = 0;
    for (int k = 0; k < N; k++) {
        for (int c : rowpat[k]) {
            const int pos = S.Lp[c] + fill[c];
            S.factor_flops += fill[c] + 2;
            S.rcol.push_back(c);
            S.rpos.push_back(pos);
            S.Li[pos] = k;
            fill[c]++;
        }
        S.rp[k + 1] = (int)S.rcol.size();
    }
    auto cols_of = [](const std::vector<int>& ptr, std::vector<int>& out) { out.assign(ptr.back(), 0); for (size_t c = 0; c + 1 < ptr.size(); c++) for (int p = ptr[c]; p < ptr[c + 1]; p++) out[p] = (int)c; };
    cols_of(S.Pp, S.Pcol); cols_of(S.Ap, S.Acol); cols_of(S.Qp, S.Qcol);
    auto by_rows = [](int nrows, const std::vector<int>& ptr, const std::vector<int>& idx, std::vector<int>& rp, std::vector<int>& re) {
        rp.assign(nrows + 1, 0); re.assign(idx.size(), 0);
        for (int r : idx) rp[r + 1]++;
        for (int r = 0; r < nrows; r++) rp[r + 1] += rp[r];
        std::vector<int> fillr(nrows, 0);
        for (size_t c = 0; c + 1 < ptr.size(); c++) for (int p = ptr[c]; p < ptr[c + 1]; p++) { const int r = idx[p]; re[rp[r] + fillr[r]++] = p; }   // columns ascend
    };
    by_rows(m, S.Ap, S.Ai, S.ArP, S.ArE);
    by_rows(n, S.Pp, S.Pi, S.PrP, S.PrE);
    by_rows(n, S.Qp, S.Qi, S.QrP, S.QrE);
    {
        std::vector<int> re;
        by_rows(N, S.Lp, S.Li, S.LrP, re);             // re[slot] = column-major entry id
        std::vector<int> slot_of(S.Li.size(), 0);
        S.LrC.assign(S.Li.size(), 0);
        std::vector<int> colof(S.Li.size(), 0);
        for (int c = 0; c < N; c++) for (int p = S.Lp[c]; p < S.Lp[c + 1]; p++) colof[p] = c;
        for (size_t sl = 0; sl < re.size(); sl++) { slot_of[re[sl]] = (int)sl; S.LrC[sl] = colof[re[sl]]; }
        S.rposr.assign(S.rpos.size(), 0);
        for (size_t t = 0; t < S.rpos.size(); t++) S.rposr[t] = slot_of[S.rpos[t]];
        // levels
        std::vector<int> lev(N, 0), blev(N, 0);
        int maxl = 0, maxb = 0;
        for (int i = 0; i < N; i++) { int l = 0; for (int sl = S.LrP[i]; sl < S.LrP[i + 1]; sl++) l = std::max(l, lev[S.LrC[sl]] + 1); lev[i] = l; maxl = std::max(maxl, l); }
        for (int i = N - 1; i >= 0; i--) { int l = 0; for (int p = S.Lp[i]; p < S.Lp[i + 1]; p++) l = std::max(l, blev[S.Li[p]] + 1); blev[i] = l; maxb = std::max(maxb, l); }
        S.flP.assign(1, 0); S.flR.clear();
        for (int l = 1; l <= maxl; l++) { for (int i = 0; i < N; i++) if (lev[i] == l) S.flR.push_back(i); S.flP.push_back((int)S.flR.size()); }
        S.blP.assign(1, 0); S.blC.clear();
        for (int l = 1; l <= maxb; l++) { for (int i = N - 1; i >= 0; i--) if (blev[i] == l) S.blC.push_back(i); S.blP.push_back((int)S.blC.size()); }
    }
    {
        // The streamed form of the two sweeps.  Unknown i of a sweep is a chain of subtractions x_i -= L(.,.) x_c in a fixed
        // order (forward: the entries of row i of L, ascending c; backward: the entries of column i, ascending row) and an
        // entry may be applied as soon as ITS x_c is final -- a row need not wait for its last column as in a level set.
        // Greedy step schedule: in every step each unfinished chain whose next entry is ready takes up to B consecutive ready
        // entries (at most 32 chains per step, one per lane); a step is one "level" of the stream, a chain appears in as many
        // levels as it takes steps (its partial sum lives in the unknown's own slot).  C4: the forward sweep takes 367 steps
        // of ~27 entries instead of 345 levels of ~3 rows x 13 entries, the backward sweep 705 steps (B = 4) instead of a
        // critical path of 2 647 single entries.  B is chosen per sweep by a cost model of the device loop.
        std::vector<std::vector<std::pair<int, int>>> fent(N), bent(N);
        for (int i = 0; i < N; i++) for (int sl = S.LrP[i]; sl < S.LrP[i + 1]; sl++) fent[i].push_back({sl, S.LrC[sl]});
        for (int i = 0; i < N; i++) for (int p = S.Lp[i]; p < S.Lp[i + 1]; p++) bent[i].push_back({p, S.Li[p]});
        auto schedule = [&](const std::vector<std::vector<std::pair<int, int>>>& ent, int B, bool descending, std::vector<std::vector<StreamRow>>& out) {
            std::vector<int> pos(N, 0), finstep(N, -1);
            std::vector<char> done(N, 0);
            int remaining = 0;
            for (int i = 0; i < N; i++) { if (ent[i].empty()) done[i] = 1; else remaining++; }
            out.clear();
            double cost = 0.0;
            for (int step = 0; remaining > 0; step++) {
                std::vector<StreamRow> lv;
                int Emax = 0;
                for (int ii = 0; ii < N && (int)lv.size() < 32; ii++) {
                    const int i = descending ? N - 1 - ii : ii;
                    if (done[i]) continue;
                    int c = 0;
                    while (c < B && pos[i] + c < (int)ent[i].size()) {
                        const int d = ent[i][pos[i] + c].second;
                        if (done[d] && finstep[d] < step) c++; else break;
                    }
                    if (c == 0) continue;
                    StreamRow r; r.id = i;
                    r.ent.assign(ent[i].begin() + pos[i], ent[i].begin() + pos[i] + c);
                    lv.push_back(std::move(r));
                    Emax = std::max(Emax, c);
                }
                if (lv.empty()) { out.clear(); return -1.0; }   // (cannot happen: L is triangular)
                for (const auto& r : lv) {
                    pos[r.id] += (int)r.ent.size();
                    if (pos[r.id] == (int)ent[r.id].size()) { done[r.id] = 1; finstep[r.id] = step; remaining--; }
                }
                cost += 110.0 + 45.0 * Emax;   // cycles of a level in the device loop (header + row, then the entries)
                out.push_back(std::move(lv));
            }
            return cost;
        };
        auto best = [&](const std::vector<std::vector<std::pair<int, int>>>& ent, bool descending, std::vector<std::vector<StreamRow>>& out) {
            double bc = -1.0;
            for (int B : {1, 2, 4}) {
                std::vector<std::vector<StreamRow>> cand;
                const double c = schedule(ent, B, descending, cand);
                if (c >= 0.0 && (bc < 0.0 || c < bc)) { bc = c; out = std::move(cand); }
            }
        };
        std::vector<std::vector<StreamRow>> fw, bw;
        if (N + 3 <= 65535) {   // (16-bit index words; beyond that the sweeps read L in place and no schedule is needed)
            best(fent, false, fw);
            best(bent, true, bw);
        }
        const bool okf = !fw.empty() && !bw.empty() && build_stream(N, fw, S.fsI, S.fsSrc, S.fsChunks);
        const bool okb = okf && build_stream(N, bw, S.bsI, S.bsSrc, S.bsChunks);
        S.stream = (okf && okb) ? 1 : 0;
        if (!S.stream) { S.fsI.clear(); S.bsI.clear(); S.fsSrc.clear(); S.bsSrc.clear(); S.fsChunks = S.bsChunks = 0; }
    }
    {
        S.sOff.assign(S.rcol.size() + 1, 0);
        S.fpIdx.clear(); S.fpLi.clear(); S.fpStep.clear();
        S.rowPair.assign(N + 1, 0);
        for (int k = 0; k < N; k++) {
            S.rowPair[k] = (int)S.fpIdx.size();
            for (int t = S.rp[k]; t < S.rp[k + 1]; t++) {
                S.sOff[t] = (int)S.fpIdx.size();
                const int c = S.rcol[t];
                for (int p = S.Lp[c]; p < S.rpos[t]; p++) { S.fpIdx.push_back(p); S.fpLi.push_back(S.Li[p]); S.fpStep.push_back(t - S.rp[k]); }
            }
        }
        S.sOff[S.rcol.size()] = (int)S.fpIdx.size();
        S.rowPair[N] = (int)S.fpIdx.size();
    }
    S.Lcol.assign(S.Li.size(), 0);
    S.Lrev.clear();
    for (int c = 0; c < N; c++) for (int p = S.Lp[c]; p < S.Lp[c + 1]; p++) S.Lcol[p] = c;
    for (int c = N - 1; c >= 0; c--) for (int p = S.Lp[c]; p < S.Lp[c + 1]; p++) S.Lrev.push_back(p);
}

// Minimum degree gives the least fill; when its elimination tree is tall (many levels = many warp-wide synchronisations
// per triangular solve) nested dissection is tried too and taken if it at least halves the levels at no more than three
// times the fill.
inline void analyse(int n, int m, const std::vector<Trip>& Qpat, const std::vector<Trip>& Apat, Symbolic& S)
{
    analyse_with(n, m, Qpat, Apat, S, 0);
    if (S.flP.size() - 1 > 64) {
        Symbolic T;
        analyse_with(n, m, Qpat, Apat, T, 1);
        if (2 * (T.flP.size() - 1) <= (S.flP.size() - 1) && T.Li.size() <= 3 * S.Li.size()) S = std::move(T);
    }
}

// Patterns from dense row-major matrices (the batch's union of non-zeros is passed in as 0/1 masks)
inline void dense_patterns(int n, int nC, int nComp, const unsigned char* Qmask, const unsigned char* Amask, const unsigned char* Lmask,
                           const unsigned char* Rmask, std::vector<Trip>& Qpat, std::vector<Trip>& Apat)
{
    Qpat.clear(); Apat.clear();
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) if (Qmask[(size_t)i * n + j] || Qmask[(size_t)j * n + i]) Qpat.push_back({i, j, i * n + j});
    for (int r = 0; r < nC; r++) for (int j = 0; j < n; j++) if (Amask[(size_t)r * n + j]) Apat.push_back({r, j, (0 << 28) | (r * n + j)});
    for (int r = 0; r < nComp; r++) for (int j = 0; j < n; j++) if (Lmask[(size_t)r * n + j]) Apat.push_back({nC + r, j, (1 << 28) | (r * n + j)});
    for (int r = 0; r < nComp; r++) for (int j = 0; j < n; j++) if (Rmask[(size_t)r * n + j]) Apat.push_back({nC + nComp + r, j, (2 << 28) | (r * n + j)});
}

// Patterns from CSC index arrays (LCQProblem::loadLCQP(const csc*...), /root/reference/src/LCQProblem.cpp:312-387)
inline void csc_patterns(int n, int nC, int nComp, const int* Qp, const int* Qi, const int* Ap, const int* Ai, const int* Lp, const int* Li,
                         const int* Rp, const int* Ri, std::vector<Trip>& Qpat, std::vector<Trip>& Apat)
{
    Qpat.clear(); Apat.clear();
    for (int c = 0; c < n; c++) for (int p = Qp[c]; p < Qp[c + 1]; p++) Qpat.push_back({Qi[p], c, p});
    if (Ap) for (int c = 0; c < n; c++) for (int p = Ap[c]; p < Ap[c + 1]; p++) Apat.push_back({Ai[p], c, (0 << 28) | p});
    for (int c = 0; c < n; c++) for (int p = Lp[c]; p < Lp[c + 1]; p++) Apat.push_back({nC + Li[p], c, (1 << 28) | p});
    for (int c = 0; c < n; c++) for (int p = Rp[c]; p < Rp[c + 1]; p++) Apat.push_back({nC + nComp + Ri[p], c, (2 << 28) | p});
}

}  // namespace osq
}  // namespace lcqp
