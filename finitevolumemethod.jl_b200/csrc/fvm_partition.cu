// libfvmcuda: METIS-style graph partitioning of the mesh's node graph for the multi-GPU path (SURVEY.md 8e,
// north_star d).  No METIS header exists offline, so the scheme is written out here: multilevel recursive
// bisection.
//   coarsening    heavy-edge matching (a node is merged with the unmatched neighbour it shares the heaviest edge
//                 with) until a few thousand weighted nodes are left;
//   initial cut   greedy graph growing on the coarsest graph from several seeds (a pseudo-peripheral node, its
//                 antipode, the node farthest from both, and pseudo-random ones), the smallest refined cut wins;
//   uncoarsening  the cut is projected level by level and improved by Fiduccia-Mattheyses-style boundary
//                 refinement (boundary nodes with positive gain change sides within a balance tolerance);
//   k parts       recursive bisection with proportional targets (k need not be a power of two); on the finest
//                 level node counts are restored exactly, so parts differ by at most one node.
// The node graph is the one `jacobian_sparsity` walks (/root/reference/src/solve.jl:56-77): an edge for every
// pair of nodes sharing a triangle.  Deterministic.  Host-only code: needs no CUDA device.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <queue>

#include "fvm_internal.h"

namespace {

// FVM_PART_VERBOSE=1: wall time of every phase on stderr
struct PhaseTimer {
    const char* name;
    int32_t n;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    PhaseTimer(const char* nm, int32_t nn) : name(nm), n(nn) {}
    ~PhaseTimer() {
        static const bool on = getenv("FVM_PART_VERBOSE") != nullptr;
        if (on) fprintf(stderr, "[partition] %-10s n = %9d  %.3f s\n", name, n, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
};

struct WGraph {
    int32_t n = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> adj, ew, vw;  // neighbours, edge weights, node weights
    int64_t total_w = 0;
};

WGraph graph_from_triangles(int64_t N, const int32_t* tri, int64_t T, int32_t base) {
    PhaseTimer pt("graph", (int32_t)N);
    WGraph g;
    g.n = (int32_t)N;
    std::vector<int64_t> cnt(N + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t)
        for (int r = 0; r < 3; ++r) {
#pragma omp atomic
            cnt[tri[3 * t + r] - base + 1] += 2;
        }
    for (int64_t i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
    std::vector<int32_t> raw(cnt[N]);
    std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        const int32_t v[3] = {tri[3 * t] - base, tri[3 * t + 1] - base, tri[3 * t + 2] - base};
        for (int r = 0; r < 3; ++r) {
            int64_t at;
#pragma omp atomic capture
            {
                at = fill[v[r]];
                fill[v[r]] += 2;
            }
            raw[at] = v[(r + 1) % 3];  // the order inside a row depends on the schedule; the rows are sorted below
            raw[at + 1] = v[(r + 2) % 3];
        }
    }
    // unique neighbours per row: sort + unique in place, then compact
    std::vector<int64_t> len(N + 1, 0);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < N; ++i) {
        std::sort(raw.begin() + cnt[i], raw.begin() + cnt[i + 1]);
        len[i + 1] = std::unique(raw.begin() + cnt[i], raw.begin() + cnt[i + 1]) - (raw.begin() + cnt[i]);
    }
    for (int64_t i = 0; i < N; ++i) len[i + 1] += len[i];
    g.ptr = len;
    g.adj.resize(len[N]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) std::copy(raw.begin() + cnt[i], raw.begin() + cnt[i] + (len[i + 1] - len[i]), g.adj.begin() + len[i]);
    g.ew.assign(g.adj.size(), 1);
    g.vw.assign(N, 1);
    g.total_w = N;
    return g;
}

int64_t cut_of(const WGraph& g, const std::vector<uint8_t>& side) {
    int64_t c = 0;
    for (int32_t v = 0; v < g.n; ++v)
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
            if (side[v] != side[g.adj[e]]) c += g.ew[e];
    return c / 2;
}

// ---- coarsening: heavy-edge matching --------------------------------------------------------------------
// Large graphs: the nodes are cut into fixed chunks of consecutive ids (independent of the thread count, so the result
// is too); every chunk is matched on its own (heavy edges to nodes of the same chunk only, ascending ids) and the coarse
// rows are built chunk by chunk with a linear search instead of a node-indexed position table.  Mesh numberings have locality, so few heavy edges cross a chunk border; if the
// numbering has none the graph barely shrinks, `ok` is false and the caller falls back to the global serial matching.
WGraph coarsen_chunked(const WGraph& g, std::vector<int32_t>& cmap, int64_t max_vw, bool& ok) {
    PhaseTimer pt("coarsen/omp", g.n);
    const int32_t n = g.n;
    const int32_t CH = std::max<int32_t>(1 << 15, (n + 63) / 64);  // at most 64 chunks: a 67M-node lattice loses 1 row in 128
    const int32_t nch = (n + CH - 1) / CH;
    std::vector<int32_t> match(n, -1);
    cmap.assign(n, -1);
    std::vector<int32_t> cbase(nch + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t c = 0; c < nch; ++c) {
        const int32_t lo = c * CH, hi = std::min<int64_t>(n, (int64_t)lo + CH);
        int32_t nc = 0;
        // ascending ids: sequential memory traffic.  On a row-major lattice this pairs along the rows at one level and
        // across them at the next (the merged pairs share the heavier edges), i.e. it halves the graph every level
        for (int32_t v = lo; v < hi; ++v) {
            if (match[v] >= 0) continue;
            int32_t best = -1, best_w = -1;
            for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
                const int32_t u = g.adj[e];
                if (u >= lo && u < hi && match[u] < 0 && u != v && g.ew[e] > best_w && (int64_t)g.vw[v] + g.vw[u] <= max_vw) {
                    best_w = g.ew[e];
                    best = u;
                }
            }
            match[v] = best >= 0 ? best : v;
            if (best >= 0) match[best] = v;
            cmap[v] = nc;
            if (best >= 0) cmap[best] = nc;
            ++nc;
        }
        cbase[c + 1] = nc;
    }
    for (int32_t c = 0; c < nch; ++c) cbase[c + 1] += cbase[c];
    const int32_t nc = cbase[nch];
    WGraph c;
    ok = nc < n - n / 5;
    if (!ok) return c;
    std::vector<int32_t> first(nc), second(nc, -1);
#pragma omp parallel for schedule(static)
    for (int32_t v = 0; v < n; ++v) {
        cmap[v] += cbase[v / CH];
        if (match[v] == v) first[cmap[v]] = v;
        else if (v < match[v]) {
            first[cmap[v]] = v;
            second[cmap[v]] = match[v];
        }
    }
    c.n = nc;
    c.vw.resize(nc);
    c.ptr.assign(nc + 1, 0);
    c.total_w = g.total_w;
    const int32_t RCH = 1 << 15;
    const int32_t nrch = (nc + RCH - 1) / RCH;
    std::vector<std::vector<int32_t>> ladj(nrch), lew(nrch);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t rc = 0; rc < nrch; ++rc) {
        const int32_t lo = rc * RCH, hi = std::min<int64_t>(nc, (int64_t)lo + RCH);
        std::vector<int32_t>& A = ladj[rc];
        std::vector<int32_t>& W = lew[rc];
        A.reserve((size_t)(hi - lo) * 7);
        W.reserve((size_t)(hi - lo) * 7);
        for (int32_t cv = lo; cv < hi; ++cv) {
            const size_t row0 = A.size();
            int32_t w = 0;
            for (int32_t v : {first[cv], second[cv]}) {
                if (v < 0) continue;
                w += g.vw[v];
                for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
                    const int32_t cu = cmap[g.adj[e]];
                    if (cu == cv) continue;
                    size_t k = row0;
                    while (k < A.size() && A[k] != cu) ++k;
                    if (k < A.size()) W[k] += g.ew[e];
                    else {
                        A.push_back(cu);
                        W.push_back(g.ew[e]);
                    }
                }
            }
            c.vw[cv] = w;
            c.ptr[cv + 1] = (int64_t)(A.size() - row0);
        }
    }
    for (int32_t cv = 0; cv < nc; ++cv) c.ptr[cv + 1] += c.ptr[cv];
    c.adj.resize(c.ptr[nc]);
    c.ew.resize(c.ptr[nc]);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t rc = 0; rc < nrch; ++rc) {
        const int64_t at = c.ptr[(int64_t)rc * RCH];
        std::copy(ladj[rc].begin(), ladj[rc].end(), c.adj.begin() + at);
        std::copy(lew[rc].begin(), lew[rc].end(), c.ew.begin() + at);
    }
    return c;
}

WGraph coarsen(const WGraph& g, std::vector<int32_t>& cmap, int64_t max_vw) {
    if (g.n >= (1 << 17)) {
        bool ok = false;
        WGraph c = coarsen_chunked(g, cmap, max_vw, ok);
        if (ok) return c;
    }
    PhaseTimer pt("coarsen", g.n);
    const int32_t n = g.n;
    std::vector<int32_t> match(n, -1);
    cmap.assign(n, -1);
    int32_t nc = 0;
    // visit in a fixed pseudo-random order (a stride coprime to n) so that the matching does not follow the numbering
    int64_t stride = (int64_t)(0.6180339887 * n) | 1;
    while (std::gcd<int64_t, int64_t>(stride, n) != 1) stride += 2;
    int64_t v64 = 0;
    for (int32_t it = 0; it < n; ++it, v64 = (v64 + stride) % n) {
        const int32_t v = (int32_t)v64;
        if (match[v] >= 0) continue;
        int32_t best = -1, best_w = -1;
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
            const int32_t u = g.adj[e];
            if (match[u] < 0 && u != v && g.ew[e] > best_w && (int64_t)g.vw[v] + g.vw[u] <= max_vw) {
                best_w = g.ew[e];
                best = u;
            }
        }
        match[v] = best >= 0 ? best : v;
        if (best >= 0) match[best] = v;
        cmap[v] = nc;
        if (best >= 0) cmap[best] = nc;
        ++nc;
    }
    WGraph c;
    c.n = nc;
    c.vw.assign(nc, 0);
    c.ptr.assign(nc + 1, 0);
    c.total_w = g.total_w;
    std::vector<int32_t> first(nc, -1), second(nc, -1);
    for (int32_t v = 0; v < n; ++v) {
        c.vw[cmap[v]] += g.vw[v];
        (first[cmap[v]] < 0 ? first : second)[cmap[v]] = v;
    }
    std::vector<int64_t> pos(nc, -1);  // position of a coarse neighbour inside the row being built
    c.adj.reserve(g.adj.size() / 2);
    c.ew.reserve(g.adj.size() / 2);
    for (int32_t cv = 0; cv < nc; ++cv) {
        const int64_t row0 = (int64_t)c.adj.size();
        for (int32_t v : {first[cv], second[cv]}) {
            if (v < 0) continue;
            for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
                const int32_t cu = cmap[g.adj[e]];
                if (cu == cv) continue;
                if (pos[cu] >= row0) {
                    c.ew[pos[cu]] += g.ew[e];
                } else {
                    pos[cu] = (int64_t)c.adj.size();
                    c.adj.push_back(cu);
                    c.ew.push_back(g.ew[e]);
                }
            }
        }
        c.ptr[cv + 1] = (int64_t)c.adj.size();
    }
    return c;
}

// ---- Fiduccia-Mattheyses-style boundary refinement --------------------------------------------------------
// side 0 should weigh target0.  `exact`: finish with |w0 - target0| minimal (node counts on the finest level).
void refine(const WGraph& g, int64_t target0, std::vector<uint8_t>& side, bool exact) {
    PhaseTimer pt("refine", g.n);
    const int32_t n = g.n;
    int32_t max_vw = 1;
    for (int32_t v = 0; v < n; ++v) max_vw = std::max(max_vw, g.vw[v]);
    const int64_t tol = std::max<int64_t>(max_vw, g.total_w / 200);  // 0.5 % while the cut is being improved
    int64_t w0 = 0;
    for (int32_t v = 0; v < n; ++v)
        if (!side[v]) w0 += g.vw[v];
    auto gain = [&](int32_t v) {
        int64_t ext = 0, in = 0;
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) (side[g.adj[e]] != side[v] ? ext : in) += g.ew[e];
        return ext - in;
    };
    // FM passes: always move the unlocked boundary node with the largest gain (negative gains included, so the
    // search can climb out of a local minimum and straighten a wiggly cut), remember the best prefix of the move
    // sequence, undo the rest.  A pass ends after `patience` moves without a new best.
    std::vector<uint8_t> locked(n);
    std::vector<std::pair<int64_t, int32_t>> cand;
    std::vector<int32_t> moves;
    std::priority_queue<std::pair<int64_t, int32_t>> heap;
    const int64_t patience = std::max<int64_t>(64, std::min<int64_t>(2000, n / 50));
    for (int pass = 0; pass < 8; ++pass) {
        std::fill(locked.begin(), locked.end(), 0);
        while (!heap.empty()) heap.pop();
        for (int32_t v = 0; v < n; ++v) {
            bool boundary = false;
            for (int64_t e = g.ptr[v]; e < g.ptr[v + 1] && !boundary; ++e) boundary = side[g.adj[e]] != side[v];
            if (boundary) heap.emplace(gain(v), v);
        }
        moves.clear();
        int64_t delta = 0, best_delta = 0, best_imb = std::llabs(w0 - target0);
        size_t best_len = 0;
        while (!heap.empty() && (int64_t)(moves.size() - best_len) < patience) {
            const auto top = heap.top();
            heap.pop();
            const int32_t v = top.second;
            if (locked[v]) continue;
            const int64_t gv = gain(v);
            if (gv != top.first) {  // stale entry: requeue with the current gain
                heap.emplace(gv, v);
                continue;
            }
            const int64_t new_w0 = w0 + (side[v] ? g.vw[v] : -g.vw[v]);
            if (std::llabs(new_w0 - target0) > tol && std::llabs(new_w0 - target0) >= std::llabs(w0 - target0)) continue;
            side[v] ^= 1;
            w0 = new_w0;
            locked[v] = 1;
            delta -= gv;
            moves.push_back(v);
            const int64_t imb = std::llabs(w0 - target0);
            if (delta < best_delta || (delta == best_delta && imb < best_imb)) {
                best_delta = delta;
                best_imb = imb;
                best_len = moves.size();
            }
            for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
                if (!locked[g.adj[e]]) heap.emplace(gain(g.adj[e]), g.adj[e]);
        }
        for (size_t q = moves.size(); q > best_len; --q) {  // undo the tail after the best prefix
            const int32_t v = moves[q - 1];
            side[v] ^= 1;
            w0 += side[v] ? -g.vw[v] : g.vw[v];
        }
        if (best_len == 0) break;
    }
    // balance: move the boundary nodes that cost the least from the heavy side
    const int64_t slack = exact ? 0 : max_vw;
    for (int guard = 0; guard < 64 && std::llabs(w0 - target0) > slack; ++guard) {
        const uint8_t from = w0 > target0 ? 0 : 1;
        cand.clear();
        for (int32_t v = 0; v < n; ++v) {
            if (side[v] != from) continue;
            bool boundary = g.ptr[v + 1] == g.ptr[v];  // isolated points can go anywhere
            for (int64_t e = g.ptr[v]; e < g.ptr[v + 1] && !boundary; ++e) boundary = side[g.adj[e]] != from;
            if (boundary) cand.emplace_back(-gain(v), v);
        }
        if (cand.empty())  // nothing touches the other side (a disconnected remainder): any node will do
            for (int32_t v = 0; v < n; ++v)
                if (side[v] == from) cand.emplace_back(0, v);
        std::sort(cand.begin(), cand.end());
        for (auto& c : cand) {
            const int64_t diff = std::llabs(w0 - target0);
            if (diff <= slack) break;
            const int32_t v = c.second;
            if (g.vw[v] > 2 * diff) continue;  // would overshoot by more than it repairs
            side[v] ^= 1;
            w0 += from == 0 ? -g.vw[v] : g.vw[v];
        }
    }
}

// ---- initial bisection of the coarsest graph: greedy graph growing from several seeds ----------------------
void bfs_order(const WGraph& g, const std::vector<uint8_t>& taken, int32_t src, std::vector<int32_t>& order, std::vector<int32_t>& mark,
               int32_t stamp, std::vector<int32_t>* dist = nullptr) {
    order.clear();
    order.push_back(src);
    mark[src] = stamp;
    if (dist) (*dist)[src] = 0;
    for (size_t head = 0; head < order.size(); ++head) {
        const int32_t v = order[head];
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
            const int32_t w = g.adj[e];
            if (!taken[w] && mark[w] != stamp) {
                mark[w] = stamp;
                if (dist) (*dist)[w] = (*dist)[v] + 1;
                order.push_back(w);
            }
        }
    }
}

void initial_bisect(const WGraph& g, int64_t target0, std::vector<uint8_t>& side) {
    const int32_t n = g.n;
    std::vector<int32_t> mark(n, 0), order, dp(n, 0), dq(n, 0);
    std::vector<uint8_t> taken(n, 0);
    int32_t stamp = 0;
    auto peripheral = [&](int32_t src) {
        for (int sweep = 0; sweep < 2; ++sweep) {
            bfs_order(g, taken, src, order, mark, ++stamp);
            src = order.back();
        }
        return src;
    };
    std::vector<int32_t> seeds;
    {
        const int32_t p = peripheral(0);
        bfs_order(g, taken, p, order, mark, ++stamp, &dp);
        const int32_t q = order.back();
        const std::vector<int32_t> comp = order;
        bfs_order(g, taken, q, order, mark, ++stamp, &dq);
        int32_t third = p, best_d = -1;
        for (int32_t v : comp) {
            const int32_t d = std::min(dp[v], dq[v]);
            if (d > best_d) {
                best_d = d;
                third = v;
            }
        }
        seeds = {p, q, third};
        uint64_t s = 0x9E3779B97F4A7C15ull;  // a few fixed pseudo-random seeds as well
        for (int k = 0; k < 5; ++k) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            seeds.push_back((int32_t)((s >> 33) % (uint64_t)n));
        }
    }
    std::vector<uint8_t> trial(n);
    int64_t best_cut = INT64_MAX;
    for (size_t t = 0; t < seeds.size(); ++t) {
        std::fill(taken.begin(), taken.end(), 0);
        std::fill(trial.begin(), trial.end(), 1);
        int64_t w0 = 0;
        int32_t scan = 0;
        bool first = true;
        while (w0 < target0) {
            int32_t src;
            if (first) {
                src = seeds[t];
            } else {  // the seed's component is exhausted: continue in the next untouched one
                while (scan < n && taken[scan]) ++scan;
                if (scan == n) break;
                src = peripheral(scan);
            }
            first = false;
            bfs_order(g, taken, src, order, mark, ++stamp);
            for (int32_t v : order) {
                if (w0 >= target0) break;
                trial[v] = 0;
                w0 += g.vw[v];
            }
            for (int32_t v : order) taken[v] = 1;  // the rest of this component stays on side 1
        }
        refine(g, target0, trial, false);
        const int64_t c = cut_of(g, trial);
        if (c < best_cut) {
            best_cut = c;
            side = trial;
        }
    }
}

void multilevel_bisect(const WGraph& g, int64_t target0, std::vector<uint8_t>& side, bool finest) {
    const int32_t coarse_enough = 4000;
    if (g.n > coarse_enough) {
        std::vector<int32_t> cmap;
        const WGraph c = coarsen(g, cmap, std::max<int64_t>(2, g.total_w / (coarse_enough / 4)));
        if (c.n < g.n - g.n / 10) {  // the matching still shrinks the graph
            std::vector<uint8_t> cside;
            multilevel_bisect(c, target0, cside, false);
            side.resize(g.n);
#pragma omp parallel for schedule(static)
            for (int32_t v = 0; v < g.n; ++v) side[v] = cside[cmap[v]];
            refine(g, target0, side, finest);
            return;
        }
    }
    initial_bisect(g, target0, side);
    if (finest) refine(g, target0, side, true);
}

// induced subgraph of the nodes with side == s
WGraph subgraph(const WGraph& g, const std::vector<uint8_t>& side, uint8_t s, const std::vector<int32_t>& ids, std::vector<int32_t>& ids_out) {
    PhaseTimer pt("subgraph", g.n);
    const int32_t n = g.n;
    const int32_t CH = 1 << 16, nch = (n + CH - 1) / CH;
    std::vector<int32_t> local(n, -1), base(nch + 1, 0);
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < nch; ++c) {
        int32_t k = 0;
        for (int32_t v = c * CH; v < std::min<int64_t>(n, (int64_t)(c + 1) * CH); ++v) k += side[v] == s;
        base[c + 1] = k;
    }
    for (int32_t c = 0; c < nch; ++c) base[c + 1] += base[c];
    WGraph h;
    h.n = base[nch];
    ids_out.resize(h.n);
    h.ptr.assign((size_t)h.n + 1, 0);
    h.vw.resize(h.n);
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < nch; ++c) {
        int32_t k = base[c];
        for (int32_t v = c * CH; v < std::min<int64_t>(n, (int64_t)(c + 1) * CH); ++v)
            if (side[v] == s) {
                local[v] = k;
                ids_out[k] = ids[v];
                h.vw[k] = g.vw[v];
                ++k;
            }
    }
    int64_t tw = 0;
#pragma omp parallel for schedule(static) reduction(+ : tw)
    for (int32_t v = 0; v < n; ++v) {
        if (local[v] < 0) continue;
        tw += g.vw[v];
        int64_t d = 0;
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) d += local[g.adj[e]] >= 0;
        h.ptr[local[v] + 1] = d;
    }
    h.total_w = tw;
    for (int32_t k = 0; k < h.n; ++k) h.ptr[k + 1] += h.ptr[k];
    h.adj.resize(h.ptr[h.n]);
    h.ew.resize(h.ptr[h.n]);
#pragma omp parallel for schedule(static)
    for (int32_t v = 0; v < n; ++v) {
        if (local[v] < 0) continue;
        int64_t at = h.ptr[local[v]];
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
            if (local[g.adj[e]] >= 0) {
                h.adj[at] = local[g.adj[e]];
                h.ew[at] = g.ew[e];
                ++at;
            }
    }
    return h;
}

void partition_rec(const WGraph& g, const std::vector<int32_t>& ids, int32_t k, int32_t part0, int32_t* owner) {
    if (k == 1) {
        for (int32_t v = 0; v < g.n; ++v) owner[ids[v]] = part0;
        return;
    }
    const int32_t k_left = k / 2;
    const int64_t target0 = g.total_w * k_left / k;
    std::vector<uint8_t> side;
    multilevel_bisect(g, target0, side, true);
    for (uint8_t s = 0; s < 2; ++s) {
        std::vector<int32_t> sub_ids;
        const WGraph h = subgraph(g, side, s, ids, sub_ids);
        partition_rec(h, sub_ids, s == 0 ? k_left : k - k_left, s == 0 ? part0 : part0 + k_left, owner);
    }
}

}  // namespace

extern "C" int32_t fvm_partition_graph(int64_t n_points, const int32_t* triangles, int64_t n_triangles, int32_t index_base,
                                       int32_t n_parts, int32_t* owner) {
    if (!triangles || !owner || n_points <= 0 || n_triangles <= 0 || n_parts < 1 || n_parts > n_points)
        return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_graph: bad arguments");
    if (n_points >= INT32_MAX) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_graph: mesh too large for int32 indices");
    for (int64_t i = 0; i < 3 * n_triangles; ++i)
        if (triangles[i] - index_base < 0 || triangles[i] - index_base >= n_points)
            return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_graph: triangle vertex out of range");
    const WGraph g = graph_from_triangles(n_points, triangles, n_triangles, index_base);
    std::vector<int32_t> ids(n_points);
    std::iota(ids.begin(), ids.end(), 0);
    partition_rec(g, ids, n_parts, 0, owner);
    return FVM_OK;
}

// number of node-graph edges whose endpoints have different owners (the halo volume is proportional to it)
extern "C" int32_t fvm_partition_edge_cut(int64_t n_points, const int32_t* triangles, int64_t n_triangles, int32_t index_base,
                                          const int32_t* owner, int64_t* cut) {
    if (!triangles || !owner || !cut || n_points <= 0 || n_triangles <= 0) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_edge_cut: bad arguments");
    const WGraph g = graph_from_triangles(n_points, triangles, n_triangles, index_base);
    int64_t c = 0;
    for (int64_t v = 0; v < n_points; ++v)
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) c += owner[v] != owner[g.adj[e]];
    *cut = c / 2;
    return FVM_OK;
}
