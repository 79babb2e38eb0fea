// libfvmcuda: METIS-style graph partitioning of the mesh's node graph for the multi-GPU path (SURVEY.md 8e,
// north_star d).  No METIS header exists offline, so this is the classical recursive-bisection scheme METIS's
// initial partitioner is built on: greedy graph growing (a BFS level structure from a pseudo-peripheral node
// fills the first half) followed by Fiduccia-Mattheyses-style boundary refinement (boundary nodes with
// positive gain change sides while the balance stays within tolerance), recursively for k parts with
// proportional targets, so k need not be a power of two.  The node graph is the one `jacobian_sparsity`
// walks (/root/reference/src/solve.jl:56-77): an edge for every pair of nodes sharing a triangle.
// Host-only code: needs no CUDA device.
#include <algorithm>
#include <numeric>
#include <queue>

#include "fvm_internal.h"

namespace {

struct Graph {
    int64_t n = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> adj;
};

Graph build_graph(int64_t N, const int32_t* tri, int64_t T, int32_t base) {
    Graph g;
    g.n = N;
    std::vector<int64_t> cnt(N + 1, 0);
    for (int64_t t = 0; t < T; ++t)
        for (int r = 0; r < 3; ++r) cnt[tri[3 * t + r] - base + 1] += 2;
    for (int64_t i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
    std::vector<int32_t> raw(cnt[N]);
    std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
    for (int64_t t = 0; t < T; ++t) {
        const int32_t v[3] = {tri[3 * t] - base, tri[3 * t + 1] - base, tri[3 * t + 2] - base};
        for (int r = 0; r < 3; ++r) {
            raw[fill[v[r]]++] = v[(r + 1) % 3];
            raw[fill[v[r]]++] = v[(r + 2) % 3];
        }
    }
    g.ptr.assign(N + 1, 0);
    g.adj.reserve(raw.size() / 2 + N);
    for (int64_t i = 0; i < N; ++i) {  // unique neighbours
        std::sort(raw.begin() + cnt[i], raw.begin() + cnt[i + 1]);
        auto e = std::unique(raw.begin() + cnt[i], raw.begin() + cnt[i + 1]);
        g.adj.insert(g.adj.end(), raw.begin() + cnt[i], e);
        g.ptr[i + 1] = (int64_t)g.adj.size();
    }
    return g;
}

// BFS over the nodes of `part` (label[v] == lab) from `src`; returns the visiting order (only the component of src)
void bfs(const Graph& g, const std::vector<int32_t>& label, int32_t lab, int32_t src, std::vector<int32_t>& order,
         std::vector<int32_t>& mark, int32_t stamp, std::vector<int32_t>* dist = nullptr) {
    order.clear();
    order.push_back(src);
    mark[src] = stamp;
    if (dist) (*dist)[src] = 0;
    for (size_t head = 0; head < order.size(); ++head) {
        const int32_t v = order[head];
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
            const int32_t w = g.adj[e];
            if (label[w] == lab && mark[w] != stamp) {
                mark[w] = stamp;
                if (dist) (*dist)[w] = (*dist)[v] + 1;
                order.push_back(w);
            }
        }
    }
}

struct Bisector {
    const Graph& g;
    std::vector<int32_t>& label;  // current part of every node
    std::vector<int32_t> mark, order, dp, dq;
    int32_t stamp = 0;
    int32_t next_label;

    Bisector(const Graph& g_, std::vector<int32_t>& l, int32_t first_free) : g(g_), label(l), mark(g_.n, 0), next_label(first_free) {}

    // grows `lab` from `seed` over the unassigned nodes (label == other) until n_left nodes are taken; further
    // components are entered from their own pseudo-peripheral nodes
    void grow(const std::vector<int32_t>& nodes, int32_t lab, int32_t other, int64_t n_left, int32_t seed) {
        for (int32_t v : nodes) label[v] = other;
        int64_t grown = 0;
        size_t scan = 0;
        bool first = true;
        while (grown < n_left) {
            int32_t src;
            if (first && seed >= 0) {
                src = seed;
            } else {
                while (scan < nodes.size() && label[nodes[scan]] != other) ++scan;  // next unassigned component
                src = nodes[scan];
                for (int sweep = 0; sweep < 2; ++sweep) {
                    bfs(g, label, other, src, order, mark, ++stamp);
                    src = order.back();
                }
            }
            first = false;
            bfs(g, label, other, src, order, mark, ++stamp);
            for (int32_t v : order) {
                if (grown == n_left) break;
                label[v] = lab;
                ++grown;
            }
        }
    }

    int64_t cut_between(const std::vector<int32_t>& nodes, int32_t a, int32_t b) const {
        int64_t c = 0;
        for (int32_t v : nodes)
            if (label[v] == a)
                for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) c += label[g.adj[e]] == b;
        return c;
    }

    // splits the nodes `nodes` (all labelled `lab`) into `lab` (n_left nodes) and a new label; returns the new label.
    // Like METIS's initial partitioner, several growing seeds are tried and the smallest refined cut is kept: a
    // pseudo-peripheral node p, its antipode q, and the node farthest from both (a "third corner": its front
    // runs across the one grown from p).
    int32_t bisect(std::vector<int32_t>& nodes, int32_t lab, int64_t n_left) {
        const int32_t other = next_label++;
        const int64_t n = (int64_t)nodes.size();
        for (int32_t v : nodes) label[v] = other;
        int32_t seeds[3] = {-1, -1, -1};
        {
            int32_t src = nodes[0];
            for (int sweep = 0; sweep < 2; ++sweep) {
                bfs(g, label, other, src, order, mark, ++stamp);
                src = order.back();
            }
            seeds[0] = src;
            if (dp.size() < (size_t)g.n) {
                dp.assign(g.n, 0);
                dq.assign(g.n, 0);
            }
            bfs(g, label, other, seeds[0], order, mark, ++stamp, &dp);  // graph distances from p
            seeds[1] = order.back();
            const std::vector<int32_t> comp = order;                      // the component of p
            bfs(g, label, other, seeds[1], order, mark, ++stamp, &dq);  // ... and from q
            int32_t best_d = -1;
            for (int32_t v : comp) {
                const int32_t d = std::min(dp[v], dq[v]);
                if (d > best_d) {
                    best_d = d;
                    seeds[2] = v;
                }
            }
        }
        std::vector<int32_t> best_label;
        int64_t best_cut = INT64_MAX;
        for (int trial = 0; trial < 3; ++trial) {
            if (seeds[trial] < 0 || (trial > 0 && seeds[trial] == seeds[trial - 1])) continue;
            grow(nodes, lab, other, n_left, seeds[trial]);
            refine(nodes, lab, other, n_left, n);
            const int64_t c = cut_between(nodes, lab, other);
            if (c < best_cut) {
                best_cut = c;
                best_label.resize(nodes.size());
                for (size_t i = 0; i < nodes.size(); ++i) best_label[i] = label[nodes[i]];
            }
        }
        for (size_t i = 0; i < nodes.size(); ++i) label[nodes[i]] = best_label[i];
        return other;
    }

    // Fiduccia-Mattheyses flavoured boundary refinement: alternate sides, always move the boundary node with the
    // largest positive gain (external - internal degree) that keeps |left| within the tolerance of its target
    void refine(const std::vector<int32_t>& nodes, int32_t a, int32_t b, int64_t target_a, int64_t n) {
        const int64_t tol = std::max<int64_t>(1, n / 200);  // 0.5 % imbalance
        int64_t size_a = 0;
        for (int32_t v : nodes) size_a += label[v] == a;
        auto gain = [&](int32_t v) {
            int ext = 0, in = 0;
            const int32_t mine = label[v], oth = mine == a ? b : a;
            for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
                ext += label[g.adj[e]] == oth;
                in += label[g.adj[e]] == mine;
            }
            return ext - in;
        };
        for (int pass = 0; pass < 8; ++pass) {
            int64_t moved = 0;
            // bucket the boundary nodes by gain once per pass; a moved node is locked for the rest of the pass
            std::vector<std::pair<int, int32_t>> cand;
            for (int32_t v : nodes) {
                const int gv = gain(v);
                if (gv > 0) cand.emplace_back(-gv, v);
            }
            if (cand.empty()) break;
            std::sort(cand.begin(), cand.end());
            ++stamp;
            for (auto& c : cand) {
                const int32_t v = c.second;
                if (mark[v] == stamp) continue;
                if (gain(v) <= 0) continue;  // a neighbour moved meanwhile
                const bool from_a = label[v] == a;
                const int64_t new_a = size_a + (from_a ? -1 : 1);
                if (std::llabs(new_a - target_a) > tol) continue;
                label[v] = from_a ? b : a;
                size_a = new_a;
                mark[v] = stamp;
                ++moved;
            }
            if (moved == 0) break;
        }
        // restore the exact target (the tolerance above lets the cut improve; the caller wants equal counts):
        // move the boundary nodes that cost the least, best gain first
        while (size_a != target_a) {
            const bool need_more_a = size_a < target_a;
            const int32_t from = need_more_a ? b : a, to = need_more_a ? a : b;
            const int64_t need = std::llabs(target_a - size_a);
            std::vector<std::pair<int, int32_t>> cand;
            for (int32_t v : nodes) {
                if (label[v] != from) continue;
                bool boundary = g.ptr[v + 1] == g.ptr[v];  // isolated points can go anywhere
                for (int64_t e = g.ptr[v]; e < g.ptr[v + 1] && !boundary; ++e) boundary = label[g.adj[e]] == to;
                if (boundary) cand.emplace_back(-gain(v), v);
            }
            if (cand.empty())  // nothing touches the other side (disconnected remainder): take any node
                for (int32_t v : nodes)
                    if (label[v] == from) {
                        cand.emplace_back(0, v);
                        if ((int64_t)cand.size() == need) break;
                    }
            std::sort(cand.begin(), cand.end());
            const int64_t take = std::min<int64_t>(need, (int64_t)cand.size());
            for (int64_t q = 0; q < take; ++q) label[cand[q].second] = to;
            size_a += need_more_a ? take : -take;
        }
    }

    // recursive k-way split of `nodes` (labelled lab) into labels written to `out` as part ids part0 .. part0+k-1
    void split(std::vector<int32_t>& nodes, int32_t lab, int32_t k, int32_t part0, std::vector<int32_t>& out) {
        if (k == 1) {
            for (int32_t v : nodes) out[v] = part0;
            return;
        }
        const int32_t k_left = k / 2;
        const int64_t n_left = (int64_t)nodes.size() * k_left / k;
        const int32_t other = bisect(nodes, lab, n_left);
        std::vector<int32_t> left, right;
        left.reserve(n_left);
        right.reserve(nodes.size() - n_left);
        for (int32_t v : nodes) (label[v] == lab ? left : right).push_back(v);
        std::vector<int32_t>().swap(nodes);
        split(left, lab, k_left, part0, out);
        split(right, other, k - k_left, part0 + k_left, out);
    }
};

}  // namespace

extern "C" int32_t fvm_partition_graph(int64_t n_points, const int32_t* triangles, int64_t n_triangles, int32_t index_base,
                                       int32_t n_parts, int32_t* owner) {
    if (!triangles || !owner || n_points <= 0 || n_triangles <= 0 || n_parts < 1 || n_parts > n_points)
        return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_graph: bad arguments");
    if (n_points >= INT32_MAX) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_graph: mesh too large for int32 indices");
    for (int64_t i = 0; i < 3 * n_triangles; ++i)
        if (triangles[i] - index_base < 0 || triangles[i] - index_base >= n_points)
            return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_graph: triangle vertex out of range");
    const Graph g = build_graph(n_points, triangles, n_triangles, index_base);
    std::vector<int32_t> label(n_points, 0), out(n_points, 0), nodes(n_points);
    std::iota(nodes.begin(), nodes.end(), 0);
    Bisector B(g, label, 1);
    B.split(nodes, 0, n_parts, 0, out);
    std::copy(out.begin(), out.end(), owner);
    return FVM_OK;
}

// number of node-graph edges whose endpoints have different owners (the halo volume is proportional to it)
extern "C" int32_t fvm_partition_edge_cut(int64_t n_points, const int32_t* triangles, int64_t n_triangles, int32_t index_base,
                                          const int32_t* owner, int64_t* cut) {
    if (!triangles || !owner || !cut || n_points <= 0 || n_triangles <= 0) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_partition_edge_cut: bad arguments");
    const Graph g = build_graph(n_points, triangles, n_triangles, index_base);
    int64_t c = 0;
    for (int64_t v = 0; v < n_points; ++v)
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) c += owner[v] != owner[g.adj[e]];
    *cut = c / 2;
    return FVM_OK;
}
