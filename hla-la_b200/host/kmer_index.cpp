// Host builder of the k-mer index (see kmer_index.h for what is enumerated and in which order).
#include "kmer_index.h"
#include "../csrc/kmer_hash.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>

namespace hlala {

namespace {

struct Enumerator {
    const FlatGraph& g; int k;
    std::vector<int32_t> stack;    // current edge path
    Enumerator(const FlatGraph& graph, int kk) : g(graph), k(kk) {}
    // all continuations from `node` with exactly `need` more non-gap edges, ending with a non-gap edge, depth first in canonical
    // edge order (forwardScanRec, GraphAndEdgeIndex.cpp:1061-1160; the collapsing of single-gap runs there only shortens the recursion)
    template <class F> void dfs(int32_t node, int need, bool first_must_be_base, F&& emit) {
        for (int32_t i = g.node_out_off[node]; i < g.node_out_off[node + 1]; i++) {
            const int32_t e = g.node_out[i]; const bool gap = g.edge_emis[e] == '_';
            if (first_must_be_base && gap) continue;
            stack.push_back(e);
            if (!gap) { if (need == 1) emit(); else dfs(g.edge_to[e], need - 1, false, emit); }
            else dfs(g.edge_to[e], need, false, emit);
            stack.pop_back();
        }
    }
};

} // namespace

int64_t KmerIndex::find(const uint8_t* kmer) const {
    if (n_kmers == 0) return -1;
    uint32_t h = (uint32_t)kmer_hash(kmer, k) & ht_mask;
    for (;;) {
        const int32_t v = ht[h];
        if (v == 0) return -1;
        if (memcmp(kmer_bytes.data() + (size_t)(v - 1) * k, kmer, (size_t)k) == 0) return v - 1;
        h = (h + 1) & ht_mask;
    }
}

void build_kmer_index(const FlatGraph& g, int k, KmerIndex& out, int64_t max_edges) {
    if (k < 2 || k > 255) throw std::runtime_error("k-mer index: k must be in [2, 255]");
    out = KmerIndex(); out.k = k;
    // nodes reachable from the first node of level 0 (the reference seeds its window there, :443,:504)
    std::vector<uint8_t> reach((size_t)g.n_nodes, 0);
    if (g.n_nodes > 0) {
        reach[0] = 1;   // flat nodes are sorted by (level, ordinal): node 0 is *NodesPerLevel[0].begin()
        for (int32_t n = 0; n < g.n_nodes; n++) if (reach[n]) for (int32_t i = g.node_out_off[n]; i < g.node_out_off[n + 1]; i++) reach[g.edge_to[g.node_out[i]]] = 1;
    }
    // records in recording order: (k-mer bytes, path)
    std::vector<uint8_t> rec_kmer; std::vector<int64_t> rec_off(1, 0); std::vector<int32_t> rec_edges;
    Enumerator E(g, k);
    auto record = [&]() {
        for (int32_t e : E.stack) if (g.edge_emis[e] != '_') rec_kmer.push_back(g.edge_emis[e]);
        rec_edges.insert(rec_edges.end(), E.stack.begin(), E.stack.end()); rec_off.push_back((int64_t)rec_edges.size());
        if ((int64_t)rec_edges.size() > max_edges) throw std::runtime_error("k-mer index: more than max_edges path edges (graph too branched for this k)");
    };
    struct Basis { int32_t last_ord; int64_t off; int32_t len; };
    std::vector<Basis> bases; std::vector<int32_t> basis_edges;
    const int32_t last_start_level = g.n_levels - 1 - k;
    for (int32_t n = 0; n < g.n_nodes; n++) {
        const int32_t lvl = g.node_level[n];
        if (lvl > last_start_level) break;
        if (lvl == 0) { if (n == 0) E.dfs(0, k, true, record); continue; }
        if (!reach[n]) continue;
        bases.clear(); basis_edges.clear();
        E.dfs(n, k - 1, true, [&]() { bases.push_back({g.node_ord[g.edge_to[E.stack.back()]], (int64_t)basis_edges.size(), (int32_t)E.stack.size()}); basis_edges.insert(basis_edges.end(), E.stack.begin(), E.stack.end()); });
        std::stable_sort(bases.begin(), bases.end(), [](const Basis& a, const Basis& b) { return a.last_ord < b.last_ord; });   // depth-first order already is the edge-sequence order
        for (const Basis& b : bases) {
            E.stack.assign(basis_edges.begin() + b.off, basis_edges.begin() + b.off + b.len);
            E.dfs(g.edge_to[E.stack.back()], 1, false, record);
        }
        E.stack.clear();
    }
    const int64_t n_rec = (int64_t)rec_off.size() - 1;
    // group by k-mer string (ascending bytes), recording order kept inside a group
    std::vector<int64_t> order((size_t)n_rec); std::iota(order.begin(), order.end(), 0);
    const uint8_t* kb = rec_kmer.data();
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return memcmp(kb + a * k, kb + b * k, (size_t)k) < 0; });
    out.n_pos = n_rec; out.pos_edge_off.assign(1, 0); out.pos_edges.reserve(rec_edges.size()); out.kmer_pos_off.clear();
    for (int64_t i = 0; i < n_rec; i++) {
        const int64_t r = order[(size_t)i];
        if (i == 0 || memcmp(kb + r * k, kb + order[(size_t)i - 1] * k, (size_t)k) != 0) { out.kmer_pos_off.push_back(i); out.kmer_bytes.insert(out.kmer_bytes.end(), kb + r * k, kb + (r + 1) * k); }
        out.pos_edges.insert(out.pos_edges.end(), rec_edges.begin() + rec_off[(size_t)r], rec_edges.begin() + rec_off[(size_t)r + 1]);
        out.pos_edge_off.push_back((int64_t)out.pos_edges.size());
    }
    out.kmer_pos_off.push_back(n_rec);
    out.n_kmers = (int64_t)out.kmer_pos_off.size() - 1;
    uint64_t sz = 16; while (sz < (uint64_t)out.n_kmers * 2) sz <<= 1;
    if (sz > ((uint64_t)1 << 31)) throw std::runtime_error("k-mer index: too many distinct k-mers");
    out.ht.assign((size_t)sz, 0); out.ht_mask = (uint32_t)(sz - 1);
    for (int64_t i = 0; i < out.n_kmers; i++) {
        uint32_t h = (uint32_t)kmer_hash(out.kmer_bytes.data() + (size_t)i * k, k) & out.ht_mask;
        while (out.ht[h]) h = (h + 1) & out.ht_mask;
        out.ht[h] = (int32_t)(i + 1);
    }
}

} // namespace hlala
