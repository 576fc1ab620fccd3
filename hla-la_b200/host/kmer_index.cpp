// Host builder of the k-mer index (see kmer_index.h for what is enumerated and in which order).
#include "kmer_index.h"
#include "../csrc/kmer_hash.h"

#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <functional>
#include <atomic>
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
    const int32_t last_start_level = g.n_levels - 1 - k;
    int32_t n_start = 0; while (n_start < g.n_nodes && g.node_level[n_start] <= last_start_level) n_start++;   // start nodes are a prefix of the flat node array
    // ---- enumeration: chunks of consecutive start nodes on a thread pool; chunk order == recording order
    struct Chunk { std::vector<uint8_t> kmer; std::vector<int64_t> off; std::vector<int32_t> edges; };
    const int32_t chunk_nodes = 16384; const size_t n_chunks = ((size_t)n_start + chunk_nodes - 1) / chunk_nodes;
    std::vector<Chunk> chunks(n_chunks);
    std::atomic<size_t> next(0); std::atomic<int64_t> total_edges(0); std::atomic<bool> too_many(false);
    auto enumerate = [&]() {
        Enumerator E(g, k);
        struct Basis { int32_t last_ord; int64_t off; int32_t len; };
        std::vector<Basis> bases; std::vector<int32_t> basis_edges;
        for (;;) {
            const size_t ci = next.fetch_add(1); if (ci >= n_chunks || too_many) break;
            Chunk& C = chunks[ci]; C.off.assign(1, 0);
            auto record = [&]() {
                for (int32_t e : E.stack) if (g.edge_emis[e] != '_') C.kmer.push_back(g.edge_emis[e]);
                C.edges.insert(C.edges.end(), E.stack.begin(), E.stack.end()); C.off.push_back((int64_t)C.edges.size());
            };
            const int32_t n0 = (int32_t)(ci * chunk_nodes), n1 = std::min<int32_t>(n_start, n0 + chunk_nodes);
            for (int32_t n = n0; n < n1; n++) {
                if (g.node_level[n] == 0) { if (n == 0) E.dfs(0, k, true, record); continue; }
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
            if (total_edges.fetch_add((int64_t)C.edges.size()) + (int64_t)C.edges.size() > max_edges) too_many = true;
        }
    };
    const bool trace = getenv("HLALA_KMER_TRACE") != nullptr; auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (!trace) return; auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[kmer-index] %-12s %7.2f s\n", what, std::chrono::duration<double>(t1 - t0).count()); t0 = t1; };
    unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("HLALA_KMER_THREADS")) nt = (unsigned)std::max(1, atoi(e));
    if (n_chunks < 2) nt = 1;
    auto run_pool = [&](const std::function<void()>& fn) { std::vector<std::thread> th; for (unsigned t = 1; t < nt; t++) th.emplace_back(fn); fn(); for (auto& t : th) t.join(); };
    run_pool(enumerate); lap("enumerate");
    if (too_many) throw std::runtime_error("k-mer index: more than max_edges path edges (graph too branched for this k)");
    // ---- flat view of the records in recording order
    std::vector<int64_t> chunk_first(n_chunks + 1, 0);
    for (size_t c = 0; c < n_chunks; c++) chunk_first[c + 1] = chunk_first[c] + (int64_t)chunks[c].off.size() - 1;
    const int64_t n_rec = chunk_first[n_chunks];
    std::vector<const uint8_t*> key((size_t)n_rec); std::vector<const int32_t*> path((size_t)n_rec); std::vector<int32_t> plen((size_t)n_rec);
    next = 0;
    run_pool([&]() { for (;;) { const size_t c = next.fetch_add(1); if (c >= n_chunks) break; const Chunk& C = chunks[c];
        for (size_t i = 0; i + 1 < C.off.size(); i++) { const size_t r = (size_t)chunk_first[c] + i; key[r] = C.kmer.data() + i * (size_t)k; path[r] = C.edges.data() + C.off[i]; plen[r] = (int32_t)(C.off[i + 1] - C.off[i]); } } });
    lap("flatten");
    // ---- group by k-mer string (ascending bytes), recording order kept inside a group: buckets by the first two bytes (order-preserving),
    //      records scattered to their bucket in recording order, every bucket sorted stably on its own
    std::vector<int64_t> bucket_off(65537, 0);
    for (int64_t r = 0; r < n_rec; r++) bucket_off[(((size_t)key[(size_t)r][0]) << 8 | key[(size_t)r][1]) + 1]++;
    for (size_t b = 0; b < 65536; b++) bucket_off[b + 1] += bucket_off[b];
    std::vector<int64_t> order((size_t)n_rec);
    { std::vector<int64_t> at(bucket_off.begin(), bucket_off.end() - 1); for (int64_t r = 0; r < n_rec; r++) order[(size_t)at[((size_t)key[(size_t)r][0]) << 8 | key[(size_t)r][1]]++] = r; }
    lap("scatter");
    std::vector<size_t> busy; for (size_t b = 0; b < 65536; b++) if (bucket_off[b + 1] > bucket_off[b]) busy.push_back(b);
    std::sort(busy.begin(), busy.end(), [&](size_t x, size_t y) { return bucket_off[x + 1] - bucket_off[x] > bucket_off[y + 1] - bucket_off[y]; });   // largest buckets first
    next = 0;
    run_pool([&]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= busy.size()) break; const size_t b = busy[i];
        std::stable_sort(order.begin() + bucket_off[b], order.begin() + bucket_off[b + 1], [&](int64_t x, int64_t y) { return memcmp(key[(size_t)x], key[(size_t)y], (size_t)k) < 0; }); } });
    lap("sort");
    // ---- output arrays
    out.n_pos = n_rec; out.pos_edge_off.assign((size_t)n_rec + 1, 0); out.kmer_pos_off.clear();
    {   // first record of every k-mer: flags in parallel, then one cheap sequential pass for the offsets
        std::vector<uint8_t> first((size_t)n_rec, 0);
        next = 0; const int64_t blk = 262144;
        run_pool([&]() { for (;;) { const int64_t i0 = (int64_t)next.fetch_add(1) * blk; if (i0 >= n_rec) break; const int64_t i1 = std::min(n_rec, i0 + blk);
            for (int64_t i = i0; i < i1; i++) first[(size_t)i] = (i == 0 || memcmp(key[(size_t)order[(size_t)i]], key[(size_t)order[(size_t)i - 1]], (size_t)k) != 0) ? 1 : 0; } });
        int64_t nk = 0; for (int64_t i = 0; i < n_rec; i++) nk += first[(size_t)i];
        out.kmer_pos_off.reserve((size_t)nk + 1);
        for (int64_t i = 0; i < n_rec; i++) { if (first[(size_t)i]) out.kmer_pos_off.push_back(i); out.pos_edge_off[(size_t)i + 1] = out.pos_edge_off[(size_t)i] + plen[(size_t)order[(size_t)i]]; }
        out.kmer_bytes.resize((size_t)nk * k);
        next = 0;
        run_pool([&]() { for (;;) { const int64_t j0 = (int64_t)next.fetch_add(1) * blk; if (j0 >= nk) break; const int64_t j1 = std::min(nk, j0 + blk);
            for (int64_t j = j0; j < j1; j++) memcpy(out.kmer_bytes.data() + (size_t)j * k, key[(size_t)order[(size_t)out.kmer_pos_off[(size_t)j]]], (size_t)k); } });
    }
    out.kmer_pos_off.push_back(n_rec);
    out.n_kmers = (int64_t)out.kmer_pos_off.size() - 1;
    lap("unique");
    out.pos_edges.resize((size_t)out.pos_edge_off[(size_t)n_rec]);
    next = 0; const int64_t copy_block = 65536;
    run_pool([&]() { for (;;) { const int64_t i0 = (int64_t)next.fetch_add(1) * copy_block; if (i0 >= n_rec) break; const int64_t i1 = std::min(n_rec, i0 + copy_block);
        for (int64_t i = i0; i < i1; i++) { const int64_t r = order[(size_t)i]; memcpy(out.pos_edges.data() + out.pos_edge_off[(size_t)i], path[(size_t)r], (size_t)plen[(size_t)r] * 4); } } });
    lap("copy");
    uint64_t sz = 16; while (sz < (uint64_t)out.n_kmers * 2) sz <<= 1;
    if (sz > ((uint64_t)1 << 31)) throw std::runtime_error("k-mer index: too many distinct k-mers");
    out.ht.assign((size_t)sz, 0); out.ht_mask = (uint32_t)(sz - 1);
    next = 0;   // linear probing with compare-and-swap: the slot a k-mer lands in depends on the insertion order, what a lookup finds does not
    run_pool([&]() { const int64_t blk = 65536; for (;;) { const int64_t i0 = (int64_t)next.fetch_add(1) * blk; if (i0 >= out.n_kmers) break; const int64_t i1 = std::min<int64_t>(out.n_kmers, i0 + blk);
        for (int64_t i = i0; i < i1; i++) {
            uint32_t h = (uint32_t)kmer_hash(out.kmer_bytes.data() + (size_t)i * k, k) & out.ht_mask;
            while (!__sync_bool_compare_and_swap(&out.ht[h], 0, (int32_t)(i + 1))) h = (h + 1) & out.ht_mask;
        } } });
    lap("hash");
}

} // namespace hlala
