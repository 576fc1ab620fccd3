// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see hlala_oracle.cpp for who may load this library).
//
// Plain C++ restatement of the reference's k-mer seeding, following its own procedure step by step with STL containers
// (std::map keyed by canonical ordinals where the reference keys by pointer or by a string of pointers):
//   forwardScan / forwardScanRec                Graph/GraphAndEdgeIndex.cpp:1016-1160
//   GraphAndEdgeIndex::Index                    Graph/GraphAndEdgeIndex.cpp:428-959 (the level-by-level window of attached (k-1)-mers)
//   GraphAndEdgeIndex::queryIndex               Graph/GraphAndEdgeIndex.cpp:986-997
//   GraphAndEdgeIndex::findChains               Graph/GraphAndEdgeIndex.cpp:39-356
// It shares no code with the product (hla-la_b200/host/kmer_index.cpp enumerates paths directly and sorts; this file slides the
// window like the reference does). Pinning: tests/test_kmer_oracle.py compares the whole index (k-mers, positions, their order,
// edge paths) and findChains' results with the UNMODIFIED GraphAndEdgeIndex.cpp compiled into oracle/_ref/libhlala_ref.so.
//
// Order conventions (oracle/ref_driver.cpp pins pointer order == ordinal order): std::set<Edge*> iterates edge ordinals ascending;
// the edgeTargetCache key "string of setw(15) pointers" compares like the vector of edge ordinals.
// nodes_jumpOverGaps stays empty in the reference (fillEdgeJumper is never called), so the jump loops are not restated.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct KGraph {
    int n_levels = 0;
    std::vector<int> node_level; std::vector<int> efrom, eto; std::vector<char> emis;
    std::vector<std::vector<int>> out;            // per node: edge ordinals ascending
    std::vector<std::vector<int>> level_nodes;    // per level: node ordinals ascending
};

struct KInfo { std::string chars; std::vector<int> edges; bool gapEdge = false; };   // kMerInfo (GraphAndEdgeIndex.h:26-50)
struct KAtNode { std::string chars; std::vector<int> edges; };                        // kMerAtNode (:52-75)

// forwardScanRec (:1061-1160). The reference accumulates back to front and reverses in forwardScan (:1030-1032); built front to back here.
void scan_rec(const KGraph& g, int node, int depth, int realdepth, int limit, int firstEdgeGap, KInfo& cur, std::vector<KInfo>& out) {
    if (limit == depth) { out.push_back(cur); return; }
    for (int e : g.out[node]) {
        const bool gap = g.emis[e] == '_';
        if (realdepth == 0) { if (firstEdgeGap == 1 && !gap) continue; if (firstEdgeGap == -1 && gap) continue; }   // :1090-1108
        if (!gap) {
            cur.chars.push_back(g.emis[e]); cur.edges.push_back(e);
            scan_rec(g, g.eto[e], depth + 1, realdepth + 1, limit, firstEdgeGap, cur, out);
            cur.chars.pop_back(); cur.edges.pop_back();
        } else {
            // :1125-1154: a run of single gap edges is walked without recursion, then the scan continues from its end
            size_t before = cur.edges.size(); cur.edges.push_back(e); int t = g.eto[e];
            if (g.out[t].size() == 1 && g.emis[g.out[t][0]] == '_')
                while (g.out[t].size() == 1 && g.emis[g.out[t][0]] == '_') { cur.edges.push_back(g.out[t][0]); t = g.eto[g.out[t][0]]; }
            scan_rec(g, t, depth, realdepth + 1, limit, firstEdgeGap, cur, out);
            cur.edges.resize(before);
        }
    }
}
std::vector<KInfo> forward_scan(const KGraph& g, int start, int limit, int firstEdgeGap) { std::vector<KInfo> r; KInfo cur; scan_rec(g, start, 0, 0, limit, firstEdgeGap, cur, r); return r; }

struct KIndex {
    KGraph g; int k = 0;
    std::map<std::string, std::vector<std::vector<int>>> kmers;    // kMers (GraphAndEdgeIndex.h:127)
    // results of the last find_chains batch
    std::vector<int64_t> r_chain_off; std::vector<int32_t> r_begin, r_end; std::vector<int64_t> r_edge_off; std::vector<int32_t> r_edges;

    void index() {   // :428-959
        const int levels = g.n_levels;
        const int N0 = g.level_nodes.at(0).at(0);   // classicalN0 (:443)
        std::map<int, std::vector<KAtNode>> assigned;   // originalGraphAssignedKMers (:468)
        for (int level = 0; level <= levels - 1 - k; level++) {
            std::map<int, std::map<std::vector<int>, std::vector<int>>> cache;   // edgeTargetCache (:494): target node -> tail edges -> new k-mer edges
            std::vector<KInfo> infos;                                             // newEdgeNodeInfos (:495), indexed by new k-mer edge
            auto attach = [&](const KInfo& ki, int target) {
                std::vector<int> tail(ki.edges.begin() + 1, ki.edges.end());      // substr(pointerStrintLength): everything but the first edge
                cache[target][tail].push_back((int)infos.size()); infos.push_back(ki);
            };
            if (level == 0) {
                for (KInfo& ki : forward_scan(g, N0, k, -1)) { kmers[ki.chars].push_back(ki.edges); attach(ki, g.eto[ki.edges.back()]); }   // :501-530
                for (KInfo& ki : forward_scan(g, N0, k - 1, 1)) { ki.gapEdge = true; attach(ki, g.eto[ki.edges.back()]); }                  // :534-565
            } else {
                for (int node : g.level_nodes[level]) {   // :570-576
                    auto it = assigned.find(node);
                    if (it == assigned.end()) throw std::runtime_error("reference assertion would fail: originalGraphAssignedKMers.count(originalNode) > 0");   // :601
                    for (const KAtNode& basis : it->second) {   // :605
                        if (g.emis[basis.edges.at(0)] != '_') {   // :621
                            const int last = g.eto[basis.edges.back()];
                            std::vector<KInfo> ext = forward_scan(g, last, 1, 0);   // :645
                            if (ext.empty()) {   // :648-688: end of graph, a gap k-mer keeps the window moving
                                KInfo ni; ni.gapEdge = true; ni.chars = basis.chars; ni.edges = basis.edges; attach(ni, g.eto[ni.edges.back()]);
                            } else for (const KInfo& x : ext) {   // :694-760
                                KInfo ni; ni.gapEdge = false; ni.chars = basis.chars + x.chars; ni.edges = basis.edges; ni.edges.insert(ni.edges.end(), x.edges.begin(), x.edges.end());
                                if ((int)ni.chars.size() != k) throw std::runtime_error("reference assertion would fail: kMer_string.size() == kMerSize");
                                kmers[ni.chars].push_back(ni.edges); attach(ni, g.eto[x.edges.back()]);
                            }
                        } else {   // :763-812: the attached window starts with a gap edge; it only shifts
                            KInfo ni; ni.gapEdge = true; ni.chars = basis.chars; ni.edges = basis.edges; attach(ni, g.eto[ni.edges.back()]);
                        }
                    }
                }
            }
            assigned.clear();   // :816
            for (auto& tn : cache) for (auto& ts : tn.second) {   // :821-905, map order: target node, then tail
                const KInfo& first = infos[ts.second.at(0)];
                KAtNode a; a.chars = first.chars; a.edges.assign(first.edges.begin() + 1, first.edges.end());
                if (!first.gapEdge) a.chars.erase(0, 1);   // :858-862
                assigned[g.eto[first.edges.at(0)]].push_back(a);   // :880,:898
            }
        }
    }

    // findChains (:39-356)
    struct Chain { int begin = -1, end = -1; std::vector<int> edges; };
    std::vector<std::vector<int>> chain_forward_scan(int startEdge) const {   // the lambda at :134-187
        std::vector<std::vector<int>> found, running; running.push_back(std::vector<int>(1, startEdge));
        while (!running.empty()) {
            for (int eI = (int)running.size() - 1; eI >= 0; eI--) {
                const int tip = running[eI].back();
                if (g.emis[tip] == '_') {
                    const std::vector<int>& next = g.out[g.eto[tip]];
                    if (next.empty()) running.erase(running.begin() + eI);
                    else if (next.size() == 1) running[eI].push_back(next[0]);
                    else {
                        const std::vector<int> tmpl = running[eI];
                        running[eI].push_back(next[0]);
                        for (size_t m = 1; m < next.size(); m++) { std::vector<int> n2 = tmpl; n2.push_back(next[m]); running.push_back(n2); }
                    }
                } else { found.push_back(running[eI]); running.erase(running.begin() + eI); }
            }
        }
        return found;
    }
    std::vector<Chain> find_chains(const std::string& seq) const {
        std::vector<Chain> ret, running;
        if ((int)seq.size() < k) return ret;   // :50-54 (partitionStringIntokMers, Utilities.cpp:1381)
        auto query = [&](const std::string& s) -> const std::vector<std::vector<int>>* { auto it = kmers.find(s); return it == kmers.end() ? nullptr : &it->second; };
        if (auto* pos = query(seq.substr(0, (size_t)k))) for (auto& p : *pos) { Chain c; c.begin = 0; c.end = k - 1; c.edges = p; running.push_back(c); }   // :60-69
        for (int seqI = k; seqI < (int)seq.size(); seqI++) {
            const char ch = seq[(size_t)seqI];
            for (int chainI = (int)running.size() - 1; chainI >= 0; chainI--) {   // :80
                const int target = g.eto[running[chainI].edges.back()];
                std::vector<std::vector<int>> compat;
                for (int e : g.out[target]) for (auto& path : chain_forward_scan(e)) if (g.emis[path.back()] == ch) compat.push_back(path);   // :190-236
                if (compat.empty()) { ret.push_back(running[chainI]); running.erase(running.begin() + chainI); }   // :250-279
                else if (compat.size() == 1) { running[chainI].end = seqI; running[chainI].edges.insert(running[chainI].edges.end(), compat[0].begin(), compat[0].end()); }
                else {   // :286-304
                    const Chain tmpl = running[chainI];
                    running[chainI].end = seqI; running[chainI].edges.insert(running[chainI].edges.end(), compat[0].begin(), compat[0].end());
                    for (size_t i = 1; i < compat.size(); i++) { Chain n2 = tmpl; n2.end = seqI; n2.edges.insert(n2.edges.end(), compat[i].begin(), compat[i].end()); running.push_back(n2); }
                }
            }
            if (auto* pos = query(seq.substr((size_t)(seqI - k + 1), (size_t)k))) for (auto& p : *pos) {   // :307-341
                bool represented = false;
                for (int chainI = (int)running.size() - 1; chainI >= 0; chainI--) {
                    const Chain& ex = running[chainI];
                    if (g.eto[ex.edges.back()] == g.eto[p.back()] && ex.edges.size() >= p.size() && g.efrom[ex.edges[ex.edges.size() - p.size()]] == g.efrom[p[0]]) { represented = true; break; }
                }
                if (!represented) { Chain c; c.begin = seqI - k + 1; c.end = seqI; c.edges = p; running.push_back(c); }
            }
        }
        for (auto& c : running) ret.push_back(c);   // :344-348
        return ret;
    }
};

std::string g_kerr;

} // namespace

extern "C" {

const char* hlala_oracle_kmer_last_error() { return g_kerr.c_str(); }

// graph arrays in canonical (ordinal) order, as hlala_oracle_graph_export / hlala_ref_graph_export deliver them
void* hlala_oracle_kmer_open(long long n_nodes, const int32_t* node_level, long long n_edges, const int32_t* edge_from, const int32_t* edge_to, const uint8_t* edge_emis, int k) {
    try {
        KIndex* X = new KIndex(); X->k = k; KGraph& g = X->g;
        g.node_level.assign(node_level, node_level + n_nodes); g.efrom.assign(edge_from, edge_from + n_edges); g.eto.assign(edge_to, edge_to + n_edges);
        g.emis.assign((const char*)edge_emis, (const char*)edge_emis + n_edges);
        for (long long i = 0; i < n_nodes; i++) g.n_levels = std::max(g.n_levels, node_level[i] + 1);
        g.out.resize((size_t)n_nodes); g.level_nodes.resize((size_t)g.n_levels);
        for (long long i = 0; i < n_nodes; i++) g.level_nodes[(size_t)node_level[i]].push_back((int)i);
        for (long long e = 0; e < n_edges; e++) g.out[(size_t)edge_from[e]].push_back((int)e);
        X->index();
        return X;
    } catch (const std::exception& e) { g_kerr = e.what(); return nullptr; }
}
void hlala_oracle_kmer_free(void* idx) { delete (KIndex*)idx; }

int hlala_oracle_kmer_dump_sizes(void* idx, long long* n_kmers, long long* n_pos, long long* n_edges) {
    KIndex* X = (KIndex*)idx; *n_kmers = (long long)X->kmers.size(); *n_pos = 0; *n_edges = 0;
    for (auto& kv : X->kmers) { *n_pos += (long long)kv.second.size(); for (auto& p : kv.second) *n_edges += (long long)p.size(); }
    return 0;
}
int hlala_oracle_kmer_dump(void* idx, uint8_t* kmer_bytes, int64_t* pos_off, int64_t* edge_off, int32_t* edges) {
    KIndex* X = (KIndex*)idx; int64_t np = 0, ne = 0; size_t i = 0;
    for (auto& kv : X->kmers) {
        memcpy(kmer_bytes + i * (size_t)X->k, kv.first.data(), (size_t)X->k); pos_off[i++] = np;
        for (auto& p : kv.second) { edge_off[np++] = ne; for (int e : p) edges[ne++] = e; }
    }
    pos_off[i] = np; edge_off[np] = ne; return 0;
}
int hlala_oracle_find_chains(void* idx, long long n_reads, const int64_t* read_off, const uint8_t* bases, long long* n_chains, long long* n_edges, double* seconds) {
    KIndex* X = (KIndex*)idx;
    try {
        X->r_chain_off.assign(1, 0); X->r_begin.clear(); X->r_end.clear(); X->r_edge_off.assign(1, 0); X->r_edges.clear();
        auto t0 = std::chrono::steady_clock::now();
        for (long long r = 0; r < n_reads; r++) {
            std::string seq((const char*)bases + read_off[r], (size_t)(read_off[r + 1] - read_off[r]));
            for (const KIndex::Chain& c : X->find_chains(seq)) {
                X->r_begin.push_back(c.begin); X->r_end.push_back(c.end); X->r_edges.insert(X->r_edges.end(), c.edges.begin(), c.edges.end()); X->r_edge_off.push_back((int64_t)X->r_edges.size());
            }
            X->r_chain_off.push_back((int64_t)X->r_begin.size());
        }
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *n_chains = (long long)X->r_begin.size(); *n_edges = (long long)X->r_edges.size(); return 0;
    } catch (const std::exception& e) { g_kerr = e.what(); return -1; }
}
int hlala_oracle_find_chains_fetch(void* idx, int64_t* chain_off, int32_t* begin, int32_t* end, int64_t* edge_off, int32_t* edges) {
    KIndex* X = (KIndex*)idx;
    memcpy(chain_off, X->r_chain_off.data(), X->r_chain_off.size() * 8); memcpy(begin, X->r_begin.data(), X->r_begin.size() * 4); memcpy(end, X->r_end.data(), X->r_end.size() * 4);
    memcpy(edge_off, X->r_edge_off.data(), X->r_edge_off.size() * 8); memcpy(edges, X->r_edges.data(), X->r_edges.size() * 4);
    return 0;
}

} // extern "C"
