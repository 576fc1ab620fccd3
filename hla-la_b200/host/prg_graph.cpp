#include "prg_graph.h"
#include "arrayfile.h"
#include <sys/stat.h>
#include <dirent.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_map>

namespace hlala {

namespace {


std::string read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open graph file: " + path);
    std::string s; fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    s.resize((size_t)n);
    if (n > 0 && fread(&s[0], 1, (size_t)n, f) != (size_t)n) { fclose(f); throw std::runtime_error("short read: " + path); }
    fclose(f);
    return s;
}

// fields of one line split on "|||" (Graph.cpp:26, boost::iter_split + first_finder)
void split_fields(const char* b, const char* e, std::vector<std::pair<const char*, const char*>>& out) {
    out.clear();
    const char* p = b;
    while (true) {
        const char* hit = nullptr;
        for (const char* q = p; q + 3 <= e;) {
            q = (const char*)memchr(q, '|', (size_t)(e - q));
            if (!q || q + 3 > e) break;
            if (q[1] == '|' && q[2] == '|') { hit = q; break; }
            q++;
        }
        if (!hit) { out.emplace_back(p, e); break; }
        out.emplace_back(p, hit); p = hit + 3;
    }
}
long long to_ll(const std::pair<const char*, const char*>& f) {
    const char* p = f.first; bool neg = false; long long v = 0;
    if (p < f.second && (*p == '-' || *p == '+')) { neg = (*p == '-'); p++; }
    if (p == f.second) throw std::runtime_error("graph.txt: cannot parse integer field '" + std::string(f.first, f.second) + "'");
    for (; p < f.second; p++) { if (*p < '0' || *p > '9') throw std::runtime_error("graph.txt: cannot parse integer field '" + std::string(f.first, f.second) + "'"); v = v * 10 + (*p - '0'); }
    return neg ? -v : v;
}
uint64_t fnv(const char* b, const char* e) { uint64_t h = 1469598103934665603ull; for (const char* p = b; p < e; p++) { h ^= (unsigned char)*p; h *= 1099511628211ull; } return h; }

} // namespace

int32_t FlatGraph::z_of(int32_t node) const { return node - level_node_off[node_level[node]]; }

static void compute_gap_paths(FlatGraph& g);
static void compute_gap_stretches(FlatGraph& g);
static void load_contigs(const std::string& dir, FlatGraph& g);

// Binary cache of the flat arrays next to graph.txt, reused when newer than its inputs — the role serializedGRAPH plays for the
// reference (processBAM.cpp:37-53); its Boost text archive cannot be read without Boost and holds nothing graph.txt lacks.
static const int64_t CACHE_VERSION = 4;
static time_t mtime_of(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 ? st.st_mtime : 0; }

#define FG_I32(X) X(level_node_off) X(node_ord) X(node_level) X(level_edge_off) X(edge_from) X(edge_to) X(edge_ord) X(ord_to_edge) X(node_out_off) X(node_out) \
    X(node_in_off) X(node_in) X(path_off) X(path_edges) X(path_from) X(path_to) X(jump_fwd_off) X(jump_fwd_path) X(jump_bwd_off) X(jump_bwd_path) \
    X(contig_prg_id) X(contig_level) X(contig_tr_len) X(anchor_off) X(anchor_prg_id) X(anchor_pos)
#define FG_U8(X) X(edge_emis) X(gap_stretch) X(contig_seq)

// Everything the cache is derived from: size and mtime of graph.txt, sequences.txt, the PRG-only FASTA and every translation/*.txt (contig_level and the
// anchors are baked into the cache), folded into one 64-bit value that is stored with the arrays and compared on load.
static int64_t inputs_signature(const std::string& dir) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](uint64_t v) { for (int i = 0; i < 8; i++) { h ^= (v >> (8 * i)) & 255u; h *= 0x100000001b3ull; } };
    auto add = [&](const std::string& p) { struct stat st; if (stat(p.c_str(), &st) != 0) { mix(0); return; } mix((uint64_t)st.st_size); mix((uint64_t)st.st_mtime); for (char c : p) mix((unsigned char)c); };
    add(dir + "/PRG/graph.txt"); add(dir + "/sequences.txt"); add(dir + "/mapping_PRGonly/referenceGenome.fa");
    std::vector<std::string> tr;
    if (DIR* dp = opendir((dir + "/translation").c_str())) { while (dirent* e = readdir(dp)) { std::string n = e->d_name; if (n != "." && n != "..") tr.push_back(n); } closedir(dp); }
    std::sort(tr.begin(), tr.end());
    for (const std::string& n : tr) add(dir + "/translation/" + n);
    return (int64_t)(h & 0x7fffffffffffffffull);
}

static void save_cache(const std::string& path, const FlatGraph& g, int64_t sig) {
    try {
        ArrayFile f;
        std::vector<int64_t> sigv{sig}; f.put("inputs_signature", DT_I64, sigv);
        std::vector<int64_t> meta{CACHE_VERSION, g.n_levels, g.n_nodes, g.n_edges, g.max_nodes_per_level, g.max_edges_per_level, g.n_paths, g.n_contigs};
        f.put("meta", DT_I64, meta);
#define X(n) f.put(#n, DT_I32, g.n);
        FG_I32(X)
#undef X
#define X(n) f.put(#n, DT_U8, g.n);
        FG_U8(X)
#undef X
        f.put("contig_off", DT_I64, g.contig_off);
        auto blob = [](const std::vector<std::string>& v) { std::vector<uint8_t> b; for (const std::string& s : v) { b.insert(b.end(), s.begin(), s.end()); b.push_back('\n'); } return b; };
        f.put("level_names", DT_U8, blob(g.level_names)); f.put("contig_bam_name", DT_U8, blob(g.contig_bam_name));
        const std::string tmp = path + ".tmp." + std::to_string((long long)getpid());      // two processes preparing the same PRG must not interleave their writes
        f.write(tmp);
        if (rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
    } catch (...) { /* a read-only PRG directory just means no cache */ }
}
static bool load_cache(const std::string& path, FlatGraph& g, int64_t sig) {
    try {
        ArrayFile f; f.read(path);
        uint64_t n = 0;
        { const int64_t* s = f.get<int64_t>("inputs_signature", &n); if (n != 1 || s[0] != sig) return false; }
        const int64_t* meta = f.get<int64_t>("meta", &n);
        if (n != 8 || meta[0] != CACHE_VERSION) return false;
        g.n_levels = (int32_t)meta[1]; g.n_nodes = (int32_t)meta[2]; g.n_edges = (int32_t)meta[3]; g.max_nodes_per_level = (int32_t)meta[4]; g.max_edges_per_level = (int32_t)meta[5];
        g.n_paths = (int32_t)meta[6]; g.n_contigs = (int32_t)meta[7];
#define X(nm) { uint64_t c; const int32_t* p = f.get<int32_t>(#nm, &c); g.nm.assign(p, p + c); }
        FG_I32(X)
#undef X
#define X(nm) { uint64_t c; const uint8_t* p = f.get<uint8_t>(#nm, &c); g.nm.assign(p, p + c); }
        FG_U8(X)
#undef X
        { uint64_t c; const int64_t* p = f.get<int64_t>("contig_off", &c); g.contig_off.assign(p, p + c); }
        auto unblob = [&](const char* nm, std::vector<std::string>& out) { uint64_t c; const uint8_t* p = f.get<uint8_t>(nm, &c); out.clear(); std::string cur; for (uint64_t i = 0; i < c; i++) { if (p[i] == '\n') { out.push_back(cur); cur.clear(); } else cur.push_back((char)p[i]); } };
        unblob("level_names", g.level_names); unblob("contig_bam_name", g.contig_bam_name);
        return true;
    } catch (...) { return false; }
}

void load_prg_dir(const std::string& dir, FlatGraph& g) {
    const std::string graph_path = dir + "/PRG/graph.txt";
    const std::string cache_path = dir + "/PRG/graph.hlala_b200.cache";
    {
        if (mtime_of(cache_path) && !getenv("HLALA_NO_GRAPH_CACHE")) {
            if (load_cache(cache_path, g, inputs_signature(dir))) return;
            g = FlatGraph();
        }
    }
    std::string txt = read_file(graph_path);

    // ---- section scan (Graph::readFromFile, Graph.cpp:2329-2559)
    struct Line { const char* b; const char* e; };
    std::vector<Line> code, nodes, edges;
    std::vector<std::string> patched;   // lines rewritten for the "|||||||" -> "|||SLASH|||" escape (Graph.cpp:2338-2365)
    patched.reserve(16);
    int mode = -1;
    const char* p = txt.data(); const char* end = p + txt.size();
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        const char* lb = p; p = nl ? nl + 1 : end;
        while (le > lb && (le[-1] == '\r' || le[-1] == '\n')) le--;
        if (le == lb) continue;
        size_t len = (size_t)(le - lb);
        if (len == 5 && !memcmp(lb, "CODE:", 5)) { mode = 1; continue; }
        if (len == 6 && !memcmp(lb, "NODES:", 6)) { mode = 2; continue; }
        if (len == 6 && !memcmp(lb, "EDGES:", 6)) { mode = 3; continue; }
        if (mode < 0) throw std::runtime_error("graph.txt: content before the first section header");
        Line L{lb, le};
        // the 7-bar escape
        for (const char* q = lb; q + 7 <= le; q++) {
            if (!memcmp(q, "|||||||", 7)) {
                std::string s(lb, le); s.replace((size_t)(q - lb), 7, "|||SLASH|||");
                if (patched.size() == patched.capacity()) throw std::runtime_error("graph.txt: too many '|' emission codes for this loader");
                patched.push_back(s); L.b = patched.back().data(); L.e = L.b + patched.back().size();
                break;
            }
        }
        (mode == 1 ? code : mode == 2 ? nodes : edges).push_back(L);
    }

    // ---- CODE: (locus, code) -> allele (LocusCodeAllocation::readFromVector, LocusCodeAllocation.cpp:287)
    struct CodeEnt { char sym[8]; uint8_t code[8]; int n = 0; };
    std::unordered_map<uint64_t, CodeEnt> codes; codes.reserve(code.size());
    std::vector<std::pair<const char*, const char*>> f;
    for (const Line& L : code) {
        split_fields(L.b, L.e, f);
        if (f.size() != 3) throw std::runtime_error("Cannot read CODE from line, expect 3 fields! Line: " + std::string(L.b, L.e));
        long long c = to_ll(f[2]);
        if (c < 0 || c > 250) throw std::runtime_error("Weird codedChar value: cannot convert back! " + std::string(f[2].first, f[2].second));
        if (f[1].second - f[1].first != 1) throw std::runtime_error("graph.txt: only single-character alleles are supported on the alignment path (Graph.cpp:2517-2519)");
        CodeEnt& ce = codes[fnv(f[0].first, f[0].second)];
        if (ce.n >= 8) throw std::runtime_error("graph.txt: more than 8 alleles at one locus");
        ce.sym[ce.n] = *f[1].first; ce.code[ce.n] = (uint8_t)c; ce.n++;
    }

    // ---- NODES (ordinal = order of appearance)
    const int32_t nn = (int32_t)nodes.size();
    std::vector<int32_t> ord_level(nn);
    std::unordered_map<long long, int32_t> idx2ord; idx2ord.reserve((size_t)nn * 2);
    int32_t max_level = -1;
    for (int32_t i = 0; i < nn; i++) {
        split_fields(nodes[i].b, nodes[i].e, f);
        if (f.size() != 3) throw std::runtime_error("Cannot node-parse this line (expext 6 fields): " + std::string(nodes[i].b, nodes[i].e));
        long long idx = to_ll(f[0]), lev = to_ll(f[1]);
        if (lev < 0) throw std::runtime_error("graph.txt: negative node level");
        ord_level[i] = (int32_t)lev; idx2ord[idx] = i; max_level = std::max(max_level, (int32_t)lev);
    }
    g.n_levels = max_level + 1; g.n_nodes = nn;
    // flat node order = (level, ordinal): counting sort keeps ordinal order inside a level
    g.level_node_off.assign((size_t)g.n_levels + 1, 0);
    for (int32_t i = 0; i < nn; i++) g.level_node_off[(size_t)ord_level[i] + 1]++;
    for (int32_t l = 0; l < g.n_levels; l++) g.level_node_off[l + 1] += g.level_node_off[l];
    std::vector<int32_t> ord2flat(nn); g.node_ord.resize(nn); g.node_level.resize(nn);
    {
        std::vector<int32_t> cur(g.level_node_off.begin(), g.level_node_off.end() - 1);
        for (int32_t i = 0; i < nn; i++) { int32_t fl = cur[ord_level[i]]++; ord2flat[i] = fl; g.node_ord[fl] = i; g.node_level[fl] = ord_level[i]; }
    }
    for (int32_t l = 0; l < g.n_levels; l++) {
        int32_t w = g.level_node_off[l + 1] - g.level_node_off[l];
        if (w == 0) throw std::runtime_error("graph.txt: a level without nodes");
        g.max_nodes_per_level = std::max(g.max_nodes_per_level, w);
    }

    // ---- EDGES
    const int32_t ne = (int32_t)edges.size();
    g.n_edges = ne;
    std::vector<int32_t> o_from(ne), o_to(ne); std::vector<uint8_t> o_em(ne); std::vector<uint64_t> o_loc(ne);
    std::vector<std::pair<const char*, const char*>> o_locname(ne);
    for (int32_t i = 0; i < ne; i++) {
        split_fields(edges[i].b, edges[i].e, f);
        if (f.size() != 6 && f.size() != 8) throw std::runtime_error("Cannot edge-parse this line (expext 6/8 fields): " + std::string(edges[i].b, edges[i].e));
        std::pair<const char*, const char*> cf = f[3];
        uint8_t codeChar;
        if (cf.second - cf.first == 5 && !memcmp(cf.first, "SLASH", 5)) codeChar = (uint8_t)'|';
        else if (cf.second - cf.first == 1) codeChar = (uint8_t)*cf.first;
        else throw std::runtime_error("Cannot cast to unsigned char: " + std::string(cf.first, cf.second));
        auto itf = idx2ord.find(to_ll(f[4])), itt = idx2ord.find(to_ll(f[5]));
        if (itf == idx2ord.end() || itt == idx2ord.end()) throw std::runtime_error("graph.txt: edge refers to an unknown node: " + std::string(edges[i].b, edges[i].e));
        o_from[i] = itf->second; o_to[i] = itt->second;
        uint64_t lh = fnv(f[1].first, f[1].second); o_loc[i] = lh; o_locname[i] = f[1];
        auto ci = codes.find(lh);
        if (ci == codes.end()) throw std::runtime_error("graph.txt: no CODE entry for locus " + std::string(f[1].first, f[1].second));
        int hit = -1; for (int k = 0; k < ci->second.n; k++) if (ci->second.code[k] == codeChar) hit = k;
        if (hit < 0) throw std::runtime_error("graph.txt: cannot decode emission of edge line " + std::string(edges[i].b, edges[i].e));
        o_em[i] = (uint8_t)ci->second.sym[hit];
        if (ord_level[o_to[i]] != ord_level[o_from[i]] + 1) throw std::runtime_error("graph.txt: edge does not join level l to l+1");
    }
    g.level_edge_off.assign((size_t)g.n_levels + 1, 0);
    for (int32_t i = 0; i < ne; i++) g.level_edge_off[(size_t)ord_level[o_from[i]] + 1]++;
    for (int32_t l = 0; l < g.n_levels; l++) g.level_edge_off[l + 1] += g.level_edge_off[l];
    g.edge_from.resize(ne); g.edge_to.resize(ne); g.edge_emis.resize(ne); g.edge_ord.resize(ne); g.ord_to_edge.resize(ne);
    g.level_names.assign((size_t)std::max(0, g.n_levels - 1), std::string());
    {
        std::vector<int32_t> cur(g.level_edge_off.begin(), g.level_edge_off.end() - 1);
        for (int32_t i = 0; i < ne; i++) {
            int32_t l = ord_level[o_from[i]]; int32_t fe = cur[l]++;
            g.edge_from[fe] = ord2flat[o_from[i]]; g.edge_to[fe] = ord2flat[o_to[i]]; g.edge_emis[fe] = o_em[i]; g.edge_ord[fe] = i; g.ord_to_edge[i] = fe;
        }
    }
    for (int32_t l = 0; l + 1 < g.n_levels; l++) g.max_edges_per_level = std::max(g.max_edges_per_level, g.level_edge_off[l + 1] - g.level_edge_off[l]);
    // adjacency (flat edge order inside a level is ordinal order, so pushing in flat order keeps set<Edge*> order)
    g.node_out_off.assign((size_t)nn + 1, 0); g.node_in_off.assign((size_t)nn + 1, 0);
    for (int32_t e = 0; e < ne; e++) { g.node_out_off[(size_t)g.edge_from[e] + 1]++; g.node_in_off[(size_t)g.edge_to[e] + 1]++; }
    for (int32_t i = 0; i < nn; i++) { g.node_out_off[i + 1] += g.node_out_off[i]; g.node_in_off[i + 1] += g.node_in_off[i]; }
    g.node_out.resize(ne); g.node_in.resize(ne);
    {
        std::vector<int32_t> co(g.node_out_off.begin(), g.node_out_off.end() - 1), ci(g.node_in_off.begin(), g.node_in_off.end() - 1);
        for (int32_t e = 0; e < ne; e++) { g.node_out[co[g.edge_from[e]]++] = e; g.node_in[ci[g.edge_to[e]]++] = e; }
    }
    // Graph::getOneLocusIDforLevel (Graph.cpp:1253): locus of the first outgoing edge of the first node of the level
    for (int32_t l = 0; l + 1 < g.n_levels; l++) {
        int32_t n0 = g.level_node_off[l];
        if (g.node_out_off[n0 + 1] == g.node_out_off[n0]) throw std::runtime_error("graph.txt: node without outgoing edges before the last level");
        int32_t e0 = g.node_out[g.node_out_off[n0]];
        const auto& nm = o_locname[g.edge_ord[e0]];
        g.level_names[l].assign(nm.first, nm.second);
    }

    compute_gap_paths(g);
    compute_gap_stretches(g);
    load_contigs(dir, g);
    if (!getenv("HLALA_NO_GRAPH_CACHE")) save_cache(cache_path, g, inputs_signature(dir));
}

// Graph::computeGapEdgePaths (Graph.cpp:347-476), containers keyed by canonical ordinals instead of pointers.
static void compute_gap_paths(FlatGraph& g) {
    typedef std::map<int32_t, std::vector<int32_t>> ByFrom;     // from-node ordinal -> edge path (flat edges)
    std::map<int32_t, ByFrom> running;                          // current-node ordinal -> paths ending there
    std::vector<int32_t> flat_of_ord(g.n_nodes);
    for (int32_t i = 0; i < g.n_nodes; i++) flat_of_ord[g.node_ord[i]] = i;
    g.path_off.assign(1, 0);
    for (int32_t l = 0; l < g.n_levels; l++) {
        std::map<int32_t, ByFrom> next;
        std::set<int32_t> seen_gap_edge;
        for (auto& cur : running) {
            int32_t node = flat_of_ord[cur.first];
            int non_gap = 0;
            for (int32_t k = g.node_out_off[node]; k < g.node_out_off[node + 1]; k++) {
                int32_t e = g.node_out[k];
                if (g.edge_emis[e] == '_') {
                    seen_gap_edge.insert(e);
                    int32_t tgt = g.node_ord[g.edge_to[e]];
                    for (auto& fp : cur.second) {
                        auto it = next.find(tgt);
                        if (it == next.end() || it->second.count(fp.first) == 0) { std::vector<int32_t> path = fp.second; path.push_back(e); next[tgt][fp.first] = std::move(path); }
                    }
                } else non_gap++;
            }
            if (non_gap != 0 || l == g.n_levels - 1) {
                for (auto& fp : cur.second) {
                    g.path_edges.insert(g.path_edges.end(), fp.second.begin(), fp.second.end());
                    g.path_off.push_back((int32_t)g.path_edges.size());
                    g.path_from.push_back(flat_of_ord[fp.first]); g.path_to.push_back(node);
                }
            }
        }
        if (l + 1 < g.n_levels) {
            for (int32_t e = g.level_edge_off[l]; e < g.level_edge_off[l + 1]; e++) {   // getEdgesEmanatingFromLevel: set<Edge*> order
                if (g.edge_emis[e] != '_' || seen_gap_edge.count(e)) continue;
                int32_t from = g.node_ord[g.edge_from[e]], tgt = g.node_ord[g.edge_to[e]];
                auto it = next.find(tgt);
                if (it == next.end() || it->second.count(from) == 0) next[tgt][from] = std::vector<int32_t>(1, e);
            }
        }
        running.swap(next);
    }
    g.n_paths = (int32_t)g.path_from.size();
    // jump lists: gapEdgePaths_connectedNodes_forwards[first][last] / _backwards[last][first]; inner maps iterate by target pointer
    auto build = [&](const std::vector<int32_t>& key, const std::vector<int32_t>& tgt, std::vector<int32_t>& off, std::vector<int32_t>& lst) {
        std::vector<std::vector<std::pair<int32_t, int32_t>>> per((size_t)g.n_nodes);
        for (int32_t pI = 0; pI < g.n_paths; pI++) per[key[pI]].push_back({g.node_ord[tgt[pI]], pI});
        off.assign((size_t)g.n_nodes + 1, 0); lst.clear();
        for (int32_t n = 0; n < g.n_nodes; n++) {
            std::sort(per[n].begin(), per[n].end());
            for (size_t k = 1; k < per[n].size(); k++) if (per[n][k].first == per[n][k - 1].first) throw std::runtime_error("gap paths: two paths between the same pair of nodes (Graph.cpp:461-462 asserts)");
            for (auto& x : per[n]) lst.push_back(x.second);
            off[n + 1] = (int32_t)lst.size();
        }
    };
    build(g.path_from, g.path_to, g.jump_fwd_off, g.jump_fwd_path);
    build(g.path_to, g.path_from, g.jump_bwd_off, g.jump_bwd_path);
}

// processBAM ctor, processBAM.cpp:91-149: runs of >= 3 consecutive levels that each have at least one '_' edge
static void compute_gap_stretches(FlatGraph& g) {
    int32_t L = g.n_levels - 1;
    g.gap_stretch.assign((size_t)std::max(0, L), 0);
    int32_t start = -1;
    auto add = [&](int32_t a, int32_t b) { if (b - a + 1 >= 3) for (int32_t i = a; i <= b; i++) g.gap_stretch[i] = 1; };
    for (int32_t l = 0; l < L; l++) {
        bool have = false;
        for (int32_t e = g.level_edge_off[l]; e < g.level_edge_off[l + 1]; e++) if (g.edge_emis[e] == '_') { have = true; break; }
        if (have) { if (start == -1) start = l; }
        else if (start != -1) { add(start, l - 1); start = -1; }
    }
    if (start != -1) add(start, g.n_levels - 2);
}

static void load_contigs(const std::string& dir, FlatGraph& g) {
    std::ifstream s(dir + "/sequences.txt");
    if (!s.is_open()) throw std::runtime_error("Cannot open " + dir + "/sequences.txt");
    std::string line; std::getline(s, line);
    std::vector<std::string> header; { std::string cur; for (char c : line) { if (c == '\r' || c == '\n') continue; if (c == '\t') { header.push_back(cur); cur.clear(); } else cur.push_back(c); } header.push_back(cur); }
    int c_id = -1, c_chr = -1;
    for (size_t i = 0; i < header.size(); i++) { if (header[i] == "SequenceID") c_id = (int)i; if (header[i] == "Chr") c_chr = (int)i; }
    if (c_id < 0 || c_chr < 0) throw std::runtime_error("sequences.txt: missing SequenceID/Chr columns");
    // FASTA: id = text up to the first space (Utilities.cpp:757-808)
    std::map<std::string, std::string> fa;
    {
        std::string txt = read_file(dir + "/mapping_PRGonly/referenceGenome.fa");
        const char* p = txt.data(); const char* end = p + txt.size(); std::string* cur = nullptr;
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p)); const char* le = nl ? nl : end;
            const char* lb = p; p = nl ? nl + 1 : end;
            while (le > lb && (le[-1] == '\r')) le--;
            if (le == lb) continue;
            if (*lb == '>') { std::string id(lb + 1, le); size_t sp = id.find(' '); if (sp != std::string::npos) id = id.substr(0, sp); cur = &fa[id]; cur->clear(); }
            else if (cur) cur->append(lb, le);
        }
    }
    g.contig_off.assign(1, 0);
    struct Tr { int32_t id; std::vector<int32_t> lv; };
    std::vector<Tr> trs;
    while (std::getline(s, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> f; { std::string cur; for (char c : line) { if (c == '\t') { f.push_back(cur); cur.clear(); } else cur.push_back(c); } f.push_back(cur); }
        if (f.size() != header.size()) throw std::runtime_error("sequences.txt: field count mismatch");
        int32_t id = atoi(f[c_id].c_str());
        std::string bam = f[c_chr].empty() ? ("PRG_" + f[c_id]) : f[c_chr];
        // processBAM.cpp:86-88: the PRG-only FASTA must NOT hold PRG_5 (assert) and the reference injects "N" for it
        auto it = fa.find(bam);
        if (bam == "PRG_5") { if (it != fa.end()) throw std::runtime_error("PRG_5 must not be part of " + dir + "/mapping_PRGonly/referenceGenome.fa (the reference asserts this and uses \"N\")"); }
        else if (it == fa.end()) throw std::runtime_error(bam + " cannot be found in the PRG-only reference genome " + dir + "/mapping_PRGonly/referenceGenome.fa");
        const std::string seq = (bam == "PRG_5") ? std::string("N") : it->second;
        // translation file with the reference's getline loop semantics (processBAM.cpp:4409-4414): every line is parsed with
        // StrtoI, and a file ending in a newline yields one more (empty) line that parses as 0
        std::string tp = dir + "/translation/" + f[c_id] + ".txt";
        std::string tt;
        try { tt = read_file(tp); } catch (...) { throw std::runtime_error("Expected coordinate translation file not found: " + tp); }
        Tr tr; tr.id = id; tr.lv.reserve(seq.size() + 1);
        { const char* p = tt.data(); const char* end = p + tt.size();
          while (true) {
              const char* nl = (p < end) ? (const char*)memchr(p, '\n', (size_t)(end - p)) : nullptr; const char* le = nl ? nl : end;
              long long v = 0; bool neg = false; const char* q = p; if (q < le && *q == '-') { neg = true; q++; }
              for (; q < le && *q >= '0' && *q <= '9'; q++) v = v * 10 + (*q - '0');
              tr.lv.push_back((int32_t)(neg ? -v : v));
              if (!nl) break;
              p = nl + 1;
          } }
        g.contig_prg_id.push_back(id); g.contig_bam_name.push_back(bam);
        g.contig_seq.insert(g.contig_seq.end(), seq.begin(), seq.end());
        size_t n = seq.size();
        if (tr.lv.size() < n) throw std::runtime_error("translation shorter than contig " + bam);
        g.contig_level.insert(g.contig_level.end(), tr.lv.begin(), tr.lv.begin() + (long)n);
        g.contig_tr_len.push_back((int32_t)tr.lv.size());
        g.contig_off.push_back((int64_t)g.contig_seq.size());
        for (int32_t lv : tr.lv) if (lv < 0 || lv >= g.n_levels) throw std::runtime_error("translation level out of range for contig " + bam);
        trs.push_back(std::move(tr));
    }
    g.n_contigs = (int32_t)g.contig_prg_id.size();
    // graphLevel_2_underlyingSequencePositions (processBAM.cpp:4441-4456): per level a std::map<PRG id, position>; a later
    // position of the same contig overwrites an earlier one. Built as a CSR: contigs in ascending id order, counting sort by level.
    std::sort(trs.begin(), trs.end(), [](const Tr& a, const Tr& b) { return a.id < b.id; });
    for (size_t i = 1; i < trs.size(); i++) if (trs[i].id == trs[i - 1].id) throw std::runtime_error("sequences.txt: duplicate SequenceID");
    g.anchor_off.assign((size_t)g.n_levels + 1, 0);
    std::vector<int32_t> lastpos;   // per contig: dedupe repeated levels (keep the last position)
    for (const Tr& t : trs) {
        // a level can repeat inside one contig only through the bogus trailing 0; count distinct levels
        std::vector<int32_t> seen_level0;   // handled by the overwrite pass below
        for (size_t pos = 0; pos < t.lv.size(); pos++) {
            bool repeat = false;
            if (pos + 1 == t.lv.size() && t.lv[pos] == 0) { for (size_t q = 0; q < pos && !repeat; q++) { if (t.lv[q] == 0) repeat = true; if (t.lv[q] > 0) break; } }
            if (!repeat) g.anchor_off[(size_t)t.lv[pos] + 1]++;
        }
    }
    for (int32_t l = 0; l < g.n_levels; l++) g.anchor_off[l + 1] += g.anchor_off[l];
    g.anchor_prg_id.assign((size_t)g.anchor_off[g.n_levels], 0); g.anchor_pos.assign((size_t)g.anchor_off[g.n_levels], 0);
    {
        std::vector<int32_t> cur(g.anchor_off.begin(), g.anchor_off.end() - 1);
        for (const Tr& t : trs) {
            int32_t slot0 = -1;   // slot of this contig at level 0, if any
            for (size_t pos = 0; pos < t.lv.size(); pos++) {
                int32_t lv = t.lv[pos];
                if (lv == 0 && slot0 >= 0) { g.anchor_pos[slot0] = (int32_t)pos; continue; }     // overwrite (last position wins)
                int32_t k = cur[lv]++;
                if (k >= g.anchor_off[lv + 1]) throw std::runtime_error("translation is not strictly increasing for a contig (levels repeat)");
                g.anchor_prg_id[k] = t.id; g.anchor_pos[k] = (int32_t)pos;
                if (lv == 0) slot0 = k;
            }
        }
        for (int32_t l = 0; l < g.n_levels; l++) if (cur[l] != g.anchor_off[l + 1]) throw std::runtime_error("translation is not strictly increasing for a contig (levels repeat)");
    }
}

} // namespace hlala
