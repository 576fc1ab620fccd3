// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library. The product path
// (hla-la_b200/) never links, imports or executes anything under oracle/.
//
// A plain, single-threaded C++ restatement of the reference's read-to-PRG alignment path, written for clarity with
// STL containers (maps keyed by canonical ordinals where the reference keys by pointer). Each function cites the
// reference lines it follows. It shares no code with the CUDA implementation.
//
// Pinning: tests/test_oracle_vs_reference.py compares every output of this file with the UNMODIFIED reference
// translation units compiled into oracle/_ref/libhlala_ref.so (graph arrays, gap paths, per-chain columns and
// log-likelihoods bit for bit, per-pair columns, mapping qualities and per-column phred characters), and
// tests/golden/ holds fixtures generated from that compiled reference (tests/golden/make_golden.py) for boxes where
// the reference tree is absent. The reference has no golden vectors of its own for this path.
//
// Canonical order: nodes and edges are numbered in order of appearance in graph.txt; the reference's pointer-ordered
// std::set iteration equals that order (see oracle/ref_driver.cpp for how that is enforced when pinning).
// The reference's single random draw (extensionAligner.cpp:1459) is restated as "first candidate", matching the
// pinned reference build.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace {

struct OracleError : std::runtime_error { explicit OracleError(const std::string& m) : std::runtime_error(m) {} };
#define REQUIRE(cond, msg) do { if (!(cond)) throw OracleError(std::string("reference assertion would fail: ") + msg); } while (0)

std::vector<std::string> split_bars(const std::string& s) {
    std::vector<std::string> out; size_t p = 0;
    for (;;) { size_t h = s.find("|||", p); if (h == std::string::npos) { out.push_back(s.substr(p)); break; } out.push_back(s.substr(p, h - p)); p = h + 3; }
    return out;
}
void chomp(std::string& s) { while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); }

// ------------------------------------------------------------------------------------------------ graph
struct OEdge { int from, to; char emis; };
struct OGraph {
    int n_levels = 0;
    std::vector<int> node_level;                       // by node ordinal
    std::vector<OEdge> edges;                          // by edge ordinal
    std::vector<std::vector<int>> out, in;             // per node: edge ordinals ascending (= std::set<Edge*> order)
    std::vector<std::vector<int>> level_nodes;         // per level: node ordinals ascending (= z order, alignerBase.cpp:27-37)
    std::vector<int> node_z;
    std::vector<std::vector<int>> gap_paths;           // completedGapEdgePaths (Graph.cpp:347-476)
    std::vector<std::map<int, int>> jump_fwd, jump_bwd;   // node -> (other end node -> path id), iterated by node ordinal
    std::vector<char> gap_stretch;                     // processBAM.cpp:91-149
    // contigs
    std::vector<int> contig_id; std::vector<std::string> contig_seq; std::vector<std::vector<int>> contig_tr;
    std::vector<std::map<int, int>> anchors;           // graphLevel_2_underlyingSequencePositions (processBAM.cpp:4441-4456)
};

// Graph::readFromFile (Graph.cpp:2329-2559) + LocusCodeAllocation::readFromVector (LocusCodeAllocation.cpp:287)
void read_graph(const std::string& path, OGraph& g) {
    std::ifstream f(path); if (!f.is_open()) throw OracleError("Cannot open graph file: " + path);
    std::string line; int mode = 0;
    std::map<std::string, std::map<int, std::string>> code;
    std::map<long long, int> node_ord;
    std::vector<std::string> edge_lines;
    while (std::getline(f, line)) {
        chomp(line); if (line.empty()) continue;
        size_t bad = line.find("|||||||"); if (bad != std::string::npos) line.replace(bad, 7, "|||SLASH|||");
        if (line == "CODE:") { mode = 1; continue; } if (line == "NODES:") { mode = 2; continue; } if (line == "EDGES:") { mode = 3; continue; }
        REQUIRE(mode > 0, "graph.txt section header");
        std::vector<std::string> fld = split_bars(line);
        if (mode == 1) { REQUIRE(fld.size() == 3, "CODE line has 3 fields"); code[fld[0]][atoi(fld[2].c_str())] = fld[1]; }
        else if (mode == 2) { REQUIRE(fld.size() == 3, "NODES line has 3 fields"); node_ord[atoll(fld[0].c_str())] = (int)g.node_level.size(); g.node_level.push_back(atoi(fld[1].c_str())); }
        else edge_lines.push_back(line);
    }
    int maxl = -1; for (int l : g.node_level) maxl = std::max(maxl, l);
    g.n_levels = maxl + 1;
    g.out.resize(g.node_level.size()); g.in.resize(g.node_level.size()); g.level_nodes.resize(g.n_levels); g.node_z.resize(g.node_level.size());
    for (size_t n = 0; n < g.node_level.size(); n++) { g.node_z[n] = (int)g.level_nodes[g.node_level[n]].size(); g.level_nodes[g.node_level[n]].push_back((int)n); }
    for (const std::string& l : edge_lines) {
        std::vector<std::string> fld = split_bars(l);
        REQUIRE(fld.size() == 6 || fld.size() == 8, "EDGES line has 6 or 8 fields");
        std::string c = fld[3]; if (c == "SLASH") c = "|";
        REQUIRE(c.size() == 1, "edge code is one byte");
        const std::string& allele = code.at(fld[1]).at((unsigned char)c[0]);
        REQUIRE(allele.size() == 1, "emission is one character (Graph.cpp:2517-2519)");
        OEdge e; e.from = node_ord.at(atoll(fld[4].c_str())); e.to = node_ord.at(atoll(fld[5].c_str())); e.emis = allele[0];
        int id = (int)g.edges.size(); g.edges.push_back(e); g.out[e.from].push_back(id); g.in[e.to].push_back(id);
    }
}

// Graph::computeGapEdgePaths (Graph.cpp:347-476)
void gap_paths(OGraph& g) {
    typedef std::map<int, std::vector<int>> ByStart;
    std::map<int, ByStart> running;
    for (int l = 0; l < g.n_levels; l++) {
        std::map<int, ByStart> next; std::set<int> seen;
        for (auto& at : running) {
            int node = at.first; int plain = 0;
            for (int e : g.out[node]) {
                if (g.edges[e].emis != '_') { plain++; continue; }
                seen.insert(e);
                for (auto& p : at.second) if (!next.count(g.edges[e].to) || !next[g.edges[e].to].count(p.first)) { std::vector<int> q = p.second; q.push_back(e); next[g.edges[e].to][p.first] = q; }
            }
            if (plain || l == g.n_levels - 1) for (auto& p : at.second) g.gap_paths.push_back(p.second);
        }
        std::vector<int> level_edges;
        for (int n : g.level_nodes[l]) for (int e : g.out[n]) level_edges.push_back(e);
        std::sort(level_edges.begin(), level_edges.end());     // getEdgesEmanatingFromLevel returns a std::set<Edge*>
        for (int e : level_edges) {
            if (g.edges[e].emis != '_' || seen.count(e)) continue;
            int a = g.edges[e].from, b = g.edges[e].to;
            if (!next.count(b) || !next[b].count(a)) next[b][a] = std::vector<int>(1, e);
        }
        running.swap(next);
    }
    g.jump_fwd.resize(g.node_level.size()); g.jump_bwd.resize(g.node_level.size());
    for (size_t p = 0; p < g.gap_paths.size(); p++) {
        int a = g.edges[g.gap_paths[p].front()].from, b = g.edges[g.gap_paths[p].back()].to;
        REQUIRE(!g.jump_fwd[a].count(b), "one gap path per node pair (Graph.cpp:461)");
        g.jump_fwd[a][b] = (int)p; g.jump_bwd[b][a] = (int)p;
    }
}

void gap_stretches(OGraph& g) {   // processBAM.cpp:91-149
    g.gap_stretch.assign(g.n_levels - 1, 0);
    int start = -1;
    auto close = [&](int a, int b) { if (b - a + 1 >= 3) for (int i = a; i <= b; i++) g.gap_stretch[i] = 1; };
    for (int l = 0; l < g.n_levels - 1; l++) {
        bool has = false; for (int n : g.level_nodes[l]) for (int e : g.out[n]) if (g.edges[e].emis == '_') has = true;
        if (has) { if (start < 0) start = l; } else if (start >= 0) { close(start, l - 1); start = -1; }
    }
    if (start >= 0) close(start, g.n_levels - 2);
}

void read_contigs(const std::string& dir, OGraph& g) {   // processBAM.cpp:85-88, 1216-1320, 4389-4459
    std::map<std::string, std::string> fa;
    { std::ifstream f(dir + "/mapping_PRGonly/referenceGenome.fa"); std::string l, id; while (std::getline(f, l)) { chomp(l); if (l.empty()) continue; if (l[0] == '>') { id = l.substr(1); size_t sp = id.find(' '); if (sp != std::string::npos) id = id.substr(0, sp); fa[id] = ""; } else fa[id] += l; } }
    std::ifstream s(dir + "/sequences.txt"); if (!s.is_open()) throw OracleError("cannot open sequences.txt");
    std::string line; std::getline(s, line);
    g.anchors.assign(g.n_levels, std::map<int, int>());
    while (std::getline(s, line)) {
        chomp(line); if (line.empty()) continue;
        std::vector<std::string> f; { std::stringstream ss(line); std::string c; while (std::getline(ss, c, '\t')) f.push_back(c); while (f.size() < 6) f.push_back(""); }
        int id = atoi(f[0].c_str()); std::string name = f[3].empty() ? "PRG_" + f[0] : f[3];
        g.contig_id.push_back(id); g.contig_seq.push_back(name == "PRG_5" ? std::string("N") : fa.at(name));
        std::ifstream t(dir + "/translation/" + f[0] + ".txt"); if (!t.is_open()) throw OracleError("Expected coordinate translation file not found");
        std::vector<int> tr; std::string tl;
        while (t.good()) { std::getline(t, tl); chomp(tl); tr.push_back(atoi(tl.c_str())); }   // a trailing newline yields a bogus 0 (Utilities.cpp:644-650)
        for (size_t p = 0; p < tr.size(); p++) { REQUIRE(tr[p] >= 0 && tr[p] < g.n_levels, "translation level in range"); g.anchors[tr[p]][id] = (int)p; }
        g.contig_tr.push_back(tr);
    }
}

// ------------------------------------------------------------------------------------------------ alignment records
struct Chain {   // verboseSeedChain (mapper/reads/verboseSeedChain.h:22-60)
    std::vector<int> level, edge; std::string gchar, schar; std::vector<char> from_seed; std::string mapq;
    int seq_begin = 0, seq_end = 0; bool reverse = false; double chain_mapq = -1;
    int first_level() const { for (int l : level) if (l != -1) return l; return -1; }
    int last_level() const { for (size_t i = level.size(); i-- > 0;) if (level[i] != -1) return level[i]; return -1; }
};

struct Batch {
    long long n_reads; const int64_t* read_off; const uint8_t* bases; const uint8_t* quals; const int32_t* chain_off; const int32_t* chain_contig; const int32_t* chain_pos;
    const uint16_t* chain_flag; const int32_t* chain_as; const int32_t* cigar_off; const uint32_t* cigar;
};

struct Oracle {
    OGraph g;

    // transformBAMreadToInternalAlignment (processBAM.cpp:4794-5337)
    void bam_to_columns(const Batch& b, int c, const std::string& seq, std::vector<int>& lv, std::string& gc, std::string& sc, int& start_raw, int& stop_raw) const {
        const std::string& ref = g.contig_seq.at(b.chain_contig[c]); const std::vector<int>& tr = g.contig_tr.at(b.chain_contig[c]);
        int gpos = b.chain_pos[c], ridx = 0; start_raw = -1; stop_raw = -1;
        int n_ops = b.cigar_off[c + 1] - b.cigar_off[c];
        REQUIRE(n_ops >= 1, "CIGAR not empty");
        for (int k = 0; k < n_ops; k++) {
            uint32_t cg = b.cigar[b.cigar_off[c] + k]; char op = "MIDNSHP=X"[cg & 15]; int len = (int)(cg >> 4);
            switch (op) {
            case 'M': case '=': case 'X': case 'D':
                for (int i = 0; i < len; i++) {
                    REQUIRE(gpos >= 0 && gpos < (int)ref.size() && gpos < (int)tr.size(), "reference position in range");
                    lv.push_back(tr[gpos]); gc.push_back(ref[gpos]);
                    if (op == 'D') sc.push_back('_'); else { REQUIRE(ridx < (int)seq.size(), "read index in range"); sc.push_back(seq[ridx]); }
                    if (start_raw < 0) start_raw = ridx;
                    gpos++; if (op != 'D') ridx++;
                    stop_raw = ridx - 1;
                }
                break;
            case 'I':
                for (int i = 0; i < len; i++) { REQUIRE(ridx < (int)seq.size(), "read index in range"); lv.push_back(-1); gc.push_back('_'); sc.push_back(seq[ridx]); if (start_raw < 0) start_raw = ridx; ridx++; stop_raw = ridx - 1; }
                break;
            case 'S': ridx += len; break;
            case 'H': if (k == 0) ridx += len; break;     // leading hard clip: offset into the primary's SEQ (processBAM.cpp:4868-4874)
            case 'P': break;
            default: throw OracleError("N character in CIGAR - should only be the case for RNASeq data!");
            }
        }
        REQUIRE(start_raw < stop_raw, "sequence_aligned_startInRaw < stopInRaw (processBAM.cpp:5245)");
    }

    // cleanInitialAlignment (processBAM.cpp:4621-4792)
    static void clean(std::vector<int>& lv, std::string& gc, std::string& sc) {
        bool in_run = false, changed = false; int run_start = -1, balance = 0;
        for (size_t p = 0; p < lv.size(); p++) {
            bool ins = lv[p] == -1, dgap = gc[p] == '_' && sc[p] == '_';
            if (ins || dgap) { if (!in_run) { in_run = true; run_start = (int)p; } if (ins) balance++; if (dgap) balance--; continue; }
            if (!in_run) continue;
            int run_stop = (int)p - 1;
            if (balance == 0) {
                std::string chars; std::vector<int> levels;
                for (int q = run_start; q <= run_stop; q++) { if (lv[q] == -1) chars.push_back(sc[q]); else levels.push_back(lv[q]); }
                int half = (run_stop - run_start + 1) / 2;
                for (int q = run_start; q <= run_stop; q++) { int k = q - run_start; if (k < half) { lv[q] = levels[k]; gc[q] = '_'; sc[q] = chars[k]; } else { lv[q] = -1; gc[q] = '_'; sc[q] = '_'; } }
                changed = true;
            }
            in_run = false; balance = 0;
        }
        if (!changed) return;
        std::vector<int> l2; std::string g2, s2;
        for (size_t p = 0; p < lv.size(); p++) if (!(lv[p] == -1 && gc[p] == '_' && sc[p] == '_')) { l2.push_back(lv[p]); g2.push_back(gc[p]); s2.push_back(sc[p]); }
        lv = l2; gc = g2; sc = s2;
    }

    // restrictInitialAlignmentToNoGapAreas (processBAM.cpp:4461-4619)
    void restrict_nogap(std::vector<int>& lv, std::string& gc, std::string& sc, int& start_raw, int& stop_raw) const {
        std::vector<std::pair<int, int>> raw; int run = -1, total_chars = 0;
        for (size_t i = 0; i < lv.size(); i++) {
            if (sc[i] != '_') total_chars++;
            if (lv[i] != -1 && g.gap_stretch.at(lv[i])) { if (run != -1) { raw.push_back({run, (int)i - 1}); run = -1; } }
            else if (run == -1) run = (int)i;
        }
        if (run != -1 && run != 0) raw.push_back({run, (int)lv.size() - 1});
        std::vector<std::pair<int, int>> cand;
        for (auto st : raw) {
            while (lv[st.first] == -1) { st.first++; if (st.first > st.second || st.first > (int)lv.size() - 1) break; }
            if (st.first <= st.second) while (lv[st.second] == -1) { st.second--; if (st.second < st.first || st.second < 0) break; }
            if (st.second >= st.first) cand.push_back(st);
        }
        if (cand.empty()) return;
        std::sort(cand.begin(), cand.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return (a.second - a.first) < (b.second - b.first); });
        std::pair<int, int> sel = cand.back();
        int inside = 0, before = 0, after = 0;
        for (int i = 0; i < (int)lv.size(); i++) if (sc[i] != '_') { if (i < sel.first) before++; else if (i > sel.second) after++; else inside++; }
        if (!(((double)inside / (double)total_chars) > 0.3)) return;
        lv = std::vector<int>(lv.begin() + sel.first, lv.begin() + sel.second + 1); gc = gc.substr(sel.first, sel.second - sel.first + 1); sc = sc.substr(sel.first, sel.second - sel.first + 1);
        start_raw += before; stop_raw -= after;
    }

    // PRGContigAlignment2Seed (processBAM.cpp:2491-3017), sequence seed only (the one alignment2Chain returns, :3126)
    Chain project(const Batch& b, int c, const std::string& seq, bool reverse) const {
        std::vector<int> rl; std::string rg, rs; int start_raw, stop_raw;
        bam_to_columns(b, c, seq, rl, rg, rs, start_raw, stop_raw);
        size_t first = 0; while (first < rl.size() && rl[first] == -1) { first++; start_raw++; }
        size_t last = rl.size() - 1; while (last > 0 && rl[last] == -1) { last--; stop_raw--; }
        REQUIRE(first < last, "firstColumn < lastColumn (processBAM.cpp:2532)");
        std::vector<int> lv; std::string gc, sc; int prev = -1;
        for (size_t i = first; i <= last; i++) {
            if (rl[i] != -1) {
                if (prev != -1) { REQUIRE(prev + 1 <= rl[i], "levels increase"); for (int l = prev + 1; l < rl[i]; l++) { lv.push_back(l); gc.push_back('_'); sc.push_back('_'); } }
                prev = rl[i];
            }
            lv.push_back(rl[i]); gc.push_back(rg[i]); sc.push_back(rs[i]);
        }
        clean(lv, gc, sc);
        restrict_nogap(lv, gc, sc, start_raw, stop_raw);
        REQUIRE(lv.front() != -1 && lv.back() != -1, "alignment ends carry levels");
        // column-synchronous Viterbi over the level DAG (processBAM.cpp:2707-2835): per node best score and all arg-max incoming edges
        struct Cell { double s; std::set<int> taken; };
        std::vector<std::map<int, Cell>> col(lv.size() + 1);
        for (int n : g.level_nodes.at(lv[0])) col[0][n].s = 0;
        size_t last_real = 0;
        for (size_t ci = 1; ci <= lv.size(); ci++) {
            if (lv[ci - 1] == -1) continue;
            REQUIRE(!col[last_real].empty(), "non-empty column map (processBAM.cpp:2729)");
            bool is_match = sc[ci - 1] == gc[ci - 1];
            for (auto& nc : col[last_real]) {
                REQUIRE(g.node_level[nc.first] == lv[ci - 1], "level contiguity");
                for (int e : g.out[nc.first]) {
                    const OEdge& ed = g.edges[e];
                    if (is_match && ed.emis != sc[ci - 1]) continue;
                    double s = nc.second.s + (ed.emis == sc[ci - 1] ? 1 : 0);
                    auto it = col[ci].find(ed.to);
                    if (it == col[ci].end()) { col[ci][ed.to].s = s; col[ci][ed.to].taken.insert(e); }
                    else if (it->second.s == s) it->second.taken.insert(e);
                    else if (it->second.s < s) { it->second.s = s; it->second.taken.clear(); it->second.taken.insert(e); }
                }
            }
            last_real = ci;
        }
        REQUIRE(!col[lv.size()].empty(), "non-empty final column (processBAM.cpp:2850)");
        double best = -1; int node = -1;
        for (auto& nc : col[lv.size()]) if (node < 0 || nc.second.s > best) { best = nc.second.s; node = nc.first; }   // first maximum = lowest ordinal among the best
        Chain ch; ch.level.resize(lv.size()); ch.edge.resize(lv.size()); ch.gchar.resize(lv.size()); ch.schar = sc; ch.from_seed.assign(lv.size(), 1);
        for (size_t ci = lv.size(); ci >= 1; ci--) {
            if (lv[ci - 1] == -1) { ch.level[ci - 1] = -1; ch.edge[ci - 1] = -1; ch.gchar[ci - 1] = '_'; continue; }
            int e = *col[ci].at(node).taken.begin();     // first taken edge in set order (processBAM.cpp:2889-2916)
            ch.level[ci - 1] = lv[ci - 1]; ch.edge[ci - 1] = e; ch.gchar[ci - 1] = g.edges[e].emis; node = g.edges[e].from;
        }
        ch.seq_begin = start_raw; ch.seq_end = stop_raw; ch.reverse = reverse;
        return ch;
    }

    // extensionAligner::fullNeedleman_diagonal_extension_gapJumper (extensionAligner.cpp:335-1556). Returns the extension columns
    // (in alignment order) or an empty chain when the reference returns no extension.
    struct Step { int x = -1, y = -1, z = -1, mat = -1, edge = -1; bool jump = false; };
    struct Score { double D, GG, SG; Step bD, bGG, bSG; };
    typedef std::tuple<int, int, int> Key;
    bool extend(const std::string& seq, int start_seq, int start_level, int start_z, bool pos, Chain& ext) const {
        const double NEG = -std::numeric_limits<double>::max();
        const int max_level = g.n_levels - 1, max_seq = (int)seq.size(), dir = pos ? 1 : -1, end_seq = pos ? max_seq : 0;
        std::map<Key, Score> done;
        Score s0; s0.D = 0; s0.GG = s0.SG = NEG; done[Key(start_level, start_seq, start_z)] = s0;
        std::vector<Key> m1(1, Key(start_level, start_seq, start_z)), m2;
        double cur_max = 0; std::vector<Key> maxima(1, m1[0]); int last_inc = 0;
        std::set<std::string> complete;
        struct Cand { std::vector<double> D, GG, SG; std::vector<Step> bD, bGG, bSG; };
        auto first_max = [](const std::vector<double>& v) { size_t b = 0; for (size_t i = 1; i < v.size(); i++) if (v[i] > v[b]) b = i; return b; };
        for (int diag = 1;; diag++) {
            if (diag - last_inc > 40) break;
            if (m1.empty() && m2.empty()) break;
            std::map<Key, Cand> touched;
            for (const Key& k : m2) {
                int x = std::get<0>(k), y = std::get<1>(k), z = std::get<2>(k); int nx = x + dir, ny = y + dir;
                if (nx > max_level || ny > max_seq || nx < 0 || ny < 0) continue;
                char sc = pos ? seq[y] : seq[y - 1]; int node = g.level_nodes[x][z];
                for (int e : (pos ? g.out[node] : g.in[node])) {
                    int other = pos ? g.edges[e].to : g.edges[e].from; Step st; st.x = x; st.y = y; st.z = z; st.mat = 0; st.edge = e;
                    Cand& t = touched[Key(nx, ny, g.node_z[other])]; t.D.push_back(done.at(k).D + (g.edges[e].emis == sc ? 2 : -5)); t.bD.push_back(st);
                }
            }
            for (const Key& k : m1) {
                int x = std::get<0>(k), y = std::get<1>(k), z = std::get<2>(k); const Score& ps = done.at(k); int node = g.level_nodes[x][z];
                if (pos ? (x <= max_level && y + 1 <= max_seq) : (x >= 0 && y - 1 >= 0)) {
                    Cand& t = touched[Key(x, y + dir, z)]; Step a; a.x = x; a.y = y; a.z = z; a.mat = 0; Step bb = a; bb.mat = 1;
                    t.GG.push_back(ps.D - 4 - 2); t.bGG.push_back(a); t.GG.push_back(ps.GG == NEG ? NEG : ps.GG - 2); t.bGG.push_back(bb);
                }
                if (pos ? (x + 1 <= max_level && y <= max_seq) : (x - 1 >= 0 && y >= 0)) {
                    for (int e : (pos ? g.out[node] : g.in[node])) {
                        int other = pos ? g.edges[e].to : g.edges[e].from; bool gap = g.edges[e].emis == '_';
                        Cand& t = touched[Key(x + dir, y, g.node_z[other])]; Step a; a.x = x; a.y = y; a.z = z; a.edge = e; a.mat = 0; Step c2 = a; c2.mat = 2;
                        t.SG.push_back(gap ? NEG : ps.D - 4 - 2); t.bSG.push_back(a);
                        t.SG.push_back(ps.SG == NEG ? NEG : (gap ? ps.SG + 0 : ps.SG - 2)); t.bSG.push_back(c2);
                        if (gap) { t.D.push_back(ps.D + 0); t.bD.push_back(a); }
                    }
                }
                for (auto& j : (pos ? g.jump_fwd[node] : g.jump_bwd[node])) {
                    int len = (int)g.gap_paths[j.second].size(); int jx = x + dir * len;
                    if (!(pos ? (jx <= max_level && y <= max_seq) : (jx >= 0 && y >= 0))) continue;
                    Step a; a.x = x; a.y = y; a.z = z; a.mat = 0; a.edge = j.second; a.jump = true;
                    Cand& t = touched[Key(jx, y, g.node_z[j.first])]; t.D.push_back(ps.D + 0); t.bD.push_back(a);
                }
            }
            std::vector<Key> next;
            for (auto& tc : touched) {
                const Key& k = tc.first; Cand& t = tc.second; int x = std::get<0>(k), y = std::get<1>(k), z = std::get<2>(k);
                double sGG = NEG, sSG = NEG; Step bGG, bSG;
                if (!t.GG.empty()) { size_t i = first_max(t.GG); sGG = t.GG[i]; bGG = t.bGG[i]; }
                if (!t.SG.empty()) { size_t i = first_max(t.SG); sSG = t.SG[i]; bSG = t.bSG[i]; }
                Step self1; self1.x = x; self1.y = y; self1.z = z; self1.mat = 1; Step self2 = self1; self2.mat = 2;
                t.D.push_back(sGG); t.bD.push_back(self1); t.D.push_back(sSG); t.bD.push_back(self2);
                size_t iD = first_max(t.D); double sD = t.D[iD];
                if (!(sD >= -16)) continue;
                bool is_new = !done.count(k); bool over = false; Score& sc = done[k];
                if (is_new || sc.D < sD) { over = !is_new; sc.D = sD; sc.bD = t.bD[iD]; }
                if (is_new || sc.GG < sGG) { over = !is_new; sc.GG = sGG; sc.bGG = bGG; }
                if (is_new || sc.SG < sSG) { over = !is_new; sc.SG = sSG; sc.bSG = bSG; }
                if (y == end_seq) complete.insert(std::to_string(x) + "/" + std::to_string(z));
                next.push_back(k);
                Step st = sc.bD;
                while (st.x == x && st.y == y) st = st.mat == 1 ? done.at(Key(st.x, st.y, st.z)).bGG : done.at(Key(st.x, st.y, st.z)).bSG;
                const Score& src = done.at(Key(st.x, st.y, st.z));
                int prev = (int)(st.mat == 0 ? src.D : st.mat == 1 ? src.GG : src.SG);
                int diff = (int)(sD - prev);
                if (sD == cur_max) { if (diff != 0) { maxima.push_back(k); last_inc = diag; } }
                else if (sD > cur_max) { cur_max = sD; maxima.assign(1, k); last_inc = diag; }
                if (over) last_inc = diag;
            }
            if (!next.empty()) {
                double mx = done.at(next[0]).D; for (const Key& k : next) mx = std::max(mx, done.at(k).D);
                std::vector<Key> keep; for (const Key& k : next) if (mx - done.at(k).D <= 15) keep.push_back(k);
                next = keep;
            }
            m2 = m1; m1 = next;
        }
        Key endk; bool have = false;
        if (!complete.empty()) {
            double best = 0; std::string bestid;
            for (const std::string& id : complete) {     // std::set<std::string>: lexicographic order; the reference's draw among ties is pinned to the first
                int x = atoi(id.substr(0, id.find('/')).c_str()), z = atoi(id.substr(id.find('/') + 1).c_str());
                double s = done.at(Key(x, end_seq, z)).D;
                if (!have || s > best) { best = s; endk = Key(x, end_seq, z); have = true; }
            }
        } else if (cur_max > 0) { endk = maxima[0]; have = true; }
        if (!have) return false;
        Key k = endk; int mat = 0; std::vector<int> lv, ed; std::string gc, sc;
        while (!(std::get<0>(k) == start_level && std::get<1>(k) == start_seq)) {
            const Score& s = done.at(k); const Step& st = mat == 0 ? s.bD : mat == 1 ? s.bGG : s.bSG;
            int x = std::get<0>(k), y = std::get<1>(k);
            if (st.jump) {
                std::vector<int> path = g.gap_paths[st.edge]; if (pos) std::reverse(path.begin(), path.end());
                for (int e : path) { lv.push_back(g.node_level[g.edges[e].from]); ed.push_back(e); gc.push_back('_'); sc.push_back('_'); }
            } else if (pos) {
                if (st.x == x - 1 && st.y == y - 1) { lv.push_back(x - 1); ed.push_back(st.edge); gc.push_back(g.edges[st.edge].emis); sc.push_back(seq[y - 1]); }
                else if (st.x == x && st.y == y - 1) { lv.push_back(-1); ed.push_back(-1); gc.push_back('_'); sc.push_back(seq[y - 1]); }
                else if (st.x == x - 1 && st.y == y) { lv.push_back(x - 1); ed.push_back(st.edge); gc.push_back(g.edges[st.edge].emis); sc.push_back('_'); }
            } else {
                if (st.x == x + 1 && st.y == y + 1) { lv.push_back(x); ed.push_back(st.edge); gc.push_back(g.edges[st.edge].emis); sc.push_back(seq[y]); }
                else if (st.x == x && st.y == y + 1) { lv.push_back(-1); ed.push_back(-1); gc.push_back('_'); sc.push_back(seq[y]); }
                else if (st.x == x + 1 && st.y == y) { lv.push_back(x); ed.push_back(st.edge); gc.push_back(g.edges[st.edge].emis); sc.push_back('_'); }
            }
            k = Key(st.x, st.y, st.z); mat = st.mat;
        }
        if (pos) { std::reverse(lv.begin(), lv.end()); std::reverse(ed.begin(), ed.end()); std::reverse(gc.begin(), gc.end()); std::reverse(sc.begin(), sc.end()); }
        ext.level = lv; ext.edge = ed; ext.gchar = gc; ext.schar = sc; ext.from_seed.assign(lv.size(), 0);
        ext.seq_begin = pos ? start_seq : std::get<1>(endk); ext.seq_end = pos ? std::get<1>(endk) - 1 : start_seq - 1;
        REQUIRE(ext.seq_begin <= ext.seq_end, "extension covers at least one read base (VirtualNWUnique.cpp:34)");
        return true;
    }

    // extendSeedChain (extensionAligner.cpp:186-333) + extendToFullSequenceLength (verboseSeedChain.cpp:82-144)
    Chain extend_chain(const std::string& seq, const Chain& seed, bool dp = true) const {
        Chain r = seed;
        auto splice = [&](const Chain& e, bool left) {
            if (left) { r.level.insert(r.level.begin(), e.level.begin(), e.level.end()); r.edge.insert(r.edge.begin(), e.edge.begin(), e.edge.end()); r.gchar = e.gchar + r.gchar; r.schar = e.schar + r.schar; r.from_seed.insert(r.from_seed.begin(), e.from_seed.begin(), e.from_seed.end()); r.seq_begin = e.seq_begin; }
            else { r.level.insert(r.level.end(), e.level.begin(), e.level.end()); r.edge.insert(r.edge.end(), e.edge.begin(), e.edge.end()); r.gchar += e.gchar; r.schar += e.schar; r.from_seed.insert(r.from_seed.end(), e.from_seed.begin(), e.from_seed.end()); r.seq_end = e.seq_end; }
        };
        if (dp && seed.seq_begin != 0) {
            int node = g.edges[seed.edge.front()].from;
            if (g.node_level[node] > 0) { Chain e; if (extend(seq, seed.seq_begin, g.node_level[node], g.node_z[node], false, e)) splice(e, true); }
        }
        if (dp && seed.seq_end != (int)seq.size() - 1) {
            int node = g.edges[seed.edge.back()].to;
            if (g.node_level[node] < g.n_levels - 1) { Chain e; if (extend(seq, seed.seq_end + 1, g.node_level[node], g.node_z[node], true, e)) splice(e, false); }
        }
        int padL = r.seq_begin, padR = (int)seq.size() - 1 - r.seq_end;
        if (padL) { r.level.insert(r.level.begin(), padL, -1); r.edge.insert(r.edge.begin(), padL, -1); r.gchar = std::string(padL, '_') + r.gchar; r.schar = seq.substr(0, padL) + r.schar; r.from_seed.insert(r.from_seed.begin(), padL, 0); r.seq_begin = 0; }
        if (padR) { r.level.insert(r.level.end(), padR, -1); r.edge.insert(r.edge.end(), padR, -1); r.gchar += std::string(padR, '_'); r.schar += seq.substr(seq.size() - padR); r.from_seed.insert(r.from_seed.end(), padR, 0); r.seq_end = (int)seq.size() - 1; }
        return r;
    }

    // scoreOneAlignment (extensionAligner.cpp:52-182); qualities are consumed in alignment orientation
    static double score(const Chain& ch, const std::string& qual, bool long_reads = false) {
        double rate_del = log(0.001), rate_ins = log(0.001);
        if (long_reads) { rate_del = log(0.075); rate_ins = log(0.075); }      // extensionAligner.cpp:60-64
        double rate_mm = log(1 - exp(rate_del) - exp(rate_ins)), ll = 0; int idx = ch.seq_begin - 1;
        for (size_t i = 0; i < ch.schar.size(); i++) {
            if (ch.schar[i] != '_') {
                idx++;
                if (ch.gchar[i] == '_') ll += (rate_ins + log(1.0 / 4.0));
                else {
                    ll += rate_mm;
                    unsigned char q = (unsigned char)qual[idx]; REQUIRE(q >= 33, "phred >= 0");
                    double p = 1 - exp(log(10) * ((double)(q - 33) / (double)-10));
                    if (p > 0.999) p = 0.999; if (p == 0) p = 0.00001;
                    if (ch.schar[i] == ch.gchar[i]) ll += log(p); else { double w = 1 - p; w *= (1.0 / 3.0); ll += log(w); }
                }
            } else if (ch.gchar[i] != '_') ll += rate_del;
        }
        return ll;
    }

    struct ReadChains { std::vector<int> order; int primary = -1; std::string seq, qual; bool reverse = false; };
    ReadChains prepare_read(const Batch& b, long long r) const {   // sortChainsInSeeds (processBAM.cpp:1945) + primary lookup (protoSeeds.cpp:252-314)
        ReadChains rc; for (int c = b.chain_off[r]; c < b.chain_off[r + 1]; c++) rc.order.push_back(c);
        std::sort(rc.order.begin(), rc.order.end(), [&](int x, int y) { return b.chain_as[x] < b.chain_as[y]; }); std::reverse(rc.order.begin(), rc.order.end());
        for (size_t i = 0; i < rc.order.size(); i++) if (!(b.chain_flag[rc.order[i]] & 0x100)) { rc.primary = (int)i; break; }
        REQUIRE(rc.primary >= 0, "a primary alignment exists");
        rc.seq.assign((const char*)b.bases + b.read_off[r], (size_t)(b.read_off[r + 1] - b.read_off[r])); rc.qual.assign((const char*)b.quals + b.read_off[r], rc.seq.size());
        rc.reverse = (b.chain_flag[rc.order[rc.primary]] & 0x10) != 0;
        return rc;
    }

    std::pair<int, int> prg_span(const Batch& b, int c) const {   // alignment_get_startstop_PRGcoordinates (processBAM.cpp:3840)
        int reflen = 0; for (int k = b.cigar_off[c]; k < b.cigar_off[c + 1]; k++) { int op = b.cigar[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += (int)(b.cigar[k] >> 4); }
        const std::vector<int>& tr = g.contig_tr.at(b.chain_contig[c]);
        return {tr.at(b.chain_pos[c]), tr.at(b.chain_pos[c] + reflen - 1)};
    }

    static double normal_pdf(double mean, double sd, double x) { double e = x - mean; e *= -e; e /= 2 * sd * sd; double r = exp(e); r /= sd * sqrt(2 * 3.14159265358979323846264338327950288); return r; }
    static unsigned char to_phred(double p) { double w = 1 - p; if (w == 0) w = 1e-100; double ph = -10.0 * log10(w); if (ph + 33 > 255) ph = 255 - 33; return (unsigned char)(int)round(ph + 33); }

    std::map<int, int> anchors_of(const Chain& ch, bool end) const {   // verboseSeedChain.h:231-283 with nMaxLevels = 2
        std::vector<int> lv;
        if (end) { for (size_t i = ch.level.size(); i-- > 0 && lv.size() < 2;) if (ch.level[i] != -1) lv.push_back(ch.level[i]); }
        else { for (size_t i = 0; i < ch.level.size() && lv.size() < 2; i++) if (ch.level[i] != -1) lv.push_back(ch.level[i]); }
        std::map<int, int> r; for (int l : lv) for (auto& kv : g.anchors[l]) if (!r.count(kv.first)) r[kv.first] = kv.second;
        return r;
    }

    // alignOneReadPair (processBAM.cpp:3129-3616) + assignMappingQualities (:4062-4312)
    void pair(const Batch& b, long long p, double is_mean, double is_sd, Chain& out1, Chain& out2, double& pair_mapq) const {
        double pen = log(normal_pdf(is_mean, is_sd, is_mean + 8 * is_sd));
        std::vector<Chain> ch[2]; std::vector<double> ll[2];
        for (int w = 0; w < 2; w++) {
            ReadChains rc = prepare_read(b, 2 * p + w); std::map<std::pair<int, int>, int> seen;
            for (size_t i = 0; i < rc.order.size(); i++) {
                int c = rc.order[i];
                if (((b.chain_flag[c] & 0x10) != 0) != rc.reverse) continue;
                std::pair<int, int> id = prg_span(b, c);
                Chain seed = project(b, c, rc.seq, rc.reverse);      // the reference projects before it checks for duplicates
                if (seen.count(id) && seen[id] >= b.chain_as[c]) continue;
                Chain full = extend_chain(rc.seq, seed);
                ch[w].push_back(full); ll[w].push_back(score(full, rc.qual));
                if (!seen.count(id) || seen[id] < b.chain_as[c]) seen[id] = b.chain_as[c];
            }
            REQUIRE(!ch[w].empty(), "at least one chain per read");
        }
        std::vector<double> combo; std::vector<std::pair<int, int>> idx;
        for (size_t a = 0; a < ch[0].size(); a++) for (size_t c2 = 0; c2 < ch[1].size(); c2++) {
            const Chain& A = ch[0][a]; const Chain& B = ch[1][c2];
            double v = ll[0][a] + ll[1][c2];
            bool valid = false;
            if (A.first_level() != -1 && B.first_level() != -1 && A.reverse != B.reverse) valid = !A.reverse ? (A.first_level() < B.first_level()) : (A.last_level() > B.last_level());
            double lis = pen;
            if (valid) {
                std::set<int> dist;
                const Chain& up = A.first_level() < B.first_level() ? A : B; const Chain& dn = A.first_level() < B.first_level() ? B : A;
                std::map<int, int> e = anchors_of(up, true), s = anchors_of(dn, false);
                for (auto& kv : e) if (s.count(kv.first)) dist.insert(s[kv.first] - kv.second - 1);
                if (!dist.empty()) { bool first = true; for (int d : dist) { double pd = normal_pdf(is_mean, is_sd, d); double l = pd <= 0 ? pen : log(pd); if (first || l > lis) { lis = l; first = false; } } }
            }
            v += lis; combo.push_back(v); idx.push_back({(int)a, (int)c2});
        }
        size_t bi = 0; for (size_t i = 1; i < combo.size(); i++) if (combo[i] > combo[bi]) bi = i;
        out1 = ch[0][idx[bi].first]; out2 = ch[1][idx[bi].second];
        if (combo.size() == 1) { pair_mapq = 1; out1.chain_mapq = out2.chain_mapq = 1; out1.mapq.assign(out1.gchar.size(), (char)to_phred(1)); out2.mapq.assign(out2.gchar.size(), (char)to_phred(1)); return; }
        std::vector<double> pp(combo.size()); double sum = 0;
        for (size_t i = 0; i < combo.size(); i++) pp[i] = exp(combo[i] - combo[bi]);
        for (double v : pp) sum += v; for (double& v : pp) v = v / sum;
        pair_mapq = pp[bi]; double m1 = 0, m2 = 0;
        for (size_t i = 0; i < pp.size(); i++) { if (idx[i].first == idx[bi].first) m1 += pp[i]; if (idx[i].second == idx[bi].second) m2 += pp[i]; }
        out1.chain_mapq = std::min(m1, 1.0); out2.chain_mapq = std::min(m2, 1.0);
        auto keys = [](const Chain& c, const char* tag) { std::vector<std::string> k; int nb = 0, tot = 0; for (char x : c.schar) if (x != '_') tot++;
            for (size_t j = 0; j < c.gchar.size(); j++) { int si = -1; if (c.schar[j] != '_') { si = c.reverse ? tot - nb - 1 : nb; nb++; }
                k.push_back(std::string(1, c.gchar[j]) + ":" + std::to_string(c.level[j]) + ":" + tag + ":" + (c.reverse ? "minus" : "plus") + ":" + std::to_string(si)); } return k; };
        std::map<std::string, double> conf;
        for (size_t i = 0; i < pp.size(); i++) { for (const std::string& k : keys(ch[0][idx[i].first], "r1")) conf[k] += pp[i]; for (const std::string& k : keys(ch[1][idx[i].second], "r2")) conf[k] += pp[i]; }
        auto assign = [&](Chain& c, const char* tag) { c.mapq.clear(); for (const std::string& k : keys(c, tag)) { double q = conf.at(k); if (q > 1) q = 1; c.mapq.push_back((char)to_phred(q)); } };
        assign(out1, "r1"); assign(out2, "r2");
    }

    // alignOneLongRead (processBAM.cpp:3618-3838: the seed is only padded to the read, verboseSeedChain.cpp:82-144; indel rates 0.075) +
    // assignMappingQualities_unpaired (:3900-4059)
    void long_read(const Batch& b, long long r, Chain& out, double& best_ll) const {
        std::vector<Chain> ch; std::vector<double> ll;
        ReadChains rc = prepare_read(b, r); std::map<std::pair<int, int>, int> seen;
        for (size_t i = 0; i < rc.order.size(); i++) {
            int c = rc.order[i];
            if (((b.chain_flag[c] & 0x10) != 0) != rc.reverse) continue;
            std::pair<int, int> id = prg_span(b, c);
            Chain seed = project(b, c, rc.seq, rc.reverse);
            if (seen.count(id) && seen[id] >= b.chain_as[c]) continue;
            Chain full = extend_chain(rc.seq, seed, false);
            ch.push_back(full); ll.push_back(score(full, rc.qual, true));
            if (!seen.count(id) || seen[id] < b.chain_as[c]) seen[id] = b.chain_as[c];
        }
        REQUIRE(!ch.empty(), "at least one chain per read");
        size_t bi = 0; for (size_t i = 1; i < ll.size(); i++) if (ll[i] > ll[bi]) bi = i;
        out = ch[bi]; best_ll = ll[bi];
        if (ll.size() == 1) { out.chain_mapq = 1; out.mapq.assign(out.gchar.size(), (char)to_phred(1)); return; }
        std::vector<double> pp(ll.size()); double sum = 0;
        for (size_t i = 0; i < ll.size(); i++) pp[i] = exp(ll[i] - ll[bi]);
        for (double v : pp) sum += v; for (double& v : pp) v = v / sum;
        out.chain_mapq = pp[bi];
        auto keys = [](const Chain& c) { std::vector<std::string> k; int nb = 0, tot = 0; for (char x : c.schar) if (x != '_') tot++;
            for (size_t j = 0; j < c.gchar.size(); j++) { int si = -1; if (c.schar[j] != '_') { si = c.reverse ? tot - nb - 1 : nb; nb++; }
                k.push_back(std::string(1, c.gchar[j]) + ":" + std::to_string(c.level[j]) + ":r1:" + (c.reverse ? "minus" : "plus") + ":" + std::to_string(si)); } return k; };
        std::map<std::string, double> conf;
        for (size_t i = 0; i < pp.size(); i++) for (const std::string& k : keys(ch[i])) conf[k] += pp[i];
        out.mapq.clear(); for (const std::string& k : keys(out)) { double q = conf.at(k); if (q > 1) q = 1; out.mapq.push_back((char)to_phred(q)); }
    }
};

std::string g_err;
template <class F> int guarded(F&& f) { try { return f(); } catch (const std::exception& e) { g_err = e.what(); return -1; } catch (...) { g_err = "unknown"; return -1; } }

void export_chain(const Chain& ch, int cap, int32_t* n_cols, int32_t* level, int32_t* edge, uint8_t* gchar, uint8_t* schar, uint8_t* from_seed, uint8_t* mapq) {
    int n = (int)ch.level.size(); *n_cols = n;
    for (int i = 0; i < n && i < cap; i++) { level[i] = ch.level[i]; edge[i] = ch.edge[i]; gchar[i] = (uint8_t)ch.gchar[i]; schar[i] = (uint8_t)ch.schar[i]; if (from_seed) from_seed[i] = (uint8_t)ch.from_seed[i]; if (mapq) mapq[i] = ch.mapq.size() > (size_t)i ? (uint8_t)ch.mapq[i] : 0; }
}

} // namespace

extern "C" {

const char* hlala_oracle_last_error() { return g_err.c_str(); }

void* hlala_oracle_open(const char* dir) {
    Oracle* o = nullptr;
    int rc = guarded([&]() { o = new Oracle(); read_graph(std::string(dir) + "/PRG/graph.txt", o->g); gap_paths(o->g); gap_stretches(o->g); read_contigs(dir, o->g); return 0; });
    if (rc != 0) { delete o; return nullptr; }
    return o;
}
void hlala_oracle_close(void* h) { delete (Oracle*)h; }
long long hlala_oracle_n_levels(void* h) { return ((Oracle*)h)->g.n_levels; }
long long hlala_oracle_n_nodes(void* h) { return (long long)((Oracle*)h)->g.node_level.size(); }
long long hlala_oracle_n_edges(void* h) { return (long long)((Oracle*)h)->g.edges.size(); }
int hlala_oracle_graph_export(void* h, int32_t* node_level, int32_t* edge_from, int32_t* edge_to, uint8_t* edge_emis) {
    const OGraph& g = ((Oracle*)h)->g;
    for (size_t i = 0; i < g.node_level.size(); i++) node_level[i] = g.node_level[i];
    for (size_t i = 0; i < g.edges.size(); i++) { edge_from[i] = g.edges[i].from; edge_to[i] = g.edges[i].to; edge_emis[i] = (uint8_t)g.edges[i].emis; }
    return 0;
}
long long hlala_oracle_gap_paths_total(void* h, long long* n_paths) { const OGraph& g = ((Oracle*)h)->g; long long t = 0; *n_paths = (long long)g.gap_paths.size(); for (auto& p : g.gap_paths) t += (long long)p.size(); return t; }
int hlala_oracle_gap_paths_export(void* h, int64_t* path_off, int32_t* path_edges) { const OGraph& g = ((Oracle*)h)->g; int64_t o = 0; size_t i = 0; for (auto& p : g.gap_paths) { path_off[i++] = o; for (int e : p) path_edges[o++] = e; } path_off[i] = o; return 0; }
int hlala_oracle_gap_stretch(void* h, uint8_t* out) { const OGraph& g = ((Oracle*)h)->g; for (size_t i = 0; i < g.gap_stretch.size(); i++) out[i] = (uint8_t)g.gap_stretch[i]; return (int)g.gap_stretch.size(); }

int hlala_oracle_chains(void* h, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, const int32_t* chain_off, const int32_t* chain_contig,
                        const int32_t* chain_pos, const uint16_t* chain_flag, const int32_t* chain_as, const int32_t* cigar_off, const uint32_t* cigar, int cap,
                        int32_t* chain_order, int32_t* status, int32_t* n_cols, int32_t* seed_begin, int32_t* seed_end, double* ll,
                        int32_t* level, int32_t* edge, uint8_t* gchar, uint8_t* schar, uint8_t* from_seed) {
    Oracle* o = (Oracle*)h; Batch b{n_reads, read_off, bases, quals, chain_off, chain_contig, chain_pos, chain_flag, chain_as, cigar_off, cigar};
    return guarded([&]() {
        for (long long r = 0; r < n_reads; r++) {
            if (chain_off[r + 1] == chain_off[r]) continue;
            Oracle::ReadChains rc = o->prepare_read(b, r);
            for (size_t i = 0; i < rc.order.size(); i++) {
                int slot = chain_off[r] + (int)i, c = rc.order[i]; chain_order[slot] = c;
                if (((chain_flag[c] & 0x10) != 0) != rc.reverse) { status[slot] = 1; n_cols[slot] = 0; ll[slot] = 0; continue; }
                Chain seed = o->project(b, c, rc.seq, rc.reverse); seed_begin[slot] = seed.seq_begin; seed_end[slot] = seed.seq_end;
                Chain full = o->extend_chain(rc.seq, seed); ll[slot] = Oracle::score(full, rc.qual); status[slot] = 0;
                size_t off = (size_t)slot * cap; export_chain(full, cap, n_cols + slot, level + off, edge + off, gchar + off, schar + off, from_seed + off, nullptr);
            }
        }
        return 0;
    });
}

int hlala_oracle_pairs(void* h, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, const int32_t* chain_off, const int32_t* chain_contig,
                       const int32_t* chain_pos, const uint16_t* chain_flag, const int32_t* chain_as, const int32_t* cigar_off, const uint32_t* cigar, double is_mean, double is_sd, int cap, int threads,
                       double* pair_mapq, double* read_mapq, uint8_t* read_reverse, int32_t* n_cols, int32_t* level, int32_t* edge, uint8_t* gchar, uint8_t* schar, uint8_t* from_seed, uint8_t* mapq, double* seconds) {
    Oracle* o = (Oracle*)h; Batch b{n_reads, read_off, bases, quals, chain_off, chain_contig, chain_pos, chain_flag, chain_as, cigar_off, cigar}; (void)threads;
    return guarded([&]() {
        auto t0 = std::chrono::steady_clock::now();
        for (long long p = 0; p < n_reads / 2; p++) {
            Chain a, c; double mq; o->pair(b, p, is_mean, is_sd, a, c, mq);
            pair_mapq[p] = mq; read_mapq[2 * p] = a.chain_mapq; read_mapq[2 * p + 1] = c.chain_mapq; read_reverse[2 * p] = a.reverse; read_reverse[2 * p + 1] = c.reverse;
            if (level) { size_t oa = (size_t)(2 * p) * cap, ob = oa + cap; export_chain(a, cap, n_cols + 2 * p, level + oa, edge + oa, gchar + oa, schar + oa, from_seed + oa, mapq + oa); export_chain(c, cap, n_cols + 2 * p + 1, level + ob, edge + ob, gchar + ob, schar + ob, from_seed + ob, mapq + ob); }
            else { n_cols[2 * p] = (int)a.level.size(); n_cols[2 * p + 1] = (int)c.level.size(); }
        }
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    });
}

// one read per unit (long-read mode): read_mapq[r] = posterior of the chosen chain, read_ll[r] its log-likelihood
int hlala_oracle_long_reads(void* h, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, const int32_t* chain_off, const int32_t* chain_contig,
                            const int32_t* chain_pos, const uint16_t* chain_flag, const int32_t* chain_as, const int32_t* cigar_off, const uint32_t* cigar, int cap,
                            double* read_mapq, double* read_ll, uint8_t* read_reverse, int32_t* n_cols, int32_t* level, int32_t* edge, uint8_t* gchar, uint8_t* schar, uint8_t* from_seed, uint8_t* mapq, double* seconds) {
    Oracle* o = (Oracle*)h; Batch b{n_reads, read_off, bases, quals, chain_off, chain_contig, chain_pos, chain_flag, chain_as, cigar_off, cigar};
    return guarded([&]() {
        auto t0 = std::chrono::steady_clock::now();
        for (long long r = 0; r < n_reads; r++) {
            Chain a; double ll; o->long_read(b, r, a, ll);
            read_mapq[r] = a.chain_mapq; read_ll[r] = ll; read_reverse[r] = a.reverse;
            if (level) { size_t oa = (size_t)r * cap; export_chain(a, cap, n_cols + r, level + oa, edge + oa, gchar + oa, schar + oa, from_seed + oa, mapq + oa); }
            else n_cols[r] = (int)a.level.size();
        }
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    });
}

} // extern "C"
