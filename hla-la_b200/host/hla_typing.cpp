// Host side of the HLA typing stage (see hla_typing.h). Everything here is bookkeeping around the two GPU kernels: table loading,
// exon projection of the chosen alignments, the global read/allele filters, ranking, calls, QC and the text files the reference writes.
// Containers whose ITERATION ORDER reaches an output (std::map by string / by position, std::sort on ties) are the same
// containers the reference uses, because that order is part of the result.
#include "hla_typing.h"

#include <algorithm>
#include <array>
#include <exception>
#include <functional>
#include <memory>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <cmath>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <sstream>
#include <type_traits>
#include <stdexcept>
#include <sys/stat.h>
#include <unordered_set>
#include <chrono>
#include <charconv>

namespace hlala {

namespace {

[[noreturn]] void invariant(const std::string& what) { throw std::runtime_error("typing: reference assertion would fail: " + what); }
#define TY_REQUIRE(c, what) do { if (!(c)) invariant(what); } while (0)

// Utilities::split (Utilities.cpp:610): empty input -> no fields; otherwise every field, empty ones included
std::vector<std::string> split_on(const std::string& s, const std::string& d) {
    std::vector<std::string> out; if (s.empty()) return out;
    size_t at = 0;
    while (true) { size_t q = s.find(d, at); if (q == std::string::npos) { out.emplace_back(s, at); break; } out.emplace_back(s, at, q - at); at = q + d.size(); }
    return out;
}
std::string join_with(const std::vector<std::string>& v, const char* d) { std::string r; for (size_t i = 0; i < v.size(); i++) { if (i) r += d; r += v[i]; } return r; }
void chomp(std::string& s) { while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); }
// default ostream formatting == Utilities::ItoStr / DtoStr; spelled with snprintf / to_string because constructing a stream per number dominated
// the host time of the typing stage ("%g" is what operator<<(double) prints at the default precision of 6)
// std::to_chars(general, 6) is specified to produce what printf("%.6g") produces
template <class T> std::string str(T v) { if constexpr (std::is_floating_point<T>::value) { char b[48]; auto r = std::to_chars(b, b + sizeof b, (double)v, std::chars_format::general, 6); return std::string(b, r.ptr); } else if constexpr (std::is_integral<T>::value) return std::to_string(v); else { std::ostringstream o; o << v; return o.str(); } }
inline void put_num(std::string& o, double v) { char b[48]; auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::general, 6); o.append(b, r.ptr); }
inline void put_num(std::string& o, long long v) { char b[24]; auto r = std::to_chars(b, b + sizeof b, v); o.append(b, r.ptr); }

std::vector<std::string> read_lines(const std::string& path, bool keep_trailing_empty) {
    std::ifstream f(path); if (!f.is_open()) throw std::runtime_error("Cannot open file " + path);
    std::vector<std::string> L; std::string l;
    while (f.good()) { std::getline(f, l); chomp(l); L.push_back(l); }   // the reference's while(good()) { getline } loop yields a final empty line
    if (!keep_trailing_empty) while (!L.empty() && L.back().empty()) L.pop_back();
    return L;
}

double pcorrect_of(unsigned char q) {   // Utilities::PhredToPCorrect (Utilities.cpp:357-377)
    if (q == 0) return -1;
    int ph = (int)q - 33; TY_REQUIRE(ph >= 0, "phred >= 0");
    return 1 - exp(log(10) * ((double)ph / (double)-10));
}

const char* const LOCI17[] = {"A", "B", "C", "DQA1", "DQB1", "DRB1", "DPA1", "DPB1", "DRA", "DRB3", "DRB4", "E", "F", "G", "H", "K", "V"};
bool two_exon_locus(const std::string& l) { static const std::set<std::string> s = {"A", "B", "C", "E", "F", "G", "H", "J", "K", "L", "V"}; return s.count(l) != 0; }

} // namespace

// ------------------------------------------------------------------------------------------------------------ tables
void TypingTables::load(const std::string& dir) {
    prg_dir = dir; loci.clear(); gene_bounds.clear();
    std::vector<std::string> entries;   // directory order: "the last matching entry wins" (HLATyper.cpp:3127-3190)
    { DIR* dp = opendir((dir + "/PRG").c_str()); if (!dp) throw std::runtime_error("Cannot open directory " + dir + "/PRG");
      while (dirent* e = readdir(dp)) { std::string n = e->d_name; if (n != "." && n != "..") entries.push_back(n); } closedir(dp); }
    std::vector<std::string> seg = read_lines(dir + "/PRG/segments.txt", false);
    seg.erase(std::remove_if(seg.begin(), seg.end(), [](const std::string& s) { return s.empty(); }), seg.end());
    std::map<std::string, int32_t> level_of;     // Graph::readGraphLoci: level = position in the concatenated header lines
    std::vector<std::vector<std::string>> headers;
    for (const std::string& f : seg) {
        std::ifstream s(dir + "/PRG/" + f); if (!s.is_open()) throw std::runtime_error("Cannot open segment file " + dir + "/PRG/" + f);
        std::string h; std::getline(s, h); chomp(h); headers.push_back(split_on(h, " "));
        const auto& fs = headers.back();
        for (size_t i = 1; i < fs.size(); i++) { TY_REQUIRE(!level_of.count(fs[i]), "level names unique"); int32_t l = (int32_t)level_of.size(); level_of[fs[i]] = l; }
    }
    {   // gene boundaries (HLATyper.cpp:105-214), keyed by locus like the reference's std::map
        std::map<std::string, std::pair<int32_t, int32_t>> gb;
        for (size_t k = 0; k < seg.size(); k++) {
            std::vector<std::string> u = split_on(seg[k], "_"); if (u.size() < 3 || u[1] != "gene") continue;
            TY_REQUIRE(!headers[k].empty() && headers[k][0] == "IndividualID", "segment header starts with IndividualID");
            auto it = gb.find(u[2]); if (it == gb.end()) it = gb.insert({u[2], {-1, -1}}).first;
            for (size_t i = 1; i < headers[k].size(); i++) { int32_t l = level_of.at(headers[k][i]); if (it->second.first == -1 || l < it->second.first) it->second.first = l; if (it->second.second == -1 || l > it->second.second) it->second.second = l; }
        }
        for (auto& kv : gb) gene_bounds.push_back(kv.second);
    }
    for (const char* lname : LOCI17) {
        TypingLocus L; L.name = lname;
        const int n_exons = two_exon_locus(L.name) ? 2 : 1;
        std::map<std::string, std::string> seqs;      // HLA type -> concatenated exon sequence; std::map order defines the cluster order
        for (int ei = 0; ei < n_exons; ei++) {
            const std::string want = str(ei + 2) + ".txt"; std::string file;
            for (const std::string& e : entries) { std::vector<std::string> u = split_on(e, "_");
                if (u.size() >= 6 && u[1] == "gene" && (u[2] == "HLA-" + L.name || u[2] == L.name) && u[4] == "exon" && u[5] == want) file = e; }
            if (file.empty()) throw std::runtime_error("typing: no exon " + str(ei + 2) + " file for locus " + L.name + " in " + dir + "/PRG");
            std::vector<std::string> lines = read_lines(dir + "/PRG/" + file, true);
            std::vector<std::string> hf = split_on(lines.at(0), " "); TY_REQUIRE(!hf.empty() && hf[0] == "IndividualID", "exon file header");
            TY_REQUIRE(hf.size() >= 3, "exon file has at least two columns");
            const int32_t first = level_of.at(hf[1]), last = level_of.at(hf.back()); TY_REQUIRE(last > first, "last_graph_level > first_graph_level");
            const int32_t len = last - first + 1; TY_REQUIRE((int32_t)hf.size() - 1 == len, "exon columns are consecutive levels");
            for (int32_t i = 0; i < len; i++) { TY_REQUIRE(level_of.at(hf[1 + i]) == first + i, "exon columns are consecutive levels");
                L.col_level.push_back(first + i); L.col_exon.push_back(ei); L.col_exonpos.push_back(i);
                if (L.lmin == -1 || first + i < L.lmin) L.lmin = first + i;
                if (L.lmax == -1 || first + i > L.lmax) L.lmax = first + i; }
            L.exon_len.push_back(len);
            for (size_t li = 1; li < lines.size(); li++) {
                if (lines[li].empty()) continue;
                std::vector<std::string> f = split_on(lines[li], " "); TY_REQUIRE(f.size() == hf.size(), "exon file field count");
                if (f[0].find(':') == std::string::npos) continue;        // only proper HLA types (HLATyper.cpp:1270)
                std::string sq; for (size_t k = 1; k < f.size(); k++) sq += f[k];
                if (ei == 0) { TY_REQUIRE(!seqs.count(f[0]), "HLA type listed once"); seqs[f[0]] = sq; }
                else { auto it = seqs.find(f[0]); TY_REQUIRE(it != seqs.end(), "HLA type present in every exon file"); it->second += sq; }
            }
            TY_REQUIRE(!seqs.empty(), "exon file has sequences");
        }
        std::map<std::string, size_t> by_seq;
        for (auto& kv : seqs) {
            TY_REQUIRE((int32_t)kv.second.size() == L.P(), "one symbol per exon column");
            auto it = by_seq.find(kv.second);
            if (it == by_seq.end()) { by_seq[kv.second] = L.cluster_seq.size(); L.cluster_seq.push_back(kv.second); L.cluster_members.push_back({kv.first}); }
            else L.cluster_members[it->second].push_back(kv.first);      // seqs iterates in sorted order, so members stay sorted like the reference's std::set
        }
        loci.push_back(std::move(L));
    }
}

void TypingTables::load_G(const std::string& dir) {
    if (g_loaded) return;
    std::ifstream g(dir + "/hla_nom_g.txt");
    if (!g.is_open()) throw std::runtime_error("Can't open file hla_nom_g.txt - are you executing me from the right directory?");
    std::string line;
    while (g.good()) {
        std::getline(g, line); chomp(line); if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> c = split_on(line, ";"); TY_REQUIRE(c.size() >= 2, "hla_nom_g.txt fields");
        const std::string ls = c.front(); TY_REQUIRE(!ls.empty() && ls.back() == '*', "hla_nom_g.txt locus ends with *"); G_loci.insert(ls.substr(0, ls.size() - 1));
        std::string code; if (!c.back().empty()) code = c.back(); else { TY_REQUIRE(c.size() == 3, "hla_nom_g.txt three components"); code = c[1]; }
        for (const std::string& a : split_on(c[1], "/")) alleles_to_G[ls + a] = ls + code;
    }
    g_loaded = true;
}

bool pair_overlaps_genes(const TypingTables& T, int32_t f1, int32_t l1, int32_t f2, int32_t l2) {
    for (const auto& gb : T.gene_bounds) {   // HLATyper::intervalOverlapsWithGenes (HLATyper.cpp:259)
        if (f1 != -1 && gb.second >= f1 && gb.first <= l1) return true;
        if (f2 != -1 && gb.second >= f2 && gb.first <= l2) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------------------ TypingReads
void TypingReads::clear() { *this = TypingReads(); }
void TypingReads::append(const TypingReads& o) {
    if (col_off.empty()) { col_off.push_back(0); base_off.push_back(0); }
    const int64_t c0 = col_off.back(), b0 = base_off.back();
    pair_id.insert(pair_id.end(), o.pair_id.begin(), o.pair_id.end()); name.insert(name.end(), o.name.begin(), o.name.end());
    for (size_t i = 1; i < o.col_off.size(); i++) col_off.push_back(c0 + o.col_off[i]);
    for (size_t i = 1; i < o.base_off.size(); i++) base_off.push_back(b0 + o.base_off[i]);
    level.insert(level.end(), o.level.begin(), o.level.end()); g.insert(g.end(), o.g.begin(), o.g.end()); s.insert(s.end(), o.s.begin(), o.s.end()); mq.insert(mq.end(), o.mq.begin(), o.mq.end());
    bases.insert(bases.end(), o.bases.begin(), o.bases.end()); quals.insert(quals.end(), o.quals.begin(), o.quals.end());
    reverse.insert(reverse.end(), o.reverse.begin(), o.reverse.end()); mapq.insert(mapq.end(), o.mapq.begin(), o.mapq.end());
}
namespace {
template <class T> void put_vec(std::vector<uint8_t>& b, const std::vector<T>& v) { uint64_t n = v.size(); const uint8_t* p = (const uint8_t*)&n; b.insert(b.end(), p, p + 8); const uint8_t* q = (const uint8_t*)v.data(); b.insert(b.end(), q, q + n * sizeof(T)); while (b.size() % 8) b.push_back(0); }
template <class T> void get_vec(const uint8_t*& p, const uint8_t* end, std::vector<T>& v) { if (p + 8 > end) throw std::runtime_error("typing blob truncated"); uint64_t n; memcpy(&n, p, 8); p += 8; if (p + n * sizeof(T) > end) throw std::runtime_error("typing blob truncated"); v.resize(n); if (n) memcpy(v.data(), p, n * sizeof(T)); p += n * sizeof(T); while ((uintptr_t)(p - (const uint8_t*)nullptr) % 8 && p < end) p++; }
}
std::vector<uint8_t> TypingReads::serialize() const {
    std::vector<uint8_t> b; std::vector<uint64_t> magic = {long_reads ? 0x484C415459503032ull : 0x484C415459503031ull};   // "HLATYP01" paired, "HLATYP02" long reads
    put_vec(b, magic); put_vec(b, pair_id);
    std::vector<int64_t> noff(1, 0); std::vector<uint8_t> nch; for (const std::string& n : name) { nch.insert(nch.end(), n.begin(), n.end()); noff.push_back((int64_t)nch.size()); }
    put_vec(b, noff); put_vec(b, nch); put_vec(b, col_off); put_vec(b, level); put_vec(b, g); put_vec(b, s); put_vec(b, mq); put_vec(b, base_off); put_vec(b, bases); put_vec(b, quals); put_vec(b, reverse); put_vec(b, mapq);
    return b;
}
void TypingReads::deserialize_append(const uint8_t* p, size_t n) {
    // blobs are 8-byte aligned internally relative to their start; copy to an aligned buffer to keep get_vec's padding rule simple
    std::vector<uint64_t> aligned((n + 7) / 8); memcpy(aligned.data(), p, n);
    const uint8_t* q = (const uint8_t*)aligned.data(); const uint8_t* end = q + n;
    TypingReads o; std::vector<uint64_t> magic; get_vec(q, end, magic); if (magic.size() != 1 || (magic[0] != 0x484C415459503031ull && magic[0] != 0x484C415459503032ull)) throw std::runtime_error("typing blob: bad magic");
    o.long_reads = magic[0] == 0x484C415459503032ull;
    if (!pair_id.empty() && o.long_reads != long_reads) throw std::runtime_error("typing blob: paired and long-read blobs cannot be mixed");
    long_reads = o.long_reads;
    get_vec(q, end, o.pair_id); std::vector<int64_t> noff; std::vector<uint8_t> nch; get_vec(q, end, noff); get_vec(q, end, nch);
    for (size_t i = 0; i + 1 < noff.size(); i++) o.name.emplace_back((const char*)nch.data() + noff[i], (size_t)(noff[i + 1] - noff[i]));
    get_vec(q, end, o.col_off); get_vec(q, end, o.level); get_vec(q, end, o.g); get_vec(q, end, o.s); get_vec(q, end, o.mq); get_vec(q, end, o.base_off); get_vec(q, end, o.bases); get_vec(q, end, o.quals); get_vec(q, end, o.reverse); get_vec(q, end, o.mapq);
    if (o.name.size() != o.pair_id.size() || o.col_off.size() != 2 * o.pair_id.size() + 1 || o.base_off.size() != o.col_off.size()) throw std::runtime_error("typing blob: inconsistent sizes");
    append(o);
}

TypingReads long_read_typing_input(const TypingTables& T, int64_t n_reads, const char* const* names, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int32_t cap,
                                   const int32_t* n_cols, const int32_t* level, const uint8_t* g, const uint8_t* s, const uint8_t* mq, const uint8_t* reverse, const double* mapq) {
    TypingReads tr; tr.long_reads = true; tr.col_off.push_back(0); tr.base_off.push_back(0);
    for (int64_t r = 0; r < n_reads; r++) {
        const size_t o = (size_t)r * (size_t)cap; const int n = n_cols[r];
        if (n > cap) throw std::runtime_error("typing: a long-read alignment has more columns than max_columns");
        int f = -1, l = -1; for (int c = 0; c < n; c++) { const int lv = level[o + c]; if (lv != -1) { if (f == -1) f = lv; l = lv; } }
        if (!pair_overlaps_genes(T, f, l, -1, -1)) continue;
        tr.pair_id.push_back(r); tr.name.push_back(names && names[r] ? std::string(names[r]) : "r" + std::to_string(r));
        tr.level.insert(tr.level.end(), level + o, level + o + n); tr.g.insert(tr.g.end(), g + o, g + o + n); tr.s.insert(tr.s.end(), s + o, s + o + n); tr.mq.insert(tr.mq.end(), mq + o, mq + o + n);
        tr.col_off.push_back((int64_t)tr.level.size()); tr.col_off.push_back((int64_t)tr.level.size());
        tr.bases.insert(tr.bases.end(), bases + read_off[r], bases + read_off[r + 1]); tr.quals.insert(tr.quals.end(), quals + read_off[r], quals + read_off[r + 1]);
        tr.base_off.push_back((int64_t)tr.bases.size()); tr.base_off.push_back((int64_t)tr.bases.size());
        tr.reverse.push_back(reverse[r]); tr.reverse.push_back(0); tr.mapq.push_back(mapq[r]); tr.mapq.push_back(0);
    }
    return tr;
}

// ------------------------------------------------------------------------------------------------------------ inference
namespace {

struct Mate {     // one read's chosen alignment
    int n = 0; const int32_t* level = nullptr; const uint8_t* g = nullptr; const uint8_t* s = nullptr; const uint8_t* mq = nullptr;
    int len = 0; const uint8_t* bases = nullptr; const uint8_t* quals = nullptr; bool reverse = false; double mapQ = 0;
    int first_level = -1, last_level = -1, nongap_cols = 0; double weighted_ok = 0, fraction_ok = 0; const std::string* name = nullptr;
    // text of this read's observations in R1_pileup_<L>.txt / histogram_matchesPerRead.txt, formatted once (the same pieces repeat for every exon column)
    std::string pile_mid;    // ") [pairsDistance <d> | alignmentLength <n> | "
    std::string pile_tail;   // " | <mapQ> <mapQ> | <weighted self> <weighted mate> | <name> <mate name>]"
    std::string hist_base;   // "base<weighted>\n"
};

void mate_stats(Mate& m) {
    m.first_level = -1; m.last_level = -1;
    for (int c = 0; c < m.n; c++) if (m.level[c] != -1) { if (m.first_level == -1) m.first_level = m.level[c]; m.last_level = m.level[c]; }
    // alignmentFractionOK (HLATyper.cpp:3082)
    int ok = 0, chk = 0; for (int c = 0; c < m.n; c++) { if (m.g[c] == '_' && m.s[c] == '_') continue; chk++; if (m.g[c] == m.s[c]) ok++; }
    m.fraction_ok = (double)ok / (double)chk;
    // alignmentWeightedOKFraction (HLATyper.cpp:3933-4018): qualities are indexed in alignment orientation (the two strand flips cancel)
    int idx = -1, total = 0; double w = 0; m.nongap_cols = 0;
    for (int c = 0; c < m.n; c++) {
        if (m.s[c] != '_') {
            m.nongap_cols++; idx++; TY_REQUIRE(idx < m.len, "read index inside the read"); TY_REQUIRE(m.bases[idx] == m.s[c], "alignment spells the read");
            if (m.g[c] == '_') { total++; w++; }
            else if (m.s[c] != m.g[c]) { double p = pcorrect_of(m.quals[idx]); TY_REQUIRE(p >= 0 && p <= 1, "pCorrect in [0,1]"); w += p; total++; }
            else { double p = pcorrect_of(m.quals[idx]); TY_REQUIRE(p >= 0 && p <= 1, "pCorrect in [0,1]"); }
        } else if (m.g[c] != '_') { total++; w++; }
    }
    TY_REQUIRE(total >= w, "totalMismatches >= weightedMismatches");
    m.weighted_ok = 1.0 - (w / (double)m.len);
}
bool strands_ok(const Mate& a, const Mate& b) {   // alignerBase::alignedReadPair_strandsValid (alignerBase.cpp:213-245)
    if (a.first_level == -1 || b.first_level == -1 || a.reverse == b.reverse) return false;
    return !a.reverse ? (a.first_level < b.first_level) : (a.last_level > b.last_level);
}
int level_distance(const Mate& a, const Mate& b) { return a.first_level < b.first_level ? b.first_level - a.last_level - 1 : a.first_level - b.last_level - 1; }   // alignerBase.cpp:247-288

struct ExonObs {    // hla::oneExonPosition (hla/oneExonPosition.h:15-46), fields that are read anywhere in short-read mode
    uint32_t pos = 0; int32_t level = -1; std::string genotype, qualities; const Mate* self = nullptr; const Mate* mate = nullptr;
    double dist = 0, mapq_pos = 0; unsigned char mq_char = 0;    // mq_char: the per-column quality character mapq_pos was derived from
    int novel_gap = 0;                                           // runningNovelGapEitherDirection (HLATyper.cpp:3616-3662): only read in long-read mode
};

// oneReadAlignment_2_exonPositions_paired (HLATyper.cpp:3192-3565)
// and oneReadAlignment_2_exonPositions_unpaired (:3568-3930) with unpaired = true: no mate (distance -1), running novel gaps recorded
void project_read(const Mate& A, const Mate& M, const TypingLocus& L, std::vector<ExonObs>& out, bool unpaired = false) {
    TY_REQUIRE(A.first_level <= A.last_level, "alignment_firstLevel <= alignment_lastLevel");
    const int fl = A.first_level, ll = A.last_level;
    if (!((fl >= L.lmin && fl <= L.lmax) || (ll >= L.lmin && ll <= L.lmax) || (L.lmin >= fl && L.lmin <= ll) || (L.lmax >= fl && L.lmax <= ll))) return;
    const double dist = unpaired ? -1.0 : level_distance(A, M);
    std::vector<ExonObs> all; all.reserve((size_t)A.n);
    std::vector<int> novel;
    if (unpaired) {   // longest run of columns with exactly one gap character that reaches the column, from either side
        novel.assign((size_t)A.n, 0); int run = 0;
        for (int c = 0; c < A.n; c++) { if (A.g[c] != '_' && A.s[c] != '_') run = 0; else if (!(A.g[c] == '_' && A.s[c] == '_')) run++; if (run > novel[(size_t)c]) novel[(size_t)c] = run; }
        run = 0;
        for (int c = A.n - 1; c >= 0; c--) { if (A.g[c] != '_' && A.s[c] != '_') run = 0; else if (!(A.g[c] == '_' && A.s[c] == '_')) run++; if (run > novel[(size_t)c]) novel[(size_t)c] = run; }
    }
    int idx = -1;
    for (int c = 0; c < A.n; c++) {
        if (A.level[c] == -1) {       // inserted base: joins the previous position's genotype; a leading "_" is dropped (:3328-3334)
            TY_REQUIRE(A.g[c] == '_' && A.s[c] != '_', "column without level is an insertion");
            idx++; TY_REQUIRE(idx < A.len, "read index inside the read");
            if (!all.empty()) { ExonObs& b = all.back(); if (b.genotype == "_") { TY_REQUIRE(b.qualities.empty(), "gap has no quality"); b.genotype.clear(); } b.genotype.push_back((char)A.s[c]); b.qualities.push_back((char)A.quals[idx]); }
            continue;
        }
        ExonObs e; e.level = A.level[c]; e.self = &A; e.mate = &M; e.dist = dist; e.mapq_pos = pcorrect_of(A.mq[c]); e.mq_char = A.mq[c]; if (unpaired) e.novel_gap = novel[(size_t)c];
        if (A.s[c] != '_') { idx++; TY_REQUIRE(idx < A.len, "read index inside the read"); e.genotype.assign(1, (char)A.s[c]); e.qualities.assign(1, (char)A.quals[idx]); }
        else e.genotype = "_";
        all.push_back(std::move(e));
    }
    // keep exon columns; positions must be consecutive inside one visit of the exon table (:3501-3561)
    int state = 0, last = -1;
    for (ExonObs& e : all) {
        int pos = -1;
        if (e.level >= L.lmin && e.level <= L.lmax) {   // exon columns are one or two consecutive level ranges
            int cum = 0;
            for (size_t x = 0, at = 0; x < L.exon_len.size(); at += (size_t)L.exon_len[x], x++) { int32_t f = L.col_level[at]; if (e.level >= f && e.level < f + L.exon_len[x]) { pos = cum + (e.level - f); break; } cum += L.exon_len[x]; }
        }
        if (pos >= 0) { if (state == 2) last = -1; TY_REQUIRE(last == -1 || pos == last + 1, "consecutive exon positions"); e.pos = (uint32_t)pos; last = pos; state = 1; out.push_back(std::move(e)); }
        else if (state == 1) state = 2;
    }
}

// removeDoublePositionsFromRead (HLATyper.cpp:4020-4083): per level keep the observation whose worst base quality is highest (first on ties)
std::vector<ExonObs> one_per_level(std::vector<ExonObs>& v) {
    std::map<int32_t, std::vector<size_t>> by_level; for (size_t i = 0; i < v.size(); i++) by_level[v[i].level].push_back(i);
    std::vector<ExonObs> out; out.reserve(by_level.size());
    for (auto& kv : by_level) {
        size_t best = kv.second[0]; unsigned char bq = 0; bool first = true;
        for (size_t i : kv.second) {
            const ExonObs& e = v[i]; TY_REQUIRE(e.genotype == "_" || !e.qualities.empty(), "observation has qualities");
            unsigned char q = 0; if (e.genotype != "_") { q = (unsigned char)e.qualities[0]; for (char ch : e.qualities) if ((unsigned char)ch < q) q = (unsigned char)ch; }
            if (first || q > bq) { best = i; bq = q; first = false; }
        }
        out.push_back(v[best]);
    }
    return out;
}

// ---- 31-mers of the reads (HLATyper.cpp:999-1024, 4210-4256): canonical form = lexicographic minimum of a k-mer and its reverse complement.
// Only presence is ever queried, and only for k-mers over ACGT (allele k-mers containing '*' are never looked up), so the table is a
// set of 2-bit packed canonical k-mers.
struct KmerSet {
    static const int K = 31;
    std::vector<uint64_t> slots; size_t used = 0;
    KmerSet() : slots(1 << 16, ~0ull) {}
    static int code(unsigned char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; } }
    static uint64_t mix(uint64_t x) { x ^= x >> 31; x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; return x; }
    void grow() { std::vector<uint64_t> old; old.swap(slots); slots.assign(old.size() * 2, ~0ull); used = 0; for (uint64_t v : old) if (v != ~0ull) insert(v); }
    void insert(uint64_t v) { if ((used + 1) * 2 > slots.size()) grow(); size_t m = slots.size() - 1, h = mix(v) & m; while (slots[h] != ~0ull) { if (slots[h] == v) return; h = (h + 1) & m; } slots[h] = v; used++; }
    bool has(uint64_t v) const { size_t m = slots.size() - 1, h = mix(v) & m; while (slots[h] != ~0ull) { if (slots[h] == v) return true; h = (h + 1) & m; } return false; }
    // calls f(canonical) for every K-mer of s[0..n) that is made of ACGT only, f_other() for the others
    template <class F, class G> static void scan(const unsigned char* s, size_t n, F f, G f_other) {
        if (n < (size_t)K) return;
        const uint64_t mask = (1ull << (2 * K)) - 1; uint64_t fw = 0, rv = 0; int valid = 0;
        for (size_t i = 0; i < n; i++) {
            int c = code(s[i]);
            if (c < 0) { valid = 0; fw = rv = 0; } else { fw = ((fw << 2) | (uint64_t)c) & mask; rv = (rv >> 2) | ((uint64_t)(3 - c) << (2 * (K - 1))); valid++; }
            if (i + 1 >= (size_t)K) { if (valid >= K) f(fw < rv ? fw : rv); else f_other(); }
        }
    }
};

struct KmerShards {     // a KmerSet split by hash, one shard per builder thread
    std::vector<KmerSet> shard;
    explicit KmerShards(unsigned n) : shard(n ? n : 1) {}
    unsigned owner(uint64_t k) const { return (unsigned)((KmerSet::mix(k) >> 40) % shard.size()); }
    bool has(uint64_t k) const { return shard[owner(k)].has(k); }
};

// HLALA_TYPING_PROFILE=1: wall time per phase of run_typing, summed over loci / threads, on stderr
struct PhaseClock {
    static constexpr int N = 10; std::atomic<long long> ns[N]; bool on;
    PhaseClock() : on(getenv("HLALA_TYPING_PROFILE") != nullptr) { for (auto& v : ns) v = 0; }
    struct Scope { PhaseClock& c; int i; std::chrono::steady_clock::time_point t0; Scope(PhaseClock& c_, int i_) : c(c_), i(i_), t0(std::chrono::steady_clock::now()) {}
                   ~Scope() { if (c.on) c.ns[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); } };
    void report() { if (!on) return; const char* nm[N] = {"mates+kmers", "projection", "first20", "counts", "pileup+files", "device input", "device", "rank+PP file", "QC+calls", "-"};
        for (int i = 0; i < N; i++) if (ns[i]) fprintf(stderr, "[typing-profile] %-14s %9.3f ms\n", nm[i], ns[i] / 1e6); }
};

double chi2_1_pvalue(double statistic) { return 1 - (statistic <= 0 ? 0 : erf(sqrt(statistic / 2.0))); }   // boost::math::cdf(chi_squared(1), x) == erf(sqrt(x/2))

} // namespace

void run_typing(TypingTables& T, const TypingReads& in, double is_mean, double is_sd, const std::string& out_dir, const std::string& g_dir,
                TypingDevice& dev, const TypingOptions& opt, std::vector<LocusCall>& calls) {
    calls.clear();
    PhaseClock clk; std::unique_ptr<PhaseClock::Scope> ph0(new PhaseClock::Scope(clk, 0));
    const size_t NP = in.n_pairs(); TY_REQUIRE(NP > 0, "rawPairedReads.size() > 0 || rawUnpairedReads.size() > 0");
    const bool LR = in.long_reads;       // long-read mode: entry p is the single read 2p; its "mate" 2p+1 is empty and stands for "no paired read" (-1 / "" in the reference's records)
    static const std::string no_name;
    std::vector<Mate> mates(2 * NP);
    // per-read statistics and the 31-mer set of the reads, on the host thread pool: reads are dealt in blocks; the k-mer set is sharded by hash so that every
    // thread owns one shard (each scans all reads and keeps the k-mers of its shard: scanning is cheap, inserting is what costs)
    unsigned pro_threads = opt.threads > 0 ? (unsigned)opt.threads : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("HLALA_HOST_THREADS")) pro_threads = std::min<unsigned>(pro_threads, (unsigned)std::max(1, atoi(e)));
    if (2 * NP < 4096) pro_threads = 1;
    auto on_pool = [&](const std::function<void(unsigned)>& f) {
        std::vector<std::exception_ptr> errs(pro_threads); std::vector<std::thread> th;
        auto body = [&](unsigned t) { try { f(t); } catch (...) { errs[t] = std::current_exception(); } };
        for (unsigned t = 1; t < pro_threads; t++) th.emplace_back(body, t);
        body(0); for (auto& x : th) x.join();
        for (auto& e : errs) if (e) std::rethrow_exception(e);
    };
    on_pool([&](unsigned t) {
        for (size_t r = (size_t)t * 2 * NP / pro_threads; r < (size_t)(t + 1) * 2 * NP / pro_threads; r++) {
            Mate& m = mates[r]; const int64_t c0 = in.col_off[r], b0 = in.base_off[r];
            m.n = (int)(in.col_off[r + 1] - c0); m.level = in.level.data() + c0; m.g = in.g.data() + c0; m.s = in.s.data() + c0; m.mq = in.mq.data() + c0;
            m.len = (int)(in.base_off[r + 1] - b0); m.bases = in.bases.data() + b0; m.quals = in.quals.data() + b0; m.reverse = in.reverse[r] != 0; m.mapQ = in.mapq[r]; m.name = &in.name[r / 2];
            if (LR && (r & 1)) { TY_REQUIRE(m.n == 0 && m.len == 0, "long-read entries have no second read"); m.name = &no_name; m.weighted_ok = -1; m.fraction_ok = -1; continue; }   // pairedRead_* = -1, pairedRead_ID = "" (HLATyper.cpp:3574-3581)
            mate_stats(m);
            for (int i = 0; i < m.len; i++) { unsigned char c = m.bases[i]; if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N' || c == 'a' || c == 'c' || c == 'g' || c == 't' || c == 'n' || c == '_' || c == '*')) throw std::runtime_error("typing: reverse complement of unknown character"); }
        }
    });
    // constants of HLATypeInference (HLATyper.cpp:944-946, 1032, 1551-1640)
    const double min_mapq = 0.0, min_pos_mapq = 0.7, min_weighted = 0.0, f20_min_prop = 0.1, unacc_min_frac = 0.2; const int F20N = 20, f20_limit = 2, high_cov = 100, unacc_min_cov = 30;

    KmerShards read_kmers(pro_threads);     // k-mers of the raw read == k-mers of the BAM-orientation read after canonicalisation
    on_pool([&](unsigned t) {
        KmerSet& mine = read_kmers.shard[t];
        for (size_t r = 0; r < 2 * NP; r++) KmerSet::scan(mates[r].bases, (size_t)mates[r].len, [&](uint64_t k) { if (read_kmers.owner(k) == t) mine.insert(k); }, []() {});
    });

    ph0.reset();
    // out_dir empty: compute only (ranks other than 0 of a multi-GPU run); every stream then goes to /dev/null
    auto target = [&](const std::string& name) { return out_dir.empty() ? std::string("/dev/null") : out_dir + "/" + name; };
    if (!out_dir.empty()) mkdir(out_dir.c_str(), 0777);
    {   // summaryStatistics.txt (HLATyper.cpp:1026-1125)
        int valid = 0, valid_dist = 0, perfect = 0, one_perfect = 0; std::vector<double> dists; double fsum = 0;
        const size_t NPP = LR ? 0 : NP, NU = LR ? NP : 0;      // paired / unpaired alignments
        int u_perfect = 0, u_long = 0; double u_fsum = 0;
        for (size_t p = 0; p < NU; p++) { const Mate& a = mates[2 * p]; if (a.n >= 1000) u_long++; if (a.fraction_ok == 1) u_perfect++; u_fsum += a.fraction_ok; }
        for (size_t p = 0; p < NPP; p++) { const Mate& a = mates[2 * p]; const Mate& b = mates[2 * p + 1];
            if (strands_ok(a, b)) { valid++; double d = level_distance(a, b); dists.push_back(d); if (fabs(d - is_mean) <= 5 * is_sd) valid_dist++; }
            perfect += (a.fraction_ok == 1) + (b.fraction_ok == 1); one_perfect += (a.fraction_ok == 1 || b.fraction_ok == 1); fsum += a.fraction_ok; fsum += b.fraction_ok; }
        std::sort(dists.begin(), dists.end()); double dsum = 0; for (double d : dists) dsum += d;
        double dmean = 0, dmed = 0; if (!dists.empty()) { dmean = dsum / (double)dists.size(); dmed = dists[dists.size() / 2]; }
        auto pct = [](double a, double b) { volatile double q = a / b; std::ostringstream o; o << (q * 100); return o.str(); };      // printPerc: 0 / 0 prints as the stream prints that NaN
        std::ofstream s(target("summaryStatistics.txt")); if (!s.is_open()) throw std::runtime_error("Cannot open " + target("summaryStatistics.txt") + " for writing");
        s << "\nRead alignment statistics:\n" << "\t - Total number (paired) alignments:                 " << NPP << "\n"
          << "\t\t - Alignment pairs with strands OK:                  " << valid << " (" << pct(valid, NPP) << "%)\n"
          << "\t\t - Alignment pairs with strands OK && distance OK:   " << valid_dist << " (" << pct(valid_dist, NPP) << "%)\n"
          << "\t\t - Alignment pairs with strands OK, mean distance:   " << dmean << "\n" << "\t\t - Alignment pairs with strands OK, median distance: " << dmed << "\n"
          << "\t\t - Alignment pairs, average fraction alignment OK:   " << (NPP > 0 ? fsum / (2.0 * (double)NPP) : 0.0) << "\n" << "\t\t - Alignment pairs, at least one alignment perfect:   " << one_perfect << "\n"
          << "\t\t - Single alignments, perfect (total):   " << perfect << " (" << NPP * 2 << ")\n" << "\t - Total number (unpaired) alignments:                 " << NU << "\n"
          << "\t\t - Alignment pairs, average fraction alignment OK:   " << (NU > 0 ? u_fsum / (double)NU : 0.0) << "\n" << "\t\t - Single alignments, perfect (total):   " << u_perfect << " (" << NU * 2 << ")\n" << "\t\t - Alignments with length >= " << 1000 << ":   " << u_long << "\n";
    }
    std::ofstream best(target("R1_bestguess.txt")), bestG(target("R1_bestguess_G.txt")), hist(target("histogram_matchesPerRead.txt"));
    if (!best.is_open() || !bestG.is_open() || !hist.is_open()) throw std::runtime_error("Cannot open output files in " + out_dir);
    const std::string unacc_field = "NColumns_UnaccountedAllele_fGT" + str(unacc_min_frac);
    best << "Locus\tChromosome\tAllele\tQ1\tQ2\tAverageCoverage\tCoverageFirstDecile\tMinimumCoverage\tproportionkMersCovered\tLocusAvgColumnError\t" << unacc_field << "\n";
    bestG << "Locus\tChromosome\tAllele\tQ1\tQ2\tAverageCoverage\tCoverageFirstDecile\tMinimumCoverage\tproportionkMersCovered\tLocusAvgColumnError\t" << unacc_field << "\tperfectG\n";
    hist << "Locus\tLevelValue\n";

    // which pairs pass the read-pair gates (HLATyper.cpp:1405-1410) does not depend on the locus
    std::vector<uint8_t> gate(NP, 0);
    for (size_t p = 0; p < NP; p++) { const Mate& a = mates[2 * p]; const Mate& b = mates[2 * p + 1]; TY_REQUIRE(a.mapQ >= 0 && a.mapQ <= 1, "mapQ in [0,1]");
        if (LR) gate[p] = a.mapQ >= min_mapq && a.n >= 1000;       // HLATyper.cpp:1475-1476 (minAlignmentLength_unpaired)
        else gate[p] = strands_ok(a, b) && fabs(level_distance(a, b) - is_mean) <= 5 * is_sd && a.mapQ >= min_mapq && a.weighted_ok >= min_weighted && b.weighted_ok >= min_weighted; }

    std::vector<std::array<std::string, 3>> hist_pair(NP);     // the three lines a gated pair contributes to every locus' histogram (without the locus name)
    std::string mapq_pos_str[256], qual_int_str[256];
    for (int q = 0; q < 256; q++) { qual_int_str[q] = str((int)(char)q); if (q == 0 || q >= 33) mapq_pos_str[q] = str(pcorrect_of((unsigned char)q)); }
    for (size_t p = 0; p < NP; p++) {
        Mate& a = mates[2 * p]; Mate& b = mates[2 * p + 1];
        for (int m = 0; m < 2; m++) { Mate& x = m ? b : a; const Mate& y = m ? a : b;
            x.hist_base = "base" + str(x.weighted_ok) + "\n";
            if (LR && m == 1) continue;
            x.pile_mid = ") [pairsDistance " + (LR ? std::string("-1") : str((double)level_distance(x, y))) + " | alignmentLength " + str(x.nongap_cols) + " | ";
            x.pile_tail = " | " + str(x.mapQ) + " " + str(x.mapQ) + " | " + str(x.weighted_ok) + " " + str(y.weighted_ok) + " | " + *x.name + " " + *y.name + "]"; }
        if (gate[p] && !LR) hist_pair[p] = {"\tread" + str(a.weighted_ok) + "\n", "\tread" + str(b.weighted_ok) + "\n", "\treadPair" + str((a.weighted_ok + b.weighted_ok) / 2.0) + "\n"};
    }

    // The loci are independent up to the order of their lines in the shared files and the order of the device stage (with several ranks
    // the per-locus all-reduce must be issued in the same order everywhere): host work of different loci runs on a thread pool, the device
    // stage is entered in locus order, and the shared files are assembled in locus order afterwards.
    struct LocusOut { LocusCall call; std::string hist, best, bestG; std::exception_ptr err; };
    const size_t NL = T.loci.size(); std::vector<LocusOut> outs(NL);
    T.load_G(g_dir);
    std::mutex dev_mu; std::condition_variable dev_cv; size_t dev_turn = 0;
    auto take_turn = [&](size_t li) { std::unique_lock<std::mutex> lk(dev_mu); dev_cv.wait(lk, [&]() { return dev_turn == li; }); };
    auto pass_turn = [&](size_t li) { { std::lock_guard<std::mutex> lk(dev_mu); if (dev_turn == li) dev_turn = li + 1; } dev_cv.notify_all(); };
    std::vector<uint8_t> entered_device(NL, 0);
    auto process_locus = [&](size_t li) {
        TypingLocus& L = T.loci[li]; LocusCall& call = outs[li].call;
        std::string hist; std::ostringstream best, bestG;   // this locus' lines of histogram_matchesPerRead.txt / R1_bestguess.txt / R1_bestguess_G.txt
        call.locus = L.name; const int32_t C = L.C(), P = L.P(); call.C = C;
        if (const char* e = getenv("HLALA_TEST_FAIL_LOCUS")) if ((size_t)atoi(e) == li) throw std::runtime_error("typing: locus " + L.name + " failed (HLALA_TEST_FAIL_LOCUS test hook)");
        // ---- exon observations per read pair
        std::unique_ptr<PhaseClock::Scope> ph(new PhaseClock::Scope(clk, 1));
        std::vector<std::vector<ExonObs>> reads;
        for (size_t p = 0; p < NP; p++) {
            const Mate& a = mates[2 * p]; const Mate& b = mates[2 * p + 1];
            std::vector<ExonObs> obs;
            if (LR) { project_read(a, b, L, obs, true); if (!gate[p]) continue; if (!obs.empty()) reads.push_back(std::move(obs)); continue; }     // HLATyper.cpp:1467-1495: no double positions inside one read
            project_read(a, b, L, obs); project_read(b, a, L, obs);    // both are validated even if the pair is gated out
            if (!gate[p]) continue;
            if (!obs.empty()) reads.push_back(one_per_level(obs));
            for (const std::string& piece : hist_pair[p]) { hist += L.name; hist += piece; }
        }
        const size_t R = reads.size(); call.R = (int32_t)R;
        // ---- "first 20" filter (HLATyper.cpp:1551-1640)
        ph.reset(new PhaseClock::Scope(clk, 2));
        std::set<std::string> ignored_reads; std::map<uint32_t, std::set<std::string>> ignored_alleles;
        if (!LR) {     // filterFirst20 && longReadsMode.length() == 0 (HLATyper.cpp:1509)
            std::map<uint32_t, std::vector<std::string>> al; std::map<uint32_t, std::vector<double>> wq; std::map<uint32_t, std::vector<uint32_t>> rd; std::map<uint32_t, int> robust;
            for (uint32_t r = 0; r < R; r++) for (const ExonObs& e : reads[r]) { TY_REQUIRE(e.mapq_pos >= 0 && e.mapq_pos <= 1, "mapQ_position in [0,1]"); if (e.mapq_pos < min_pos_mapq) continue;
                al[e.pos].push_back(e.genotype); wq[e.pos].push_back((e.self->weighted_ok + e.mate->weighted_ok) / 2.0); rd[e.pos].push_back(r); }
            for (auto& kv : al) {
                const std::vector<std::string>& A = kv.second; const int n = (int)A.size(); if (n < F20N) continue;
                const std::vector<double>& W = wq.at(kv.first); std::vector<unsigned> order((size_t)n); for (int i = 0; i < n; i++) order[i] = (unsigned)i;
                std::sort(order.begin(), order.end(), [&](unsigned x, unsigned y) { return W.at(x) < W.at(y); }); std::reverse(order.begin(), order.end());   // same call sequence as the reference: ties fall where std::sort puts them
                std::map<std::string, int> top; for (int i = 0; i < F20N; i++) top[A[order[i]]]++;
                std::set<std::string> kicked;
                for (int i = 0; i < n; i++) { auto it = top.find(A[i]); double prop = (double)(it == top.end() ? 0 : it->second) / (double)true;   // the reference divides by the bool switch (HLATyper.cpp:1583)
                    if (prop < f20_min_prop) { kicked.insert(A[i]); ignored_alleles[kv.first].insert(A[i]); } }
                std::map<std::string, int> times; for (int i = 0; i < n; i++) if (kicked.count(A[i])) times[A[i]]++;
                for (int i = 0; i < n; i++) { auto it = times.find(A[i]); if (it != times.end() && it->second >= 2) robust[rd.at(kv.first)[i]]++; }
            }
            for (auto& kv : robust) if (kv.second > f20_limit) { const ExonObs& e = reads.at(kv.first).at(0); ignored_reads.insert(*e.self->name); ignored_reads.insert(*e.mate->name); }
        }
        auto used = [&](const ExonObs& e) {
            if (e.mapq_pos < min_pos_mapq) return false;
            auto it = ignored_alleles.find(e.pos); if (it != ignored_alleles.end() && it->second.count(e.genotype)) return false;
            return ignored_reads.count(*e.self->name) == 0;
        };
        // ---- allele counts per position, by strand / by mate (HLATyper.cpp:1663-1875); the frequency filters are off in short-read mode
        ph.reset(new PhaseClock::Scope(clk, 3));
        std::map<uint32_t, std::map<std::string, double>> min_strand_freq, read1_freq; std::map<uint32_t, std::map<std::string, int>> counts_high_cov;
        {
            struct Cnt { int n = 0, fwd = 0, rev = 0, first = 0; }; std::map<uint32_t, std::map<std::string, Cnt>> cnt;
            for (uint32_t r = 0; r < R; r++) for (const ExonObs& e : reads[r]) { if (!used(e)) continue; Cnt& c = cnt[e.pos][e.genotype]; c.n++; if (e.self->reverse) c.rev++; else c.fwd++;
                if (LR) c.first++; /* paired mode: fromFirstRead is false for both mates (processBAM.cpp:3545-3548 overwrites the flag); alignOneLongRead sets it (:3756) */ }
            // long-read mode: highCoverage_filter_alleles with minCoverage 1 and minAlleleFreq 0.15 (HLATyper.cpp:944-946, 1806-1828), strand filter for alleles seen >= 100 times (:1847-1856)
            const int cov_min = LR ? 1 : high_cov; const double min_af = 0.15, min_strand = 0.1; const int strand_min_cov = 100;
            for (auto& pk : cnt) { int tot = 0; for (auto& a : pk.second) tot += a.second.n;
                for (auto& a : pk.second) {
                    if (tot >= cov_min) { if (LR && (double)a.second.n / (double)tot < min_af) ignored_alleles[pk.first].insert(a.first); else counts_high_cov[pk.first][a.first] = a.second.n; }
                }
                for (auto& a : pk.second) { int t = a.second.fwd + a.second.rev; const double msf = (double)std::min(a.second.fwd, a.second.rev) / (double)t; min_strand_freq[pk.first][a.first] = msf; read1_freq[pk.first][a.first] = (double)a.second.first / (double)t;
                    if (LR && t >= strand_min_cov && msf < min_strand) ignored_alleles[pk.first].insert(a.first); } }
        }
        // ---- pile-up (HLATyper.cpp:1877-2037)
        ph.reset(new PhaseClock::Scope(clk, 4));
        std::map<int, std::map<int, std::vector<const ExonObs*>>> pile; std::set<std::string> utilized;
        for (uint32_t r = 0; r < R; r++) for (const ExonObs& e : reads[r]) { if (!used(e)) continue; if (LR && e.novel_gap >= 2) continue;     // HLATyper.cpp:1918
            pile[L.col_exon[e.pos]][L.col_exonpos[e.pos]].push_back(&e); hist += L.name; hist += '\t'; hist += e.self->hist_base; }
        {
            // every observation is "<genotype> (<qualities>) [pairsDistance d | alignmentLength n | mapQ_position | mapQ mapQ | w w | name name]"; all but
            // the genotype, the qualities and the per-column mapQ are per-read text prepared once (Mate::pile_mid / pile_tail)
            std::string text; text.reserve((size_t)1 << 22);
            std::unordered_set<const Mate*> utilized_mates;
            for (auto& ex : pile) { const int exon = ex.first, len = L.exon_len.at((size_t)exon);
                for (int ep = 0; ep < len; ep++) {
                    auto it = ex.second.find(ep);
                    put_num(text, (long long)exon); text += '\t'; put_num(text, (long long)ep); text += '\t';
                    if (it == ex.second.end()) { text += "0\n"; continue; }
                    const std::vector<const ExonObs*>& pu = it->second; std::map<std::string, std::pair<long long, long long>> per_allele; uint32_t this_pos = pu[0]->pos;
                    put_num(text, (long long)pu.size()); text += '\t';
                    bool first_part = true;
                    for (const ExonObs* e : pu) { TY_REQUIRE(e->pos == this_pos, "pile-up column holds one exon position");
                        if (!first_part) text += ", "; first_part = false;
                        text += e->genotype; text += " (";
                        for (size_t qi = 0; qi < e->qualities.size(); qi++) { if (qi) text += ", "; text += qual_int_str[(unsigned char)e->qualities[qi]]; }
                        text += e->self->pile_mid; text += mapq_pos_str[e->mq_char]; text += e->self->pile_tail;
                        utilized_mates.insert(e->self); auto& pa = per_allele[e->genotype]; pa.first += e->self->nongap_cols; pa.second++; }
                    text += '\t';
                    for (auto& a : per_allele) { text += a.first; text += 'x'; put_num(text, a.second.second); text += '['; put_num(text, (double)a.second.first / (double)a.second.second); text += ';';
                        put_num(text, min_strand_freq.at(this_pos).at(a.first)); text += ';'; put_num(text, read1_freq.at(this_pos).at(a.first)); text += ']'; }
                    text += '\n';
                } }
            for (const Mate* m : utilized_mates) utilized.insert(*m->name);
            std::ofstream ps(target("R1_pileup_" + L.name + ".txt")); ps.write(text.data(), (std::streamsize)text.size());
        }
        { std::ofstream rs(target("R1_readIDs_" + L.name + ".txt")); for (const std::string& id : utilized) rs << id << "\n"; }

        // ---- the two GPU stages: per-read x cluster log-likelihoods, allele-pair sums
        ph.reset(new PhaseClock::Scope(clk, 5));
        LocusDeviceInput di; di.C = C; di.P = P; di.R = (int32_t)R; di.long_reads = LR; di.cluster_seq = &L.cluster_seq; di.rec_off.assign(1, 0); long long bases_used = 0;
        for (uint32_t r = 0; r < R; r++) {
            for (const ExonObs& e : reads[r]) { if (!used(e)) continue;
                const bool gap = e.genotype == "_"; if (!gap) { TY_REQUIRE(e.genotype.find('_') == std::string::npos, "no gap inside an insertion"); TY_REQUIRE(!e.qualities.empty(), "quality present"); double pc = pcorrect_of((unsigned char)e.qualities[0]); TY_REQUIRE(pc >= 0 && pc <= 1, "pCorrect in [0,1]"); }
                di.rec_pos.push_back((int16_t)e.pos); di.rec_c0.push_back(gap ? (uint8_t)'_' : (uint8_t)e.genotype[0]); di.rec_q0.push_back(gap ? 0 : (uint8_t)e.qualities[0]); di.rec_glen.push_back((uint16_t)std::min<size_t>(e.genotype.size(), 65535)); bases_used++; }
            di.rec_off.push_back((int32_t)di.rec_pos.size());
        }
        ph.reset(new PhaseClock::Scope(clk, 6));
        take_turn(li); entered_device[li] = 1; try { dev.run_locus(di, opt.keep_read_ll, call.dev); } catch (...) { pass_turn(li); throw; } pass_turn(li);
        ph.reset(new PhaseClock::Scope(clk, 7));
        const std::vector<double>& LLs = call.dev.pair_ll; const std::vector<double>& Mavg = call.dev.pair_mavg; const std::vector<double>& Mmin = call.dev.pair_mmin;
        const size_t NPAIR = (size_t)C * ((size_t)C + 1) / 2; TY_REQUIRE(LLs.size() == NPAIR && Mavg.size() == NPAIR && Mmin.size() == NPAIR, "pair arrays complete");
        std::vector<std::pair<uint32_t, uint32_t>> ids; ids.reserve(NPAIR); for (uint32_t c1 = 0; c1 < (uint32_t)C; c1++) for (uint32_t c2 = c1; c2 < (uint32_t)C; c2++) ids.push_back({c1, c2});
        // ---- normalise, rank, call (HLATyper.cpp:2366-2538)
        std::vector<size_t> order(NPAIR); for (size_t i = 0; i < NPAIR; i++) order[i] = i;
        { const double* ll = LLs.data(); const double* ma = Mavg.data();      // same comparator, same std::sort: the order of exact ties is the reference's
          std::sort(order.begin(), order.end(), [ll, ma](unsigned a, unsigned b) { if (ll[a] == ll[b]) return ma[b] < ma[a]; return ll[a] < ll[b]; }); std::reverse(order.begin(), order.end()); }
        size_t imax = 0; for (size_t i = 1; i < NPAIR; i++) if (LLs[i] > LLs[imax]) imax = i; const double ll_max = LLs[imax];
        double psum = 0; for (double v : LLs) psum += exp(v - ll_max);
        std::vector<double> Pn(NPAIR); for (size_t i = 0; i < NPAIR; i++) { if (psum > 0) { Pn[i] = exp(LLs[i] - ll_max) / psum; TY_REQUIRE(Pn[i] >= 0 && Pn[i] <= 1, "P_normalized in [0,1]"); } else Pn[i] = 1.0 / (double)NPAIR; }
        std::vector<std::string> member_str((size_t)C); for (int32_t c = 0; c < C; c++) member_str[(size_t)c] = join_with(L.cluster_members[(size_t)c], ";");
        auto members = [&](uint32_t c) -> const std::string& { return member_str[c]; };
        std::map<int, double> marginal;
        { std::ofstream ap(target("R1_PP_" + L.name + "_pairs.txt")); ap << "ClusterID\tP\tLL\tMismatches_avg\n";
          std::vector<double> marg((size_t)C, 0.0);
          // the lines are formatted in slices of the ranked list (helper threads when the table is large and cores are free: 500 k lines per locus at 1000 alleles),
          // written in order; the marginals are summed in rank order by this thread (the order of the additions is part of the result)
          const unsigned hw = std::max(1u, std::thread::hardware_concurrency()); unsigned helpers = NPAIR >= 100000 ? std::min(6u, std::max(1u, hw / (unsigned)std::max<size_t>(NL, 1))) : 1u;
          if (const char* e = getenv("HLALA_HOST_THREADS")) helpers = std::min<unsigned>(helpers, std::max(1u, (unsigned)atoi(e) / (unsigned)std::max<size_t>(NL, 1)));
          std::vector<std::string> part(helpers);
          auto format = [&](unsigned h) { std::string& text = part[h]; const size_t k0 = NPAIR * h / helpers, k1 = NPAIR * (h + 1) / helpers; text.reserve((k1 - k0) * 72);
              for (size_t k = k0; k < k1; k++) { const size_t i = order[k];
                  text += members(ids[i].first); text += '/'; text += members(ids[i].second); text += '\t'; put_num(text, Pn[i]); text += '\t'; put_num(text, LLs[i]); text += '\t'; put_num(text, Mavg[i]); text += '\n'; } };
          { std::vector<std::thread> th; for (unsigned h = 1; h < helpers; h++) th.emplace_back(format, h);
            for (size_t k = 0; k < NPAIR; k++) { const size_t i = order[k]; marg[ids[i].first] += Pn[i]; if (ids[i].second != ids[i].first) marg[ids[i].second] += Pn[i]; }
            format(0); for (auto& t : th) t.join(); }
          for (const std::string& text : part) ap.write(text.data(), (std::streamsize)text.size());
          for (int32_t c = 0; c < C; c++) marginal[c] = marg[(size_t)c]; }      // every cluster is part of a pair: the keys the reference's map holds
        auto first_max = [](const std::map<int, double>& m) { double mx = 0; int at = 0; bool first = true; for (auto& kv : m) if (first || kv.second > mx) { mx = kv.second; at = kv.first; first = false; } return std::make_pair(mx, at); };   // Utilities::findIntMapMaxP_nonCritical
        const std::pair<double, int> b1 = first_max(marginal);
        std::map<int, double> partner_p, partner_mm;
        for (size_t i = 0; i < NPAIR; i++) { if ((int)ids[i].first == b1.second) { partner_p[(int)ids[i].second] = Pn[i]; partner_mm[(int)ids[i].second] = Mmin[i]; } else if ((int)ids[i].second == b1.second) { partner_p[(int)ids[i].first] = Pn[i]; partner_mm[(int)ids[i].first] = Mmin[i]; } }
        const std::pair<double, int> p2 = first_max(partner_p); std::map<int, double> tie_break; for (auto& kv : partner_p) if (kv.second == p2.first) tie_break[kv.first] = -1 * partner_mm.at(kv.first);
        const std::pair<double, int> b2 = first_max(tie_break);
        call.call1 = members((uint32_t)b1.second); call.call2 = members((uint32_t)b2.second); call.q1 = b1.first; call.q2 = p2.first;
        // ---- QC (HLATyper.cpp:2543-2759, 4258-4320)
        ph.reset(new PhaseClock::Scope(clk, 8));
        int total_columns = 0; for (int l : L.exon_len) total_columns += l;
        const double locus_cov = (double)bases_used / (double)total_columns;
        std::vector<double> pos_cov; for (int32_t i = 0; i < P; i++) pos_cov.push_back((double)pile[L.col_exon[i]][L.col_exonpos[i]].size());
        std::sort(pos_cov.begin(), pos_cov.end(), std::less<int>());      // the reference compares the doubles as ints
        const std::string& s1 = L.cluster_seq[(size_t)b1.second]; const std::string& s2 = L.cluster_seq[(size_t)b2.second];
        size_t all_tot = 0, all_inc = 0; std::vector<int> col_tot, col_inc; int n_unacc = 0; std::vector<std::string> ex1, ex2;
        for (int32_t i = 0; i < P; i++) {
            if (i == 0 || L.col_exon[i] != L.col_exon[i - 1]) { ex1.emplace_back(); ex2.emplace_back(); }
            const std::string u1(1, s1[i]), u2(1, s2[i]); ex1.back().push_back(s1[i]); ex2.back().push_back(s2[i]);
            int tot = 0, inc = 0; for (const ExonObs* e : pile.at(L.col_exon[i]).at(L.col_exonpos[i])) { tot++; if (e->genotype != u1 && e->genotype != u2) inc++; }
            all_tot += tot; all_inc += inc; col_tot.push_back(tot); col_inc.push_back(inc);
            auto it = counts_high_cov.find((uint32_t)i);
            if (it != counts_high_cov.end()) { int cov = 0; for (auto& a : it->second) cov += a.second;
                if (cov >= unacc_min_cov) for (auto& a : it->second) { if (a.first == u1 || a.first == u2) continue; if ((double)a.second / (double)cov >= unacc_min_frac) n_unacc++; } }
        }
        auto kmer_presence = [&](const std::vector<std::string>& exons) -> double { int tot = 0, present = 0;
            for (const std::string& e : exons) { std::string ng; for (char ch : e) if (ch != '_') ng.push_back(ch);
                KmerSet::scan((const unsigned char*)ng.data(), ng.size(), [&](uint64_t k) { tot++; if (read_kmers.has(k)) present++; }, [&]() { tot++; }); }
            return tot == 0 ? -1.0 : (double)present / (double)tot; };
        const double k1 = kmer_presence(ex1), k2 = kmer_presence(ex2);
        const double avg_err = all_tot > 0 ? (double)all_inc / (double)all_tot : 0;
        { std::ofstream ce(target("R1_columnIncompatibilities_" + L.name + ".txt")); ce << "Column\tCoverage\tExpectedIncompatible\tObservedIncompatible\tp\n";
          for (int32_t col = 0; col < P; col++) { const int cov = col_tot[col], obs = col_inc[col]; const double expd = avg_err * cov; double pv = 1;
              if (obs > expd) { double o0 = cov - obs, o1 = obs, e0 = cov - expd, e1 = expd; TY_REQUIRE(e0 > 0 && e1 > 0, "expected counts > 0"); double st = 0; st += pow(o0 - e0, 2) / e0; st += pow(o1 - e1, 2) / e1; pv = chi2_1_pvalue(st); TY_REQUIRE(pv >= 0 && pv <= 1, "p in [0,1]"); }
              ce << col << "\t" << cov << "\t" << expd << "\t" << obs << "\t" << pv << "\n"; } }
        const double first_decile = pos_cov.at((size_t)((double)pos_cov.size() / 10.0)), min_cov = pos_cov.at(0);
        auto row = [&](std::ostream& o, int chrom, const std::string& allele, double qa) { o << L.name << "\t" << chrom << "\t" << allele << "\t" << qa << "\t" << b2.first << "\t" << locus_cov << "\t" << first_decile << "\t" << min_cov << "\t" << (chrom == 1 ? k1 : k2) << "\t" << avg_err << "\t" << n_unacc; };
        row(best, 1, call.call1, b1.first); best << "\n"; row(best, 2, call.call2, p2.first); best << "\n" << std::flush;
        if (T.G_loci.count(L.name)) {   // HLATyper.cpp:4095-4148
            auto to_G = [&](const std::vector<std::string>& alleles, bool& perfect) -> std::string {
                std::map<std::string, int> groups;
                for (const std::string& a : alleles) { if (split_on(a, "*").size() != 2) throw std::runtime_error("Weird allele: " + a); auto it = T.alleles_to_G.find(a); if (it != T.alleles_to_G.end()) groups[it->second]++; }
                if (groups.empty()) { perfect = false; return join_with(alleles, ";"); }
                if (groups.size() == 1) { perfect = true; return groups.begin()->first; }
                perfect = false; std::vector<std::string> keys; for (auto& kv : groups) keys.push_back(kv.first);
                std::sort(keys.begin(), keys.end(), [&](const std::string& x, const std::string& y) { return groups.at(x) < groups.at(y); }); std::reverse(keys.begin(), keys.end());
                return keys.at(0); };
            bool p1 = false, p2g = false; const std::string g1 = to_G(L.cluster_members[(size_t)b1.second], p1), g2 = to_G(L.cluster_members[(size_t)b2.second], p2g);
            row(bestG, 1, g1, b1.first); bestG << "\t" << p1 << "\n"; row(bestG, 2, g2, p2.first); bestG << "\t" << p2g << "\n" << std::flush;
        }
        if (!opt.keep_read_ll) { call.dev.LL.clear(); call.dev.LL.shrink_to_fit(); call.dev.mism.clear(); call.dev.mism.shrink_to_fit(); }
        outs[li].hist = std::move(hist); outs[li].best = best.str(); outs[li].bestG = bestG.str();
        ph.reset();
    };
    // a locus that fails on this rank before its device stage still enters the per-locus collective (poisoned), so that all ranks fail together instead of
    // the others waiting in the all-reduce for ever
    auto process = [&](size_t li) {
        try { process_locus(li); }
        catch (...) { if (!entered_device[li]) { take_turn(li); try { dev.abort_locus(T.loci[li].C()); } catch (...) {} pass_turn(li); } throw; }
    };
    {
        std::atomic<size_t> next(0);
        auto worker = [&]() { for (;;) { const size_t li = next.fetch_add(1); if (li >= NL) break;
            try { process(li); } catch (...) { outs[li].err = std::current_exception(); }
            { std::unique_lock<std::mutex> lk(dev_mu); dev_cv.wait(lk, [&]() { return dev_turn >= li; }); if (dev_turn == li) dev_turn = li + 1; }   // a locus that failed before its device stage still hands the turn on
            dev_cv.notify_all(); } };
        unsigned nt = opt.threads > 0 ? (unsigned)opt.threads : std::max(1u, std::min<unsigned>((unsigned)NL, std::thread::hardware_concurrency()));
        if (const char* e = getenv("HLALA_HOST_THREADS")) nt = std::min<unsigned>(nt, (unsigned)std::max(1, atoi(e)));     // several ranks on one node
        if (const char* e = getenv("HLALA_TYPING_THREADS")) nt = (unsigned)std::max(1, atoi(e));
        std::vector<std::thread> th; for (unsigned t = 1; t < nt; t++) th.emplace_back(worker);
        worker(); for (auto& t : th) t.join();
    }
    std::vector<std::string> locus_names;
    for (size_t li = 0; li < NL; li++) {
        if (outs[li].err) std::rethrow_exception(outs[li].err);
        locus_names.push_back(T.loci[li].name); hist << outs[li].hist; best << outs[li].best; bestG << outs[li].bestG;
        calls.push_back(std::move(outs[li].call));
    }
    best << std::flush; bestG << std::flush;
    clk.report();
    { std::ofstream ps(target("R1_parameters.txt")); ps << "Loci = " << join_with(locus_names, ",") << "\n" << "veryConservativeReadLikelihoods = " << true << "\n"; }
}

} // namespace hlala
