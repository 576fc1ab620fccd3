// TEST INFRASTRUCTURE — not part of the product (see hlala_oracle.cpp). CPU restatement of the HLA typing stage.
//
// Restates, on the flat alignment arrays every implementation exports (level / graph char / read char / per-column
// mapQ char per read, strand, chain mapQ):
//   gene filter                                  mapper/processBAM.cpp:2427-2446, hla/HLATyper.cpp:259
//   HLATyper ctor: segments, gene boundaries      hla/HLATyper.cpp:36-256
//   HLATypeInference                              hla/HLATyper.cpp:933-2810
//     exon tables, clustering                     :1177-1372
//     oneReadAlignment_2_exonPositions_paired     :3192-3565
//     alignmentFractionOK / WeightedOKFraction    :3082, :3933
//     removeDoublePositionsFromRead               :4020
//     first-20 filter, frequency/strand tables    :1509-1875
//     pile-up + files                             :1877-2037
//     per-read x cluster log-likelihoods          :2049-2277
//     allele-pair log-likelihoods (logAvg)        :2280-2364, Utilities.cpp:1368
//     normalise, rank, call, QC, G translation    :2366-2790, :4095-4320
// The files it writes are compared byte-for-byte with the ones the unmodified reference writes (tests/test_typing_oracle.py);
// raw doubles are exported besides them so the CUDA path can be compared tighter than the files' 6 significant digits.
// Short-read (paired) mode only, like the rest of this repository.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>

namespace {

#define TREQUIRE(c, what) do { if (!(c)) throw std::runtime_error(std::string("reference assertion would fail: ") + what); } while (0)

std::vector<std::string> split(const std::string& s, const std::string& d) {   // Utilities::split (Utilities.cpp:610): keeps empty fields, incl. the trailing one
    std::vector<std::string> r; if (s.empty()) return r;
    size_t pos = 0;
    for (;;) { size_t q = s.find(d, pos); if (q == std::string::npos) { r.push_back(s.substr(pos)); break; } r.push_back(s.substr(pos, q - pos)); pos = q + d.size(); }
    return r;
}
std::string join(const std::vector<std::string>& p, const std::string& d) { std::string r; for (size_t i = 0; i < p.size(); i++) { if (i) r += d; r += p[i]; } return r; }
void erase_nl(std::string& s) { while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); }
std::string itos(long long v) { std::stringstream ss; ss << v; return ss.str(); }
std::string dtos(double v) { std::stringstream ss; ss << v; return ss.str(); }
double phred_to_pcorrect(unsigned char q) {   // Utilities.cpp:357-377
    if (q == 0) return -1;
    int ph = (int)q - 33; TREQUIRE(ph >= 0, "phred >= 0");
    double l = (double)ph / (double)-10; double w = exp(log(10) * l); return 1 - w;
}
char rc_char(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'n': return 'n'; case '_': return '_'; case '*': return '*'; default: throw std::runtime_error("reverse complement of unknown character"); } }
std::string rc_seq(const std::string& s) { std::string r(s.rbegin(), s.rend()); for (char& c : r) c = rc_char(c); return r; }
double log_avg(double a, double b) { if (a > b) return log(0.5) + (log(1 + exp(b - a)) + a); return log(0.5) + (log(1 + exp(a - b)) + b); }
std::vector<std::string> kmers_of(const std::string& s, int k) { std::vector<std::string> r; if ((int)s.size() >= k) for (size_t i = 0; i + k <= s.size(); i++) r.push_back(s.substr(i, k)); return r; }
std::string kmer_canon(const std::string& k) {   // HLATyper.cpp:4210-4256
    char f = k.at(0), lr = rc_char(k.at(k.size() - 1));
    if (f < lr) return k;
    std::string r = rc_seq(k); return r < k ? r : k;
}

struct Aln {   // one read's chosen alignment (verboseSeedChain) + its raw data in BAM (= alignment) orientation
    std::vector<int> level; std::string g, s, mapq, qual, seq_bam, name; bool reverse = false, from_first = false; double mapQ = 0;
    int first_level() const { for (int l : level) if (l != -1) return l; return -1; }
    int last_level() const { for (size_t i = level.size(); i-- > 0;) if (level[i] != -1) return level[i]; return -1; }
};

bool strands_valid(const Aln& a, const Aln& b) {   // alignerBase.cpp:213-245
    if (a.first_level() != -1 && b.first_level() != -1 && a.reverse != b.reverse) return !a.reverse ? (a.first_level() < b.first_level()) : (a.last_level() > b.last_level());
    return false;
}
int pair_distance(const Aln& a, const Aln& b) { if (a.first_level() < b.first_level()) return b.first_level() - a.last_level() - 1; return a.first_level() - b.last_level() - 1; }   // alignerBase.cpp:247-288

double fraction_ok(const Aln& r) { int ok = 0, chk = 0; for (size_t i = 0; i < r.g.size(); i++) { if (r.g[i] == '_' && r.s[i] == '_') continue; chk++; if (r.g[i] == r.s[i]) ok++; } return double(ok) / double(chk); }
double weighted_ok(const Aln& a) {   // HLATyper.cpp:3933-4018; the strand flip of index and quality string cancel, qualities are read in alignment orientation
    int idx = -1; int total = 0; double w = 0;
    for (size_t c = 0; c < a.s.size(); c++) {
        if (a.s[c] != '_') {
            idx++; TREQUIRE(idx < (int)a.qual.size(), "read index in range"); TREQUIRE(a.seq_bam[idx] == a.s[c], "underlyingReadCharacter == sequenceCharacter");
            if (a.g[c] == '_') { total++; w++; }
            else { double p = phred_to_pcorrect((unsigned char)a.qual[idx]); TREQUIRE(p >= 0 && p <= 1, "pCorrect in [0,1]"); if (a.s[c] != a.g[c]) { w += p; total++; } }
        } else if (a.g[c] != '_') { total++; w++; }
    }
    double len = (double)a.seq_bam.size(); TREQUIRE(total >= w, "totalMismatches >= weightedMismatches");
    return 1.0 - (w / len);
}

struct ExonPos {   // hla/oneExonPosition.h:15-46 (fields that are read anywhere)
    unsigned int positionInExon = 0; int graphLevel = -1; std::string genotype, qualities, thisRead_ID, pairedRead_ID;
    double thisW = 0, pairedW = 0, strands_distance = 0, mapQ = 0, mapQ_position = 0; int alnCols = 0, novelGap = 0; bool reverse = false, fromFirstRead = false;
};

void project(const Aln& A, const Aln& P, int lmin, int lmax, const std::map<int, unsigned int>& level2pos, std::vector<ExonPos>& out) {   // HLATyper.cpp:3192-3565
    int fl = A.first_level(), ll = A.last_level(); TREQUIRE(fl <= ll, "alignment_firstLevel <= alignment_lastLevel");
    double thisW = weighted_ok(A), pairedW = weighted_ok(P); double dist = pair_distance(A, P);
    bool overlap = (fl >= lmin && fl <= lmax) || (ll >= lmin && ll <= lmax) || (lmin >= fl && lmin <= ll) || (lmax >= fl && lmax <= ll);
    if (!overlap) return;
    std::vector<ExonPos> rec; int ncol = (int)A.s.size();
    int oneNonGap = 0; for (int c = 0; c < ncol; c++) if (A.s[c] != '_') oneNonGap++;   // the reference tests the read character twice (:3235)
    std::vector<int> novel(ncol, 0);
    { int run = 0; for (int c = 0; c < ncol; c++) { if (A.g[c] != '_' && A.s[c] != '_') run = 0; else if (!(A.g[c] == '_' && A.s[c] == '_')) run++; if (run > novel[c]) novel[c] = run; }
      run = 0; for (int c = ncol - 1; c >= 0; c--) { if (A.g[c] != '_' && A.s[c] != '_') run = 0; else if (!(A.g[c] == '_' && A.s[c] == '_')) run++; if (run > novel[c]) novel[c] = run; } }
    int idx = -1;
    for (int c = 0; c < ncol; c++) {
        double mq = phred_to_pcorrect((unsigned char)A.mapq[c]);
        auto mk = [&](const std::string& gt, const std::string& q) { ExonPos e; e.graphLevel = A.level[c]; e.genotype = gt; e.qualities = q; e.thisRead_ID = A.name; e.pairedRead_ID = P.name; e.thisW = thisW; e.pairedW = pairedW;
            e.strands_distance = dist; e.mapQ = A.mapQ; e.mapQ_position = mq; e.alnCols = oneNonGap; e.novelGap = novel[c]; e.reverse = A.reverse; e.fromFirstRead = A.from_first; rec.push_back(e); };
        if (A.level[c] == -1) {
            TREQUIRE(A.g[c] == '_' && A.s[c] != '_', "insertion column");
            idx++; TREQUIRE(idx < (int)A.qual.size(), "read index in range");
            if (!rec.empty()) {
                ExonPos& b = rec.back(); b.genotype.push_back(A.s[c]); b.qualities.push_back(A.qual[idx]);
                if (b.genotype.size() != b.qualities.size()) { TREQUIRE(b.genotype.size() == b.qualities.size() + 1 && b.genotype[0] == '_', "gap + insertion"); b.genotype = b.genotype.substr(1); }
            }
        } else if (A.s[c] != '_') { idx++; TREQUIRE(idx < (int)A.qual.size(), "read index in range"); mk(std::string(1, A.s[c]), std::string(1, A.qual[idx])); }
        else mk("_", "");
    }
    int mode = 0, last = -1;
    for (ExonPos& e : rec) {
        auto it = level2pos.find(e.graphLevel);
        if (it != level2pos.end()) {
            if (mode == 2) last = -1;
            e.positionInExon = it->second; TREQUIRE(last == -1 || (int)e.positionInExon == last + 1, "consecutive exon positions");
            last = (int)e.positionInExon; mode = 1; out.push_back(e);
        } else if (mode == 1) mode = 2;
    }
}

std::vector<ExonPos> remove_double(const std::vector<ExonPos>& v) {   // HLATyper.cpp:4020-4083
    std::map<int, std::vector<ExonPos>> per; for (const ExonPos& e : v) per[e.graphLevel].push_back(e);
    std::vector<ExonPos> r;
    for (auto& kv : per) {
        size_t best = 0; unsigned char bq = 0;
        for (size_t i = 0; i < kv.second.size(); i++) {
            const ExonPos& e = kv.second[i]; TREQUIRE(e.genotype == "_" || !e.qualities.empty(), "qualities present");
            unsigned char q = 0; if (e.genotype != "_") { for (size_t k = 0; k < e.qualities.size(); k++) if (k == 0 || (unsigned char)e.qualities[k] < q) q = (unsigned char)e.qualities[k]; }
            if (i == 0 || q > bq) { best = i; bq = q; }
        }
        r.push_back(kv.second[best]);
    }
    return r;
}

double chisq_p(double statistic) { double cdf = statistic <= 0 ? 0 : erf(sqrt(statistic / 2.0)); return 1 - cdf; }   // chi_squared(1) cdf as oracle/shim/boost/math/distributions/chi_squared.hpp evaluates it

struct LocusResult { std::string locus; int C = 0, R = 0; std::vector<double> LL; std::vector<int> mism; std::vector<double> pairLL, pairMavg, pairMmin; std::string call1, call2; double q1a = 0, q1b = 0; };

struct Typing {
    std::string dir; std::vector<std::string> loci = {"A", "B", "C", "DQA1", "DQB1", "DRB1", "DPA1", "DPB1", "DRA", "DRB3", "DRB4", "E", "F", "G", "H", "K", "V"};
    std::map<std::string, std::vector<std::string>> loci2exons; std::vector<std::string> files; std::map<std::string, int> level_of_name;
    std::map<std::string, std::pair<int, int>> gene_bounds; std::map<std::string, std::string> alleles_to_G; std::set<std::string> G_loci;
    std::vector<LocusResult> results;

    explicit Typing(const std::string& d) : dir(d) {
        for (const char* l : {"A", "B", "C", "E", "F", "G", "H", "J", "K", "L", "V"}) loci2exons[l] = {"exon_2", "exon_3"};
        for (const char* l : {"DQA1", "DQB1", "DRB1", "DPA1", "DPB1", "DRA", "DRB3", "DRB4"}) loci2exons[l] = {"exon_2"};
        DIR* dp = opendir((dir + "/PRG").c_str()); TREQUIRE(dp, "PRG directory readable");
        while (dirent* e = readdir(dp)) { std::string n = e->d_name; if (n != "." && n != "..") files.push_back(dir + "/PRG/" + n); } closedir(dp);
        // Graph::readGraphLoci (Graph.cpp:2563): level index = position in the concatenated header lines of the segment files
        std::ifstream seg(dir + "/PRG/segments.txt"); TREQUIRE(seg.is_open(), "segments.txt"); std::string line; std::vector<std::string> segfiles;
        while (seg.good()) { std::getline(seg, line); erase_nl(line); if (line.size()) segfiles.push_back(line); }
        int lv = 0;
        for (const std::string& f : segfiles) { std::ifstream s(dir + "/PRG/" + f); TREQUIRE(s.is_open(), "segment file"); std::string h; std::getline(s, h); erase_nl(h); std::vector<std::string> fs = split(h, " ");
            for (size_t i = 1; i < fs.size(); i++) { TREQUIRE(!level_of_name.count(fs[i]), "unique level names"); level_of_name[fs[i]] = lv++; } }
        for (const std::string& f : segfiles) {   // HLATyper.cpp:105-214 (boundaries only)
            std::vector<std::string> u = split(f, "_"); if (u.at(1) != "gene") continue;
            std::ifstream s(dir + "/PRG/" + f); std::string h; std::getline(s, h); erase_nl(h); std::vector<std::string> fs = split(h, " "); TREQUIRE(fs.at(0) == "IndividualID", "header");
            std::string locus = u.at(2); if (!gene_bounds.count(locus)) gene_bounds[locus] = {-1, -1};
            for (size_t i = 1; i < fs.size(); i++) { int l = level_of_name.at(fs[i]); auto& B = gene_bounds[locus]; if (B.first == -1 || l < B.first) B.first = l; if (B.second == -1 || l > B.second) B.second = l; }
        }
    }
    bool overlaps_genes(int a, int b) const { for (auto& kv : gene_bounds) if (kv.second.second >= a && kv.second.first <= b) return true; return false; }
    std::string exon_file(const std::string& locus, const std::string& exon) const {   // HLATyper.cpp:3127-3190: the last matching directory entry wins
        std::vector<std::string> ep = split(exon, "_"); int n = atoi(ep.at(1).c_str()); std::string r;
        for (const std::string& f : files) { std::vector<std::string> sl = split(f, "/"); std::vector<std::string> u = split(sl.back(), "_");
            if (u.size() >= 6 && u[1] == "gene" && (u[2] == "HLA-" + locus || u[2] == locus) && u[4] == "exon" && u[5] == itos(n) + ".txt") r = f; }
        return r;
    }
    void read_G(const std::string& gdir) {   // HLATyper.cpp:4150-4198
        if (!alleles_to_G.empty()) return;
        std::ifstream g(gdir + "/hla_nom_g.txt"); if (!g.is_open()) throw std::runtime_error("Can't open file hla_nom_g.txt - are you executing me from the right directory?");
        std::string line;
        while (g.good()) { std::getline(g, line); erase_nl(line); if (line.empty() || line[0] == '#') continue;
            std::vector<std::string> c = split(line, ";"); std::string ls = c.front(); TREQUIRE(ls.back() == '*', "locus with star"); G_loci.insert(ls.substr(0, ls.size() - 1));
            std::string code; if (c.back() != "") code = c.back(); else { TREQUIRE(c.size() == 3, "three components"); code = c.at(1); } code = ls + code;
            for (const std::string& a : split(c.at(1), "/")) alleles_to_G[ls + a] = code; }
    }
    std::string to_G(const std::vector<std::string>& alleles, bool& perfect) {   // HLATyper.cpp:4095-4148
        std::map<std::string, int> groups;
        for (const std::string& a : alleles) { std::vector<std::string> la = split(a, "*"); if (la.size() != 2) throw std::runtime_error("Weird allele: " + a); if (!alleles_to_G.count(a)) continue; groups[alleles_to_G.at(a)]++; }
        if (groups.empty()) { perfect = false; return join(alleles, ";"); }
        if (groups.size() == 1) { perfect = true; return groups.begin()->first; }
        perfect = false;
        // Utilities::get_map_keys_sorted_by_value (Utilities.cpp): keys sorted by value, descending, std::sort on (key) with value comparator
        std::vector<std::string> keys; for (auto& kv : groups) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end(), [&](const std::string& x, const std::string& y) { return groups.at(x) < groups.at(y); }); std::reverse(keys.begin(), keys.end());
        return keys.at(0);
    }

    void run(const std::vector<Aln>& reads /* 2 per pair */, double is_mean, double is_sd, const std::string& out_dir, const std::string& gdir);
};

void Typing::run(const std::vector<Aln>& all, double is_mean, double is_sd, const std::string& outDir, const std::string& gdir) {
    // gene filter (processBAM.cpp:2427-2446)
    std::vector<const Aln*> A1, A2;
    for (size_t p = 0; p + 1 < all.size(); p += 2) { const Aln& a = all[p]; const Aln& b = all[p + 1]; bool inc = false;
        if (a.first_level() != -1) inc = inc || overlaps_genes(a.first_level(), a.last_level());
        if (b.first_level() != -1) inc = inc || overlaps_genes(b.first_level(), b.last_level());
        if (inc) { A1.push_back(&a); A2.push_back(&b); } }
    size_t NP = A1.size(); TREQUIRE(NP > 0, "rawPairedReads.size() > 0");
    const double insertionP = 0.001, deletionP = 0.001;
    const double ll_ins = log(insertionP), ll_ins_actual = ll_ins + log(1.0 / 4.0), ll_del = log(deletionP), ll_mm = log(1 - insertionP - deletionP);
    const double minMapQ = 0.0, minPosMapQ = 0.7, minW = 0.0; const int F20N = 20; const bool F20 = true; const double F20MinProp = 0.1; const int F20Limit = 2;
    const int hc_minCov = 100; const double unacc_minFrac = 0.2; const int unacc_minCov = 30;

    const int K = 31; std::unordered_map<std::string, int> kmer_counts;
    for (size_t p = 0; p < NP; p++) for (const Aln* a : {A1[p], A2[p]}) { std::string raw = a->reverse ? rc_seq(a->seq_bam) : a->seq_bam; for (const std::string& k : kmers_of(raw, K)) kmer_counts[kmer_canon(k)]++; }

    // alignment statistics (HLATyper.cpp:1026-1125)
    int st_valid = 0, st_perfect = 0, st_onePerfect = 0, st_validDist = 0; std::vector<double> dists; double fsum = 0;
    for (size_t p = 0; p < NP; p++) {
        if (strands_valid(*A1[p], *A2[p])) { st_valid++; double d = pair_distance(*A1[p], *A2[p]); dists.push_back(d); if (fabs(d - is_mean) <= 5 * is_sd) st_validDist++; }
        double f1 = fraction_ok(*A1[p]), f2 = fraction_ok(*A2[p]); if (f1 == 1) st_perfect++; if (f2 == 1) st_perfect++; if (f1 == 1 || f2 == 1) st_onePerfect++; fsum += f1; fsum += f2;
    }
    std::sort(dists.begin(), dists.end()); double dsum = 0; for (double d : dists) dsum += d; double dmean = 0, dmed = 0; if (!dists.empty()) { dmean = dsum / (double)dists.size(); dmed = dists[dists.size() / 2]; }
    mkdir(outDir.c_str(), 0777);
    auto perc = [](double a, double b) { return dtos((a / b) * 100); };
    { std::ofstream s(outDir + "/summaryStatistics.txt");
      s << "\nRead alignment statistics:\n" << "\t - Total number (paired) alignments:                 " << NP << "\n"
        << "\t\t - Alignment pairs with strands OK:                  " << st_valid << " (" << perc(st_valid, NP) << "%)\n"
        << "\t\t - Alignment pairs with strands OK && distance OK:   " << st_validDist << " (" << perc(st_validDist, NP) << "%)\n"
        << "\t\t - Alignment pairs with strands OK, mean distance:   " << dmean << "\n" << "\t\t - Alignment pairs with strands OK, median distance: " << dmed << "\n"
        << "\t\t - Alignment pairs, average fraction alignment OK:   " << (fsum / (2.0 * (double)NP)) << "\n" << "\t\t - Alignment pairs, at least one alignment perfect:   " << st_onePerfect << "\n"
        << "\t\t - Single alignments, perfect (total):   " << st_perfect << " (" << NP * 2 << ")\n" << "\t - Total number (unpaired) alignments:                 " << 0 << "\n"
        << "\t\t - Alignment pairs, average fraction alignment OK:   " << 0.0 << "\n" << "\t\t - Single alignments, perfect (total):   " << 0 << " (" << 0 << ")\n" << "\t\t - Alignments with length >= " << 1000 << ":   " << 0 << "\n"; }
    std::ofstream best(outDir + "/R1_bestguess.txt"), bestG(outDir + "/R1_bestguess_G.txt"), hist(outDir + "/histogram_matchesPerRead.txt");
    std::string unaccField = "NColumns_UnaccountedAllele_fGT" + dtos(unacc_minFrac);
    best << "Locus\tChromosome\tAllele\tQ1\tQ2\tAverageCoverage\tCoverageFirstDecile\tMinimumCoverage\tproportionkMersCovered\tLocusAvgColumnError\t" << unaccField << "\n";
    bestG << "Locus\tChromosome\tAllele\tQ1\tQ2\tAverageCoverage\tCoverageFirstDecile\tMinimumCoverage\tproportionkMersCovered\tLocusAvgColumnError\t" << unaccField << "\tperfectG\n";
    hist << "Locus\tLevelValue\n";
    results.clear();

    for (const std::string& locus : loci) {
        LocusResult LR; LR.locus = locus; std::set<std::string> utilized; int bases_used = 0;
        std::vector<int> cLevels, cExon, cExonPos; int lmin = -1, lmax = -1; std::map<std::string, std::string> seqs; int totalColumns = 0; std::map<int, int> exon_lengths;
        const std::vector<std::string>& exons = loci2exons.at(locus);
        for (size_t ei = 0; ei < exons.size(); ei++) {
            std::string fn = exon_file(locus, exons[ei]); std::ifstream s(fn); TREQUIRE(!fn.empty() && s.is_open(), "exon file readable");
            std::vector<std::string> lines; while (s.good()) { std::string l; std::getline(s, l); erase_nl(l); lines.push_back(l); }
            std::vector<std::string> hf = split(lines.at(0), " "); TREQUIRE(hf.at(0) == "IndividualID", "header"); std::vector<std::string> names(hf.begin() + 1, hf.end());
            unsigned first = level_of_name.at(names.front()), lastl = level_of_name.at(names.back()); TREQUIRE(lastl > first, "last_graph_level > first_graph_level");
            unsigned explen = lastl - first + 1; TREQUIRE(names.size() == explen, "expected allele length"); totalColumns += (int)names.size();
            for (unsigned i = 0; i < explen; i++) { int gl = (int)(first + i); TREQUIRE(level_of_name.at(names[i]) == gl, "consecutive levels"); cLevels.push_back(gl); cExon.push_back((int)ei); cExonPos.push_back((int)i);
                if (lmin == -1 || lmin > gl) lmin = gl; if (lmax == -1 || lmax < gl) lmax = gl; }
            exon_lengths[(int)ei] = (int)explen;
            for (size_t li = 1; li < lines.size(); li++) if (lines[li].size()) { std::vector<std::string> f = split(lines[li], " "); TREQUIRE(f.size() == hf.size(), "field count"); if (f[0].find(":") == std::string::npos) continue;
                std::string sq; for (size_t k = 1; k < f.size(); k++) sq += f[k];
                if (ei == 0) { TREQUIRE(!seqs.count(f[0]), "type seen once"); seqs[f[0]] = sq; } else { TREQUIRE(seqs.count(f[0]), "type known"); seqs.at(f[0]) += sq; } }
            TREQUIRE(!seqs.empty(), "sequences present");
        }
        std::map<int, unsigned int> level2pos, level2exon, level2exonpos; for (size_t i = 0; i < cLevels.size(); i++) { level2pos[cLevels[i]] = (unsigned)i; level2exon[cLevels[i]] = cExon[i]; level2exonpos[cLevels[i]] = cExonPos[i]; }
        std::map<std::string, unsigned> type2cluster; std::vector<std::set<std::string>> clusters; std::map<std::string, unsigned> seq2cluster; std::vector<std::string> cluster_seq;
        for (auto& kv : seqs) { TREQUIRE(kv.second.size() == seqs.begin()->second.size(), "equal sequence lengths");
            if (seq2cluster.count(kv.second)) { unsigned c = seq2cluster.at(kv.second); clusters[c].insert(kv.first); type2cluster[kv.first] = c; }
            else { clusters.push_back({kv.first}); seq2cluster[kv.second] = (unsigned)clusters.size() - 1; type2cluster[kv.first] = (unsigned)clusters.size() - 1; } }
        for (auto& cl : clusters) cluster_seq.push_back(seqs.at(*cl.begin()));
        const size_t C = clusters.size();

        std::vector<std::vector<ExonPos>> reads;
        for (size_t p = 0; p < NP; p++) {
            const Aln& a = *A1[p]; const Aln& b = *A2[p]; std::vector<ExonPos> e1, e2;
            project(a, b, lmin, lmax, level2pos, e1); project(b, a, lmin, lmax, level2pos, e2);
            TREQUIRE(a.mapQ >= 0 && a.mapQ <= 1, "mapQ in [0,1]");
            if (strands_valid(a, b) && fabs(pair_distance(a, b) - is_mean) <= 5 * is_sd && a.mapQ >= minMapQ && weighted_ok(a) >= minW && weighted_ok(b) >= minW) {
                std::vector<ExonPos> t = e1; t.insert(t.end(), e2.begin(), e2.end());
                if (!t.empty()) { t = remove_double(t); reads.push_back(t); }
                hist << locus << "\t" << "read" << weighted_ok(a) << "\n"; hist << locus << "\t" << "read" << weighted_ok(b) << "\n"; hist << locus << "\t" << "readPair" << (weighted_ok(a) + weighted_ok(b)) / 2.0 << "\n";
            }
        }
        const size_t R = reads.size();
        std::set<std::string> ignore_ids; std::map<unsigned, std::map<std::string, int>> counts_post; std::map<unsigned, std::set<std::string>> ignore_alleles;
        if (F20) {
            std::map<unsigned, int> kicked, kicked_robust; std::map<unsigned, std::vector<std::string>> pa; std::map<unsigned, std::vector<double>> pw; std::map<unsigned, std::vector<unsigned>> pr;
            for (unsigned r = 0; r < R; r++) for (const ExonPos& e : reads[r]) { TREQUIRE(e.mapQ_position >= 0 && e.mapQ_position <= 1, "mapQ_position in [0,1]"); if (e.mapQ_position < minPosMapQ) continue;
                pa[e.positionInExon].push_back(e.genotype); pw[e.positionInExon].push_back((e.thisW + e.pairedW) / 2.0); pr[e.positionInExon].push_back(r); }
            for (auto& pos : pa) {
                int n = (int)pos.second.size(); if (n < F20N) continue;
                std::vector<unsigned> idx; for (int i = 0; i < n; i++) idx.push_back((unsigned)i);
                const std::vector<double>& W = pw.at(pos.first);
                std::sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return W.at(a) < W.at(b); }); std::reverse(idx.begin(), idx.end());
                std::map<std::string, int> first20; for (int i = 0; i < F20N; i++) first20[pos.second.at(idx.at(i))]++;
                std::set<std::string> kickedAlleles;
                for (int i = 0; i < n; i++) { const std::string& al = pos.second[i]; int cnt = first20.count(al) ? first20.at(al) : 0; double prop = (double)cnt / (double)F20;   // divides by the bool (HLATyper.cpp:1583)
                    if (prop < F20MinProp) { kickedAlleles.insert(al); ignore_alleles[pos.first].insert(al); kicked[pr.at(pos.first)[i]]++; } }
                std::map<std::string, int> howMany; for (int i = 0; i < n; i++) if (kickedAlleles.count(pos.second[i])) howMany[pos.second[i]]++;
                for (int i = 0; i < n; i++) if (howMany.count(pos.second[i]) && howMany.at(pos.second[i]) >= 2) kicked_robust[pr.at(pos.first)[i]]++;
            }
            for (auto& kr : kicked_robust) if (kr.second > F20Limit) { const ExonPos& e = reads.at(kr.first).at(0); ignore_ids.insert(e.thisRead_ID); ignore_ids.insert(e.pairedRead_ID); }
        }
        std::map<unsigned, std::map<std::string, double>> minStrandFreq, read1Freq;
        {
            std::map<unsigned, std::map<std::string, int>> cnt, cnt1; std::map<unsigned, std::map<std::string, std::pair<int, int>>> byStrand;
            for (unsigned r = 0; r < R; r++) for (const ExonPos& e : reads[r]) {
                if (ignore_ids.count(e.thisRead_ID)) continue; if (e.mapQ_position < minPosMapQ) continue;
                unsigned pos = e.positionInExon; if (ignore_alleles.count(pos) && ignore_alleles.at(pos).count(e.genotype)) continue;
                if (!cnt[pos].count(e.genotype)) { cnt[pos][e.genotype] = 0; byStrand[pos][e.genotype] = {0, 0}; cnt1[pos][e.genotype] = 0; }
                cnt.at(pos).at(e.genotype)++; if (e.reverse) byStrand.at(pos).at(e.genotype).second++; else byStrand.at(pos).at(e.genotype).first++; if (e.fromFirstRead) cnt1.at(pos).at(e.genotype)++;
            }
            for (auto& pos : cnt) {
                int tot = 0; for (auto& a : pos.second) tot += a.second;
                if (tot >= hc_minCov) for (auto& a : pos.second) counts_post[pos.first][a.first] = a.second;   // highCoverage_filter_alleles is off in short-read mode
                for (auto& a : byStrand.at(pos.first)) { int t = a.second.first + a.second.second; int mn = a.second.first < a.second.second ? a.second.first : a.second.second; int c1 = cnt1.at(pos.first).count(a.first) ? cnt1.at(pos.first).at(a.first) : 0;
                    minStrandFreq[pos.first][a.first] = (double)mn / (double)t; read1Freq[pos.first][a.first] = (double)c1 / (double)t; }
            }
        }
        std::map<int, std::map<int, std::vector<ExonPos>>> pile;
        for (unsigned r = 0; r < R; r++) for (const ExonPos& e : reads[r]) {
            if (e.mapQ_position < minPosMapQ) continue; if (ignore_alleles.count(e.positionInExon) && ignore_alleles.at(e.positionInExon).count(e.genotype)) continue; if (ignore_ids.count(e.thisRead_ID)) continue;
            pile[(int)level2exon.at(e.graphLevel)][(int)level2exonpos.at(e.graphLevel)].push_back(e); hist << locus << "\t" << "base" << e.thisW << "\n";
        }
        { std::ofstream ps(outDir + "/R1_pileup_" + locus + ".txt");
          for (auto& ex : pile) { int exon = ex.first, L = exon_lengths.at(exon);
            for (int ep = 0; ep < L; ep++) {
                if (ex.second.count(ep)) {
                    const std::vector<ExonPos>& pu = ex.second.at(ep); std::vector<std::string> fields = {itos(exon), itos(ep), itos((long long)pu.size())}; std::vector<std::string> gts; unsigned thisPos = 0; std::map<std::string, std::vector<int>> ac;
                    for (size_t i = 0; i < pu.size(); i++) { const ExonPos& e = pu[i]; std::vector<std::string> qs; for (char q : e.qualities) qs.push_back(itos((int)q));
                        gts.push_back(e.genotype + " (" + join(qs, ", ") + ")" + " [" + "pairsDistance " + dtos(e.strands_distance) + " | " + "alignmentLength " + itos(e.alnCols) + " | " + dtos(e.mapQ_position) + " | " + dtos(e.mapQ) + " " + dtos(e.mapQ) + " | " +
                                      dtos(e.thisW) + " " + dtos(e.pairedW) + " | " + e.thisRead_ID + " " + e.pairedRead_ID + "]");
                        utilized.insert(e.thisRead_ID); ac[e.genotype].push_back(e.alnCols); if (i == 0) thisPos = e.positionInExon; else TREQUIRE(thisPos == e.positionInExon, "same exon position"); }
                    fields.push_back(join(gts, ", ")); std::string summary;
                    for (auto& a : ac) { long long tot = 0; for (int l : a.second) tot += l; double avg = (double)tot / (double)a.second.size();
                        summary += a.first + "x" + itos((long long)a.second.size()) + "[" + dtos(avg) + ";" + dtos(minStrandFreq.at(thisPos).at(a.first)) + ";" + dtos(read1Freq.at(thisPos).at(a.first)) + "]"; }
                    ps << join(fields, "\t") << "\t" << summary << "\n";
                } else ps << join({itos(exon), itos(ep), itos(0)}, "\t") << "\n";
            } } }
        { std::ofstream rs(outDir + "/R1_readIDs_" + locus + ".txt"); for (const std::string& id : utilized) rs << id << "\n"; }

        // per-read x cluster log-likelihoods (HLATyper.cpp:2049-2277)
        LR.C = (int)C; LR.R = (int)R; LR.LL.assign(C * R, 0); LR.mism.assign(C * R, 0);
        for (size_t c = 0; c < C; c++) { const std::string& cs = cluster_seq[c];
            for (size_t r = 0; r < R; r++) { double ll = 0; int mm = 0;
                for (const ExonPos& e : reads[r]) {
                    if (e.mapQ_position < minPosMapQ) continue; if (ignore_alleles.count(e.positionInExon) && ignore_alleles.at(e.positionInExon).count(e.genotype)) continue; if (ignore_ids.count(e.thisRead_ID)) continue;
                    if (c == 0) bases_used++;
                    double lp = 0; std::string ex = cs.substr(e.positionInExon, 1); const std::string& rg = e.genotype; unsigned l_diff = (unsigned)(rg.size() - ex.size());
                    if (ex == "_") { if (rg != "_") { TREQUIRE(rg.find("_") == std::string::npos, "no gap inside insertion"); lp += (ll_ins_actual * (1 + l_diff)); } }
                    else {
                        if (rg.substr(0, 1) == "_") lp += ll_del;
                        else { lp += ll_mm; TREQUIRE(!e.qualities.empty(), "quality present"); double pc = phred_to_pcorrect((unsigned char)e.qualities.at(0)); if (pc > 0.999) pc = 0.999; TREQUIRE(pc >= 0 && pc <= 1, "pCorrect"); if (pc == 0) pc = 0.001;
                            if (ex == rg.substr(0, 1)) lp += log(pc); else { double pi = (1 - pc) * (1.0 / 3.0); lp += log(pi); } }
                        lp += (ll_ins_actual * l_diff);
                    }
                    if (rg != "_" && rg != ex) mm++;
                    ll += lp;
                }
                LR.LL[c * R + r] = ll; LR.mism[c * R + r] = mm; } }
        // allele pairs, single-threaded order (HLATyper.cpp:2280-2364)
        std::vector<double> LLs, Mavg, Mmin; std::vector<std::pair<unsigned, unsigned>> ids;
        for (unsigned c1 = 0; c1 < C; c1++) for (unsigned c2 = c1; c2 < C; c2++) { double sa = 0, sm = 0, pl = 0;
            for (size_t r = 0; r < R; r++) { double v = log_avg(LR.LL[c1 * R + r], LR.LL[c2 * R + r]); TREQUIRE(exp(v) >= 0 && exp(v) <= 1, "exp(logAvg) in [0,1]"); int m1 = LR.mism[c1 * R + r], m2 = LR.mism[c2 * R + r];
                sa += ((double)(m1 + m2) / 2.0); sm += (m1 < m2 ? m1 : m2); pl += v; }
            LLs.push_back(pl); Mavg.push_back(sa); Mmin.push_back(sm); ids.push_back({c1, c2}); }
        LR.pairLL = LLs; LR.pairMavg = Mavg; LR.pairMmin = Mmin;
        std::vector<size_t> order; for (size_t i = 0; i < LLs.size(); i++) order.push_back(i);
        std::sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { if (LLs.at(a) == LLs.at(b)) return Mavg.at(b) < Mavg.at(a); return LLs.at(a) < LLs.at(b); }); std::reverse(order.begin(), order.end());
        size_t maxI = 0; for (size_t i = 1; i < LLs.size(); i++) if (LLs[i] > LLs[maxI]) maxI = i; double LLmax = LLs[maxI];
        std::vector<double> Pn; double Psum = 0; for (double v : LLs) Psum += exp(v - LLmax);
        if (Psum > 0) for (double v : LLs) { double pn = exp(v - LLmax) / Psum; TREQUIRE(pn >= 0 && pn <= 1, "P_normalized in [0,1]"); Pn.push_back(pn); } else for (size_t i = 0; i < LLs.size(); i++) Pn.push_back(1.0 / (double)LLs.size());
        std::map<int, double> marg;
        auto members = [&](unsigned c, const char* d) { return join(std::vector<std::string>(clusters[c].begin(), clusters[c].end()), d); };
        { std::ofstream ap(outDir + "/R1_PP_" + locus + "_pairs.txt"); ap << "ClusterID\tP\tLL\tMismatches_avg\n";
          for (size_t k = 0; k < order.size(); k++) { size_t i = order[k]; ap << members(ids[i].first, ";") + "/" + members(ids[i].second, ";") << "\t" << Pn.at(i) << "\t" << LLs.at(i) << "\t" << Mavg.at(i) << "\n";
            marg[(int)ids[i].first] += Pn.at(i); if (ids[i].second != ids[i].first) marg[(int)ids[i].second] += Pn.at(i); } }
        auto map_max = [](const std::map<int, double>& m) { double mx = 0; int im = 0; bool f = true; for (auto& kv : m) if (f || kv.second > mx) { mx = kv.second; im = kv.first; f = false; } return std::make_pair(mx, im); };
        std::pair<double, int> b1 = map_max(marg); std::string b1ID = members((unsigned)b1.second, ";"), b1one = *clusters[b1.second].begin();
        std::map<int, double> alt, altmm;
        for (size_t i = 0; i < ids.size(); i++) { if ((int)ids[i].first == b1.second) { alt[(int)ids[i].second] = Pn.at(i); altmm[(int)ids[i].second] = Mmin.at(i); } else if ((int)ids[i].second == b1.second) { alt[(int)ids[i].first] = Pn.at(i); altmm[(int)ids[i].first] = Mmin.at(i); } }
        std::pair<double, int> one2 = map_max(alt); std::map<int, double> mmBest; for (auto& kv : alt) if (kv.second == one2.first) mmBest[kv.first] = -1 * altmm.at(kv.first);
        std::pair<double, int> b2 = map_max(mmBest); std::string b2one = *clusters[b2.second].begin(), b2ID = members((unsigned)b2.second, ";");
        LR.call1 = b1ID; LR.call2 = b2ID; LR.q1a = b1.first; LR.q1b = one2.first;

        double locus_cov = (double)bases_used / (double)totalColumns;
        std::vector<double> poscov; for (size_t i = 0; i < cLevels.size(); i++) poscov.push_back((double)pile[(int)level2exon.at(cLevels[i])][(int)level2exonpos.at(cLevels[i])].size());
        std::sort(poscov.begin(), poscov.end(), std::less<int>());
        size_t allTot = 0, allInc = 0; std::vector<int> colTot, colInc; double avgErr; int nUnacc = 0; double kA1, kA2;
        {
            const std::string& s1 = seqs.at(b1one); const std::string& s2 = seqs.at(b2one); std::vector<std::string> by1, by2;
            for (size_t i = 0; i < s1.size(); i++) {
                if (i == 0 || cExon[i] != cExon[i - 1]) { by1.push_back(""); by2.push_back(""); }
                std::string u1 = s1.substr(i, 1), u2 = s2.substr(i, 1); by1.back().push_back(u1[0]); by2.back().push_back(u2[0]);
                int tot = 0, inc = 0; for (const ExonPos& e : pile.at(cExon[i]).at(cExonPos[i])) { tot++; if (e.genotype != u1 && e.genotype != u2) inc++; }
                allTot += tot; allInc += inc; colTot.push_back(tot); colInc.push_back(inc);
                if (counts_post.count((unsigned)i)) { int cov = 0; for (auto& a : counts_post.at((unsigned)i)) cov += a.second;
                    if (cov >= unacc_minCov) for (auto& a : counts_post.at((unsigned)i)) { if (a.first == u1 || a.first == u2) continue; if ((double)a.second / (double)cov >= unacc_minFrac) nUnacc++; } }
            }
            auto kpres = [&](const std::vector<std::string>& ex) -> double { int tot = 0, pres = 0; for (const std::string& e : ex) { std::string ng; for (char ch : e) if (ch != '_') ng.push_back(ch);
                for (const std::string& k : kmers_of(ng, K)) { tot++; if (k.find("*") == std::string::npos) { std::string key = kmer_canon(k); auto it = kmer_counts.find(key); if (it != kmer_counts.end() && it->second > 0) pres++; } } }
                if (tot == 0) return -1; return (double)pres / (double)tot; };
            kA1 = kpres(by1); kA2 = kpres(by2);
            std::ofstream ce(outDir + "/R1_columnIncompatibilities_" + locus + ".txt"); ce << "Column\tCoverage\tExpectedIncompatible\tObservedIncompatible\tp\n";
            avgErr = allTot > 0 ? (double)allInc / (double)allTot : 0;
            for (size_t col = 0; col < s1.size(); col++) { int cov = colTot[col]; double expd = avgErr * cov; int obs = colInc[col]; double p = 1;
                if (obs > expd) { double o0 = cov - obs, o1 = obs, e0 = cov - expd, e1 = expd; TREQUIRE(e0 > 0 && e1 > 0, "expected > 0"); double st = 0; st += pow(o0 - e0, 2) / e0; st += pow(o1 - e1, 2) / e1; p = chisq_p(st); TREQUIRE(p >= 0 && p <= 1, "p in [0,1]"); }
                ce << join({itos((long long)col), itos(cov), dtos(expd), itos(obs), dtos(p)}, "\t") << "\n"; }
        }
        int dec = (int)((double)poscov.size() / 10.0); double firstDec = poscov.at(dec), minCov = poscov.at(0);
        best << locus << "\t" << 1 << "\t" << b1ID << "\t" << b1.first << "\t" << b2.first << "\t" << locus_cov << "\t" << firstDec << "\t" << minCov << "\t" << kA1 << "\t" << avgErr << "\t" << nUnacc << "\n";
        best << locus << "\t" << 2 << "\t" << b2ID << "\t" << one2.first << "\t" << b2.first << "\t" << locus_cov << "\t" << firstDec << "\t" << minCov << "\t" << kA2 << "\t" << avgErr << "\t" << nUnacc << "\n" << std::flush;
        read_G(gdir);
        if (G_loci.count(locus)) { bool p1, p2; std::string g1 = to_G(std::vector<std::string>(clusters[b1.second].begin(), clusters[b1.second].end()), p1), g2 = to_G(std::vector<std::string>(clusters[b2.second].begin(), clusters[b2.second].end()), p2);
            bestG << locus << "\t" << 1 << "\t" << g1 << "\t" << b1.first << "\t" << b2.first << "\t" << locus_cov << "\t" << firstDec << "\t" << minCov << "\t" << kA1 << "\t" << avgErr << "\t" << nUnacc << "\t" << p1 << "\n";
            bestG << locus << "\t" << 2 << "\t" << g2 << "\t" << one2.first << "\t" << b2.first << "\t" << locus_cov << "\t" << firstDec << "\t" << minCov << "\t" << kA2 << "\t" << avgErr << "\t" << nUnacc << "\t" << p2 << "\n" << std::flush; }
        results.push_back(std::move(LR));
    }
    { std::ofstream ps(outDir + "/R1_parameters.txt"); ps << "Loci = " << join(loci, ",") << "\n" << "veryConservativeReadLikelihoods = " << true << "\n"; }
}

std::string t_err;

} // namespace

extern "C" {

const char* hlala_oracle_type_last_error() { return t_err.c_str(); }

// reads 2p / 2p+1 are the mates of pair p, named "r<p>" like oracle/ref_driver.cpp does. bases/quals are the primary records' SEQ/QUAL (alignment orientation).
void* hlala_oracle_type(const char* prg_dir, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int cap, const int32_t* n_cols, const int32_t* level,
                        const uint8_t* gchar, const uint8_t* schar, const uint8_t* mapq, const uint8_t* read_reverse, const double* read_mapq, double is_mean, double is_sd, const char* out_dir, const char* g_dir) {
    try {
        Typing* T = new Typing(prg_dir);
        std::vector<Aln> all((size_t)n_reads);
        for (long long r = 0; r < n_reads; r++) { Aln& a = all[(size_t)r]; int n = n_cols[r]; if (n > cap) throw std::runtime_error("alignment longer than the column capacity"); size_t o = (size_t)r * cap;
            a.level.assign(level + o, level + o + n); a.g.assign((const char*)gchar + o, n); a.s.assign((const char*)schar + o, n); a.mapq.assign((const char*)mapq + o, n);
            a.seq_bam.assign((const char*)bases + read_off[r], (size_t)(read_off[r + 1] - read_off[r])); a.qual.assign((const char*)quals + read_off[r], a.seq_bam.size());
            a.name = "r" + std::to_string(r / 2); a.reverse = read_reverse[r] != 0; a.from_first = false;   /* processBAM.cpp:3545-3548: the flag set on read 1 is overwritten by the re-assignment of chains.first; both mates end up false */ a.mapQ = read_mapq[r]; }
        T->run(all, is_mean, is_sd, out_dir, g_dir);
        return T;
    } catch (const std::exception& e) { t_err = e.what(); return nullptr; }
}
int hlala_oracle_type_n_loci(void* h) { return (int)((Typing*)h)->results.size(); }
int hlala_oracle_type_dims(void* h, int locus, int* C, int* R) { const LocusResult& L = ((Typing*)h)->results.at(locus); *C = L.C; *R = L.R; return 0; }
int hlala_oracle_type_read_ll(void* h, int locus, double* LL, int32_t* mism) { const LocusResult& L = ((Typing*)h)->results.at(locus); memcpy(LL, L.LL.data(), L.LL.size() * 8); for (size_t i = 0; i < L.mism.size(); i++) mism[i] = L.mism[i]; return 0; }
int hlala_oracle_type_pair_ll(void* h, int locus, double* ll, double* mavg, double* mmin) { const LocusResult& L = ((Typing*)h)->results.at(locus); memcpy(ll, L.pairLL.data(), L.pairLL.size() * 8); memcpy(mavg, L.pairMavg.data(), L.pairMavg.size() * 8); memcpy(mmin, L.pairMmin.data(), L.pairMmin.size() * 8); return 0; }
void hlala_oracle_type_free(void* h) { delete (Typing*)h; }

} // extern "C"
