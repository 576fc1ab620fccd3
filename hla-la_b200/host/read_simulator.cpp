// See read_simulator.h. Written against the behaviour of simulator/readSimulator.cpp (constructor :58-334, simulate_paired_reads_from_edgePath :1194-1693).
#include "read_simulator.h"
#include "prg_graph.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>

namespace hlala {

namespace {

typedef std::map<char, double> QMap;

void chomp(std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); }
std::vector<std::string> split_tabs(const std::string& s) {   // Utilities::split(line, "\t"): empty fields kept
    std::vector<std::string> out; size_t a = 0;
    for (;;) { size_t p = s.find('\t', a); out.push_back(s.substr(a, p == std::string::npos ? std::string::npos : p - a)); if (p == std::string::npos) break; a = p + 1; }
    return out;
}
int to_int(const std::string& s) { std::stringstream ss(s); int v = 0; ss >> v; return v; }            // Utilities::StrtoI
double to_double(const std::string& s) { std::stringstream ss(s); double v = 0; ss >> v; return v; }   // Utilities::StrtoD

QMap normalized(const QMap& m) {   // Utilities::normalize_map (Utilities.cpp:1012-1027)
    double sum = 0; for (const auto& kv : m) { if (kv.second < 0) throw std::runtime_error("quality matrix: negative count"); sum += kv.second; }
    if (!(sum > 0)) throw std::runtime_error("quality matrix: a read position without observations");
    QMap r; for (const auto& kv : m) r[kv.first] = kv.second / sum; return r;
}

// keeps the lowest (size - drop) quality classes of one read position and renormalises their frequencies (:233-262, :269-298)
void drop_upper_classes(QMap& freq, QMap& correct, int drop) {
    std::vector<char> have; for (const auto& kv : freq) have.push_back(kv.first);
    std::sort(have.begin(), have.end());
    if (!(drop < (int)have.size() - 1)) throw std::runtime_error("quality matrix: cannot remove " + std::to_string(drop) + " quality classes of " + std::to_string(have.size()));
    QMap f2, c2;
    for (int i = 0; i < (int)have.size() - drop; i++) { f2[have[(size_t)i]] = freq.at(have[(size_t)i]); c2[have[(size_t)i]] = correct.at(have[(size_t)i]); }
    freq = normalized(f2); correct = c2;
}

// the three draws of Utilities.cpp:1133-1155, 1230-1244, 1248-1255 on the caller's generator
double uniform01(std::mt19937& g) { return std::generate_canonical<double, 53>(g); }
char choose(const QMap& m, std::mt19937& g) {
    const double f = uniform01(g); double acc = 0; char r = 0;
    for (const auto& kv : m) { acc += kv.second; if (f <= acc) { r = kv.first; break; } }
    return r;   // 0 when rounding leaves f above the last partial sum (the reference returns an unset char there and asserts on it one line later)
}
bool bernoulli(double p, std::mt19937& g) { return uniform01(g) <= p; }
char random_nucleotide(std::mt19937& g) { std::uniform_int_distribution<int> d(0, 3); return "ACGT"[d(g)]; }

char complement(char c) {   // Utilities::reverse_char_nucleotide (Utilities.cpp:1171-1204)
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'n': return 'n'; case '_': return '_'; case '*': return '*';
    }
    throw std::runtime_error(std::string("reverse complement of a character that is no nucleotide: ") + c);
}
std::string reverse_complement(const std::string& s) { std::string r(s.size(), 0); for (size_t k = 0; k < s.size(); k++) r[k] = complement(s[s.size() - 1 - k]); return r; }

std::string join_ints(const std::vector<int32_t>& v) { std::string s; for (size_t i = 0; i < v.size(); i++) { if (i) s += ' '; s += std::to_string(v[i]); } return s; }

} // namespace

ReadSimulator::ReadSimulator(const std::string& file, unsigned read_length, bool interpolate, int drop1, int drop2_extra) : len_(read_length) {
    if (drop1 < 0 || drop1 > 20 || drop2_extra < 0 || drop2_extra > 20) throw std::runtime_error("ReadSimulator: quality classes to remove must lie in 0..20");   // :62-65
    std::ifstream in(file.c_str());
    if (!in.is_open()) throw std::runtime_error("ReadSimulator: cannot open the quality matrix " + file);   // :319
    struct PerLength { std::vector<QMap> freq, correct; std::vector<double> indel, non_indel; };
    std::map<int, PerLength> by_length;
    std::string line; size_t line_no = 0;
    while (in.good()) {
        std::getline(in, line); line_no++; chomp(line);
        if (line.empty()) continue;
        const std::vector<std::string> f = split_tabs(line);
        if (line_no == 1) {   // :100-112
            static const char* names[6] = {"readLength", "qualityScore", "positionInRead", "N", "ExpectedCorrect", "EmpiricalCorrect"};
            if (f.size() != 6) throw std::runtime_error("ReadSimulator: " + file + ": 6 columns expected, found " + std::to_string(f.size()));
            for (int i = 0; i < 6; i++) if (f[(size_t)i] != names[i]) throw std::runtime_error("ReadSimulator: " + file + ": column " + std::to_string(i + 1) + " is not " + names[i]);
            continue;
        }
        if (f.size() < 6 || f[1].empty()) throw std::runtime_error("ReadSimulator: " + file + ": line " + std::to_string(line_no) + " is incomplete");
        const int L = to_int(f[0]);
        if (L <= 0) throw std::runtime_error("ReadSimulator: " + file + ": line " + std::to_string(line_no) + ": read length " + f[0]);
        PerLength& pl = by_length[L];
        if (pl.freq.empty()) { pl.freq.resize((size_t)L); pl.correct.resize((size_t)L); pl.indel.assign((size_t)L, 0.0); pl.non_indel.assign((size_t)L, 0.0); }
        const char q = f[1][0]; const unsigned pos = (unsigned)to_int(f[2]); const int n = to_int(f[3]);
        if (!(pos < read_length) || !(pos < (unsigned)L))   // assert(positionInRead < readLength), :127; vector::at of the per-length tables
            throw std::runtime_error("ReadSimulator: " + file + ": line " + std::to_string(line_no) + ": position " + f[2] + " beyond the read length");
        if (q == 0) { pl.indel[pos] += n; continue; }   // rows of quality character 0 count insertions / deletions (:129-132)
        pl.non_indel[pos] += n; pl.freq[pos][q] += n;
        if (pl.correct[pos].count(q)) throw std::runtime_error("ReadSimulator: " + file + ": quality defined twice, line " + std::to_string(line_no));   // :143-146
        pl.correct[pos][q] = to_double(f[5]);
    }
    std::vector<QMap> freq, correct; std::vector<double> indel, non_indel;
    if (by_length.count((int)read_length)) {
        const PerLength& pl = by_length.at((int)read_length); freq = pl.freq; correct = pl.correct; indel = pl.indel; non_indel = pl.non_indel;
    } else {   // the table of the nearest read length, positions scaled (:165-203)
        if (!interpolate) throw std::runtime_error("ReadSimulator: " + file + " holds no reads of length " + std::to_string(read_length) + " and interpolation is off");
        if (by_length.empty()) throw std::runtime_error("ReadSimulator: " + file + " holds no data");
        int which = -1, dist = -1;
        for (const auto& kv : by_length) { const int d = std::abs(kv.first - (int)read_length); if (which < 0 || d < dist) { which = kv.first; dist = d; } }
        const PerLength& pl = by_length.at(which);
        freq.resize(read_length); correct.resize(read_length); indel.resize(read_length); non_indel.resize(read_length);
        for (unsigned p = 0; p < read_length; p++) {
            const double target = ((double)p / (double)read_length) * (which - 1);
            const int t = (int)std::round(target);
            if (t < 0 || t >= (int)pl.correct.size()) throw std::runtime_error("ReadSimulator: interpolation outside the table");
            freq[p] = pl.freq[(size_t)t]; correct[p] = pl.correct[(size_t)t]; indel[p] = pl.indel[(size_t)t]; non_indel[p] = pl.non_indel[(size_t)t];
        }
    }
    indel_.resize(read_length);
    for (unsigned i = 0; i < read_length; i++) {   // :205-220
        if (correct[i].empty() || freq[i].empty()) throw std::runtime_error("ReadSimulator: " + file + ": no qualities for read position " + std::to_string(i));
        freq[i] = normalized(freq[i]);
        const double total = indel[i] + non_indel[i];
        if (total == 0) throw std::runtime_error("ReadSimulator: " + file + ": no observations for read position " + std::to_string(i));
        indel_[i] = indel[i] / total;
        if (indel_[i] == 0) indel_[i] = 1e-4;
    }
    freq_[0] = freq; correct_[0] = correct; freq_[1] = freq; correct_[1] = correct;
    if (drop1) for (unsigned p = 0; p < read_length; p++) drop_upper_classes(freq_[0][p], correct_[0][p], drop1);
    if (drop1 || drop2_extra) for (unsigned p = 0; p < read_length; p++) drop_upper_classes(freq_[1][p], correct_[1][p], drop1 + drop2_extra);
}

std::pair<double, double> ReadSimulator::average_error_rates() const {   // :28-53
    double r[2];
    for (int w = 0; w < 2; w++) {
        double sum = 0;
        for (unsigned p = 0; p < len_; p++) { double e = 0; for (const auto& kv : freq_[w][p]) e += kv.second * (1 - correct_[w][p].at(kv.first)); sum += e; }
        r[w] = sum / (double)len_;
    }
    return std::make_pair(r[0], r[1]);
}

std::vector<SimulatedPair> ReadSimulator::simulate_pairs_from_path(const std::string& path, double coverage, double diff_mean, double diff_sd, bool perfectly,
                                                                   const std::string& prefix, bool include_deletions) const {
    std::vector<SimulatedPair> out;
    std::string genome; std::vector<int32_t> origin;   // the path without its gaps, and the level every character came from (:1198-1220)
    for (size_t l = 0; l < path.size(); l++) if (path[l] != '_') { genome.push_back(path[l]); origin.push_back((int32_t)l); }
    const double start_rate = coverage / (2.0 * (double)len_);   // reads and their mates (:1222)
    const long long last = (long long)genome.size() - (long long)len_;
    if (last < 0) throw std::runtime_error("ReadSimulator: the path is shorter than one read");   // :1226-1230

    std::mt19937 gen;   // default seed
    std::poisson_distribution<int> starts(start_rate);
    std::vector<std::poisson_distribution<int>> ins, del;
    std::normal_distribution<double> jump_dist(diff_mean, diff_sd);
    for (unsigned i = 0; i < len_; i++) { ins.emplace_back(indel_[i]); del.emplace_back(indel_[i]); }
    size_t n_pairs = 0;

    auto one_base = [&](unsigned pos, char under, int which, char& base, char& qual) {   // :1267-1311
        qual = choose(freq_[which][pos], gen);
        if (!(qual > 0)) throw std::runtime_error("ReadSimulator: no quality drawn");
        const bool err = bernoulli(1 - correct_[which][pos].at(qual), gen);
        base = (err && !perfectly) ? random_nucleotide(gen) : under;
        qual = (char)(qual + 32);
    };
    struct Raw { std::string seq, qual, labels; std::vector<int32_t> coord, fa_coord; std::string fa_labels, fa_seq; bool ok; };
    auto sample = [&](long long at, int which, Raw& r) {   // :1313-1500
        r.seq.assign(len_, 0); r.qual.assign(len_, 0); r.labels.assign(len_, 0);
        r.coord.clear(); r.fa_coord.clear(); r.fa_labels.clear(); r.fa_seq.clear(); r.ok = true;
        int pending = 0;
        for (unsigned b = 0; b < len_; b++) {
            int n_ins = ins[b](gen), n_del = del[b](gen);
            if (perfectly) n_ins = n_del = 0;
            if (n_ins > 0) {
                for (int k = 0; k < n_ins; k++) {
                    const char under = random_nucleotide(gen); char base, qual; one_base(b, under, which, base, qual);
                    r.seq[b] = base; r.qual[b] = qual; r.labels[b] = '_'; b++;
                    r.coord.push_back(-1); r.fa_coord.push_back(-1); r.fa_labels.push_back('_'); r.fa_seq.push_back(base);
                    if (b >= len_) break;
                }
                if (b >= len_) break;
            }
            if (include_deletions) pending += n_del;
            else for (int k = 0; k < n_del; k++) { r.fa_coord.push_back((int32_t)at); r.fa_labels.push_back(genome.at((size_t)at)); r.fa_seq.push_back('_'); at++; }
            if ((unsigned long long)at >= genome.size()) { r.ok = false; break; }   // a start before the path counts as past its end (signed / unsigned compare, :1383)
            if (include_deletions && pending) {
                r.seq[b] = '_'; r.qual[b] = 1; r.labels[b] = genome[(size_t)at];
                r.coord.push_back((int32_t)at); r.fa_coord.push_back((int32_t)at); r.fa_labels.push_back(genome[(size_t)at]); r.fa_seq.push_back('_');
                at++; pending--;
            } else {
                char base, qual; one_base(b, genome[(size_t)at], which, base, qual);
                r.seq[b] = base; r.qual[b] = qual; r.labels[b] = genome[(size_t)at];
                r.coord.push_back((int32_t)at); r.fa_coord.push_back((int32_t)at); r.fa_labels.push_back(genome[(size_t)at]); r.fa_seq.push_back(base);
                at++;
            }
        }
        if (r.ok && r.coord.size() != r.seq.size()) throw std::runtime_error("ReadSimulator: a read without a coordinate per base");   // :1437
        // leading gap columns of the true alignment are dropped (:1446-1486; a trailing one cannot arise: every read ends in a base)
        size_t lead = 0; while (lead < r.fa_seq.size() && r.fa_seq[lead] == '_') lead++;
        if (!r.fa_seq.empty() && r.fa_seq.back() == '_') throw std::runtime_error("ReadSimulator: true alignment ending in a gap");   // the reference's scan runs off the string there (:1459)
        if (lead) {
            if (!(lead < r.fa_coord.size())) throw std::runtime_error("ReadSimulator: true alignment of gaps only");
            r.fa_coord.erase(r.fa_coord.begin(), r.fa_coord.begin() + (long)lead); r.fa_seq.erase(0, lead); r.fa_labels.erase(0, lead);
        }
    };
    auto to_levels = [&](std::vector<int32_t>& v) { for (int32_t& c : v) if (c != -1) c = origin.at((size_t)c); };

    Raw a, b;
    for (long long i = 0; i < last; i++) {   // :1515-1668
        const int n_start = starts(gen);
        for (int s = 0; s < n_start; s++) {
            const int jump = (int)std::floor(jump_dist(gen));
            const bool invert = bernoulli(0.5, gen);
            sample(i, invert ? 1 : 0, a);
            sample(i + (long long)len_ + jump, invert ? 0 : 1, b);
            if (!(a.ok && b.ok)) continue;
            n_pairs++;
            SimulatedPair p; p.jump = jump; p.first_read_minus_strand = invert;
            const std::string name = prefix + "r" + std::to_string(n_pairs) + "|||" + std::to_string(i) + "|||" + std::to_string(jump);
            SimulatedRead r1, r2;
            r1.name = name; r1.sequence = a.seq; r1.quality = a.qual; r1.levels = a.coord; to_levels(r1.levels); r1.edge_labels = a.labels;
            r1.full_levels = a.fa_coord; to_levels(r1.full_levels); r1.full_labels = a.fa_labels; r1.full_sequence = a.fa_seq;
            r2.name = name; r2.sequence = reverse_complement(b.seq); r2.quality = std::string(b.qual.rbegin(), b.qual.rend());
            r2.levels = b.coord; to_levels(r2.levels); std::reverse(r2.levels.begin(), r2.levels.end());
            r2.edge_labels = std::string(b.labels.rbegin(), b.labels.rend());   // reversed only: the truth stays on the forward strand (:1618-1630)
            r2.full_levels = b.fa_coord; to_levels(r2.full_levels); std::reverse(r2.full_levels.begin(), r2.full_levels.end());
            r2.full_labels = std::string(b.fa_labels.rbegin(), b.fa_labels.rend()); r2.full_sequence = std::string(b.fa_seq.rbegin(), b.fa_seq.rend());
            if (r1.full_levels.size() < r1.sequence.size() || r2.full_levels.size() < r2.sequence.size()) throw std::runtime_error("ReadSimulator: true alignment shorter than its read");   // :1652,1655
            if (invert) { p.first = r2; p.second = r1; } else { p.first = r1; p.second = r2; }   // oneReadPair::invert swaps the mates, nothing else (readSimulator.h:89-97)
            out.push_back(std::move(p));
        }
    }
    return out;
}

namespace {

std::vector<std::string> split_on(const std::string& s, char c) {   // Utilities::split(input, delimiter): an empty input has no fields, otherwise every delimiter separates
    std::vector<std::string> out; if (s.empty()) return out;
    size_t a = 0; for (;;) { size_t p = s.find(c, a); out.push_back(s.substr(a, p == std::string::npos ? std::string::npos : p - a)); if (p == std::string::npos) break; a = p + 1; }
    return out;
}
std::string join_strings(const std::vector<std::string>& v, const std::string& d) { std::string s; for (size_t i = 0; i < v.size(); i++) { if (i) s += d; s += v[i]; } return s; }
std::vector<std::string> file_lines(const std::string& path) {   // while(good) getline: a final newline leaves one empty line at the end, as in the reference
    std::ifstream in(path.c_str()); if (!in.is_open()) throw std::runtime_error("cannot open " + path);
    std::vector<std::string> out; std::string line; while (in.good()) { std::getline(in, line); chomp(line); out.push_back(line); } return out;
}
const std::string& uniform_choice(const std::vector<std::string>& v) { if (v.empty()) throw std::runtime_error("nothing to choose from"); return v.at((size_t)(rand() % (int)v.size())); }

struct GeneTables {
    std::vector<std::string> introns, exons, all;                                                      // segment IDs in segments.txt order
    std::map<std::string, std::map<std::string, std::vector<std::string>>> seq;                        // segment -> type -> one string per level
    std::map<std::string, std::vector<std::string>> level_names;                                       // segment -> level IDs
};

std::set<std::string> complete_genomic_types(const GeneTables& g) {   // HLATyper.cpp:353-404: the named (*) types of the first segment that holds any, introns first
    std::set<std::string> r; bool first = true;
    for (int ex = 0; ex <= 1; ex++) for (const std::string& id : (ex ? g.exons : g.introns)) {
        if (first) for (const auto& kv : g.seq.at(id)) if (kv.first.find('*') != std::string::npos) r.insert(kv.first);
        if (!g.seq.at(id).empty()) first = false;
    }
    for (const std::string& t : r) for (const auto& sg : g.seq) if (!sg.second.count(t)) throw std::runtime_error("type " + t + " is missing from segment " + sg.first);   // assert :398
    return r;
}
std::set<std::string> complete_exonic_types(const GeneTables& g) {    // HLATyper.cpp:269-335: the named types present in every exon
    if (g.exons.empty()) return complete_genomic_types(g);
    std::map<std::string, int> n; for (const std::string& id : g.exons) for (const auto& kv : g.seq.at(id)) n[kv.first]++;
    std::set<std::string> r; for (const auto& kv : n) if (kv.second == (int)g.exons.size() && kv.first.find('*') != std::string::npos) r.insert(kv.first);
    return r;
}

} // namespace

SimulatedIndividual simulate_one_individual(const std::string& prg_dir, const std::string& matrix, const std::string& out_dir, double is_mean, double is_sd, bool novel, bool with_error,
                                            unsigned seed) {
    if (matrix.empty()) throw std::runtime_error("simulate_one_individual: no quality matrix (assert(rS != 0), HLATyper.cpp:692)");
    const unsigned read_length = 101; const double coverage = 15;   // HLATyper.cpp:98-99
    ReadSimulator sim(matrix, read_length);
    // tables of the constructor (:103-215)
    std::map<std::string, GeneTables> genes;
    for (const std::string& line : file_lines(prg_dir + "/PRG/segments.txt")) {
        if (line.empty()) continue;
        const std::vector<std::string> f = split_on(line, '_');
        if (f.size() < 2) throw std::runtime_error("segments.txt: unexpected line " + line);
        if (f[1] != "gene") continue;
        if (f.size() < 5) throw std::runtime_error("segments.txt: unexpected line " + line);
        GeneTables& g = genes[f[2]];
        std::string id = f[4]; if (f.size() > 5) id += f[5];
        const bool intron = id.find("intron") != std::string::npos;
        if (!intron && id.find("exon") == std::string::npos) throw std::runtime_error("segments.txt: " + line + " is neither an exon nor an intron");
        (intron ? g.introns : g.exons).push_back(id); g.all.push_back(id);
        if (g.seq.count(id)) throw std::runtime_error("segments.txt: segment " + id + " of gene " + f[2] + " listed twice");
        const std::vector<std::string> lines = file_lines(prg_dir + "/PRG/" + line);
        const std::vector<std::string> head = split_on(lines.at(0), ' ');
        if (head.empty() || head[0] != "IndividualID") throw std::runtime_error(line + ": first field is not IndividualID");
        g.level_names[id] = std::vector<std::string>(head.begin() + 1, head.end());
        g.seq[id];
        for (size_t l = 1; l < lines.size(); l++) {
            if (lines[l].empty()) continue;
            const std::vector<std::string> fl = split_on(lines[l], ' ');
            if (fl.size() != head.size()) throw std::runtime_error(line + ": a row with " + std::to_string(fl.size()) + " fields, header has " + std::to_string(head.size()));
            g.seq[id][fl[0]] = std::vector<std::string>(fl.begin() + 1, fl.end());
        }
    }
    {   // parameters.txt (:701-710)
        std::ofstream ps((out_dir + "/parameters.txt").c_str());
        if (!ps.is_open()) throw std::runtime_error("cannot write into " + out_dir);
        const std::pair<double, double> er = sim.average_error_rates();
        ps << "graphDir: " << prg_dir << "\n" << "simulations_read_length: " << read_length << "\n" << "insertSize_mean: " << is_mean << "\n" << "insertSize_sd: " << is_sd << "\n"
           << "simulations_haploidCoverage: " << coverage << "\n" << "simulations_qualityMatrixFile: " << matrix << "\n" << "withError: " << with_error << "\n"
           << "rS average error rates: " << er.first << "\t" << er.second << "\n";
    }
    srand(seed);
    SimulatedIndividual res; std::vector<std::string> level_ids_out, haplotypes_out; bool first_write = true;
    for (const auto& gk : genes) {   // std::set<std::string> graphGenes: name order
        const std::string& gene = gk.first; const GeneTables& g = gk.second;
        const std::set<std::string> from_set = novel ? complete_exonic_types(g) : complete_genomic_types(g);
        const std::vector<std::string> from(from_set.begin(), from_set.end());
        if (from.empty()) throw std::runtime_error("gene " + gene + ": no complete type to choose from");
        std::vector<std::string> chosen; chosen.push_back(uniform_choice(from)); chosen.push_back(uniform_choice(from));
        std::vector<std::vector<std::string>> hap(2); std::vector<std::string> level_ids;
        for (const std::string& id : g.all) {
            const bool exon = id.find("intron") == std::string::npos;
            for (int h = 0; h < 2; h++) {
                std::string type = chosen[(size_t)h];
                if (!exon && novel) { std::vector<std::string> pool; for (const auto& kv : g.seq.at(id)) pool.push_back(kv.first); type = uniform_choice(pool); }
                const std::vector<std::string>& a = g.seq.at(id).at(type);
                hap[(size_t)h].insert(hap[(size_t)h].end(), a.begin(), a.end());
                if (h == 0) level_ids.insert(level_ids.end(), g.level_names.at(id).begin(), g.level_names.at(id).end());
            }
        }
        for (int h = 0; h < 2; h++) {
            if (hap[(size_t)h].size() != level_ids.size()) throw std::runtime_error("gene " + gene + ": haplotype and level names differ in length");
            for (const std::string& c : hap[(size_t)h]) if (c.size() != 1) throw std::runtime_error("gene " + gene + ": an allele of more than one character");
            const std::vector<SimulatedPair> pairs = sim.simulate_pairs_from_path(join_strings(hap[(size_t)h], ""), coverage, is_mean, is_sd, !with_error, "PRG_" + gene + "_HAPLO_" + std::to_string(h));
            write_simulated_pairs(pairs, out_dir + "/R", !first_write); first_write = false;
            res.pairs += (int64_t)pairs.size();
        }
        res.genes.push_back(gene); res.types.push_back(std::make_pair(chosen[0], chosen[1]));
        level_ids_out.push_back(join_strings(level_ids, ";")); haplotypes_out.push_back(join_strings(hap[0], ";") + "/" + join_strings(hap[1], ";"));
    }
    if (first_write) write_simulated_pairs(std::vector<SimulatedPair>(), out_dir + "/R", false);   // the four files exist even without a gene
    {   // HLAtypes.txt, haplotypes.txt (:895-930)
        std::ofstream ts((out_dir + "/HLAtypes.txt").c_str()), hs((out_dir + "/haplotypes.txt").c_str());
        if (!ts.is_open() || !hs.is_open()) throw std::runtime_error("cannot write into " + out_dir);
        std::vector<std::string> row{out_dir}, hhead, hrow{out_dir};
        for (size_t i = 0; i < res.genes.size(); i++) {
            row.push_back(res.types[i].first + "/" + res.types[i].second);
            hhead.push_back("LevelIDs_" + res.genes[i]); hhead.push_back("Haplotype_" + res.genes[i]); hrow.push_back(level_ids_out[i]); hrow.push_back(haplotypes_out[i]);
        }
        ts << "IndividualID" << "\t" << join_strings(res.genes, "\t") << "\n" << join_strings(row, "\t") << "\n";
        hs << "IndividualID" << "\t" << join_strings(hhead, "\t") << "\n" << join_strings(hrow, "\t") << "\n";
    }
    return res;
}

int64_t simulate_from_graph(const FlatGraph& g, const std::string& graph_label, const std::string& matrix, int read_length, double is_mean, double is_sd, int n_genomes,
                            const std::string& out_dir, double coverage, bool with_error, unsigned seed) {
    if (g.n_levels < 2 || g.node_out_off.size() != (size_t)g.n_nodes + 1) throw std::runtime_error("simulate_from_graph: graph without edges");
    ReadSimulator sim(matrix, (unsigned)read_length);
    {   // simulator.cpp:263-276
        std::ofstream ps((out_dir + "/parameters.txt").c_str());
        if (!ps.is_open()) throw std::runtime_error("cannot write into " + out_dir);
        const std::pair<double, double> er = sim.average_error_rates();
        ps << "Graph: " << graph_label << "\n" << "simulatedGraphGenomes: " << n_genomes << "\n" << "qualityMatrixFile: " << matrix << "\n" << "read_length: " << read_length << "\n"
           << "insertSize_mean: " << is_mean << "\n" << "insertSize_sd: " << is_sd << "\n" << "haploidCoverage: " << coverage << "\n" << "withError: " << with_error << "\n"
           << "rS average error rates: " << er.first << "\t" << er.second << "\n";
    }
    write_simulated_pairs(std::vector<SimulatedPair>(), out_dir + "/R", false);
    srand(seed);
    int64_t total = 0;
    for (int gi = 0; gi < n_genomes; gi++) {
        std::string path[2];
        for (int h = 0; h < 2; h++) {   // Graph.cpp:1485-1519
            int32_t node = g.level_node_off[0];
            while (g.node_out_off[(size_t)node + 1] > g.node_out_off[(size_t)node]) {
                const int n = g.node_out_off[(size_t)node + 1] - g.node_out_off[(size_t)node];
                int pick = (int)(((double)rand() / RAND_MAX) * n); if (pick == n) pick = n - 1;
                const int32_t e = g.node_out[(size_t)g.node_out_off[(size_t)node] + (size_t)pick];
                path[h].push_back((char)g.edge_emis[(size_t)e]); node = g.edge_to[(size_t)e];
            }
            if ((int32_t)path[h].size() != g.n_levels - 1) throw std::runtime_error("simulate_from_graph: a walk that ends before the last level");   // assert :1516
        }
        for (int h = 0; h < 2; h++) {   // both haplotypes are simulated before either is printed (simulator.cpp:333-343); the generators are per call, so the order is free
            const std::vector<SimulatedPair> pairs = sim.simulate_pairs_from_path(path[h], coverage, is_mean, is_sd, !with_error, "PRG_" + std::to_string(gi) + "h" + std::to_string(h + 1) + "_");
            write_simulated_pairs(pairs, out_dir + "/R", true); total += (int64_t)pairs.size();
        }
    }
    return total;
}

void write_simulated_pairs(const std::vector<SimulatedPair>& pairs, const std::string& prefix, bool append) {
    FILE* fq[2]; FILE* lv[2];
    for (int m = 0; m < 2; m++) {
        fq[m] = fopen((prefix + "_" + std::to_string(m + 1) + ".fq").c_str(), append ? "a" : "w"); lv[m] = fopen((prefix + "_" + std::to_string(m + 1) + ".levels").c_str(), append ? "a" : "w");
        if (!fq[m] || !lv[m]) throw std::runtime_error("cannot write " + prefix + "_*.fq / .levels");
    }
    for (const SimulatedPair& p : pairs) for (int m = 0; m < 2; m++) {
        const SimulatedRead& r = m ? p.second : p.first;
        fprintf(fq[m], "@%s\n%s\n+\n%s\n", r.name.c_str(), r.sequence.c_str(), r.quality.c_str());
        fprintf(lv[m], "@%s\n%s\n%s\n%s\n%s\n%s\n", r.name.c_str(), join_ints(r.levels).c_str(), r.edge_labels.c_str(), join_ints(r.full_levels).c_str(), r.full_labels.c_str(), r.full_sequence.c_str());
    }
    for (int m = 0; m < 2; m++) { fclose(fq[m]); fclose(lv[m]); }
}

} // namespace hlala
