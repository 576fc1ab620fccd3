#include "hla_eval.h"

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace hlala {

namespace {
[[noreturn]] void violated(const std::string& what) { throw std::runtime_error("type evaluation: reference assertion would fail: " + what); }
#define EV_REQUIRE(c, what) do { if (!(c)) violated(what); } while (0)

std::vector<std::string> fields(const std::string& s, const std::string& d) {     // Utilities::split (Utilities.cpp:610): "" -> no fields
    std::vector<std::string> out; if (s.empty()) return out;
    size_t at = 0;
    for (;;) { const size_t q = s.find(d, at); if (q == std::string::npos) { out.emplace_back(s, at); break; } out.emplace_back(s, at, q - at); at = q + d.size(); }
    return out;
}
void strip_nl(std::string& s) { while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); }
std::string after_star(const std::string& t) { const size_t p = t.find('*'); return p == std::string::npos ? t : t.substr(p + 1); }
}

void read_inferred_types(const std::string& sample_id, InferredTypes& into, const std::string& file) {
    EV_REQUIRE(into.count(sample_id) == 0, "sample not read before");
    std::ifstream in(file); EV_REQUIRE(in.is_open() && in.good(), "R1_bestguess.txt readable: " + file);
    std::string line; std::getline(in, line); strip_nl(line);
    const std::vector<std::string> head = fields(line, "\t");
    EV_REQUIRE(head.size() > 2 && head[0] == "Locus" && head[1] == "Chromosome" && head[2] == "Allele", "header Locus / Chromosome / Allele");
    while (in.good()) {
        std::getline(in, line); strip_nl(line); if (line.empty()) continue;
        const std::vector<std::string> f = fields(line, "\t"); EV_REQUIRE(f.size() == head.size(), "field count of a bestguess line");
        const int chrom = atoi(f[1].c_str()); EV_REQUIRE(chrom == 1 || chrom == 2, "chromosome is 1 or 2");
        const std::vector<std::string> alleles = fields(f[2], ";"); EV_REQUIRE(!alleles.empty(), "at least one allele");
        std::set<std::string>& dst = chrom == 1 ? into[sample_id][f[0]].first : into[sample_id][f[0]].second;
        EV_REQUIRE(dst.empty(), "one line per locus and chromosome");
        dst.insert(alleles.begin(), alleles.end());
    }
}

void read_true_types(TrueTypes& into, const std::string& file) {
    std::ifstream in(file); EV_REQUIRE(in.is_open() && in.good(), "true types readable: " + file);
    std::string line; std::getline(in, line); strip_nl(line);
    bool space = false; std::vector<std::string> head = fields(line, "\t");
    if (head.size() == 1) { space = true; head = fields(line, " "); EV_REQUIRE(head.size() > 1, "header has loci"); }
    EV_REQUIRE(head.size() > 1 && head[0] == "IndividualID", "header starts with IndividualID");
    while (in.good()) {
        std::getline(in, line); strip_nl(line); if (line.empty()) continue;
        const std::vector<std::string> f = fields(line, space ? " " : "\t"); EV_REQUIRE(f.size() == head.size(), "field count of a true-types line");
        EV_REQUIRE(into.count(f[0]) == 0, "individual listed once");
        for (size_t i = 1; i < head.size(); i++) {
            std::string locus = head[i];
            if (locus.size() > 4 && locus.compare(0, 4, "HLA-") == 0) locus = locus.substr(4);
            if (locus.size() > 3 && locus.compare(0, 3, "HLA") == 0) locus = locus.substr(3);
            const std::vector<std::string> al = fields(f[i], "/"); EV_REQUIRE(al.size() == 2, "two alleles separated by /");
            into[f[0]][locus].first = al[0]; into[f[0]][locus].second = al[1];
        }
    }
}

bool alleles_compatible(std::string inferred, std::string truth) {
    EV_REQUIRE(!inferred.empty() && !truth.empty(), "alleles not empty");
    if (truth.back() == 'G' || truth.back() == 'g') truth.pop_back();
    EV_REQUIRE(inferred.find(':') != std::string::npos, "inferred allele has fields: " + inferred);
    if (truth.find(':') == std::string::npos) { EV_REQUIRE(truth.size() == 4, "four-digit true allele: " + truth); truth = truth.substr(0, 2) + ":" + truth.substr(2); }
    const std::vector<std::string> gi = fields(inferred, ":"), gt = fields(truth, ":");
    EV_REQUIRE(gi.size() >= 2 && gt.size() >= 2, "at least two fields");
    if (gi.size() < gt.size()) return false;
    for (size_t k = 0; k < gt.size(); k++) if (gi[k] != gt[k]) return false;
    return true;
}

std::map<std::string, std::pair<int, int>> evaluate_types(const TrueTypes& truth, const InferredTypes& inferred, std::string* summary) {
    std::map<std::string, std::pair<int, int>> res;
    for (const auto& ind : inferred) {
        const auto tind = truth.find(ind.first);
        for (const auto& loc : ind.second) {
            if (tind == truth.end()) continue;
            const auto tl = tind->second.find(loc.first); if (tl == tind->second.end()) continue;
            const std::string t1 = after_star(tl->second.first), t2 = after_star(tl->second.second);
            // one side: the first inferred allele (set order) compatible with the true one counts once
            auto side = [](const std::set<std::string>& inf, const std::string& tr, int& correct, int& compared) {
                if (inf.empty()) return;
                for (const std::string& a : inf) if (alleles_compatible(after_star(a), tr)) { correct++; break; }
                compared++;
            };
            int best_correct = 0, best_compared = 0;
            for (int swap = 0; swap < 2; swap++) {
                int correct = 0, compared = 0;
                side(swap ? loc.second.second : loc.second.first, t1, correct, compared);
                side(swap ? loc.second.first : loc.second.second, t2, correct, compared);
                if (swap == 0 || correct > best_correct) { best_correct = correct; best_compared = compared; }
            }
            EV_REQUIRE(best_correct <= best_compared, "correct <= compared");
            res[loc.first].first += best_compared; res[loc.first].second += best_correct;
        }
    }
    std::ostringstream o;
    o << "HLATyper::evaluate_HLA_types(..) summary:\n";
    for (const auto& kv : res) {
        o << "\t" << kv.first << "\n";
        EV_REQUIRE(kv.second.first > 0, "alleles were evaluated");
        const double rate = kv.second.second / kv.second.first;      // integer division, as in the reference (HLATyper.cpp:522)
        o << "\t\t" << kv.second.first << " evaluated." << "\n" << "\t\t" << kv.second.second << " correct." << "\n" << "\t\t" << (rate * 100) << "%" << "\n";
    }
    if (summary) *summary = o.str();
    return res;
}

} // namespace hlala
