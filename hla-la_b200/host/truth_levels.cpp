// See truth_levels.h. Written against the behaviour of simulator/trueReadLevels.cpp (parser :203-320, evaluateAlignment :18-196).
#include "truth_levels.h"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <stdexcept>

namespace hlala {

namespace {

void chomp(std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); }   // Utilities::eraseNL

// Utilities::split(line, " ") + Utilities::StrtoI per token: every blank separates, an empty line has no tokens, a token that does not
// start with a number reads as 0 (operator>> of C++11 stores 0 on failure).
void parse_levels(const std::string& line, std::vector<int32_t>& out) {
    out.clear(); if (line.empty()) return;
    size_t s = 0;
    for (;;) {
        size_t p = line.find(' ', s);
        const std::string tok = line.substr(s, p == std::string::npos ? std::string::npos : p - s);
        char* end = nullptr; long v = strtol(tok.c_str(), &end, 10);
        out.push_back(end == tok.c_str() ? 0 : (int32_t)v);
        if (p == std::string::npos) break;
        s = p + 1;
    }
}

void read_one_file(const std::string& file, int mate, std::unordered_map<std::string, TruthLevels::Entry>& reads) {
    std::ifstream in(file.c_str());
    if (!in.is_open()) throw std::runtime_error("cannot open the true read levels " + file);
    std::string line;
    while (in.good()) {
        std::getline(in, line); chomp(line);
        if (line.empty()) continue;
        if (line[0] != '@') throw std::runtime_error(file + ": expected a line starting with @, got: " + line.substr(0, 60));
        TruthLevels::Mate& m = reads[line.substr(1)].mate[mate];
        if (!in.good()) throw std::runtime_error(file + ": record of " + line + " is cut off");
        std::string l2, l4;
        std::getline(in, l2); chomp(l2); parse_levels(l2, m.levels);
        std::getline(in, m.edge_labels); chomp(m.edge_labels);
        std::getline(in, l4); chomp(l4); parse_levels(l4, m.full_levels);
        std::getline(in, m.full_labels); chomp(m.full_labels);
        std::getline(in, m.full_seq); chomp(m.full_seq);
        m.present = true;
        if (m.full_levels.size() != m.full_labels.size() || m.full_levels.size() != m.full_seq.size())   // :311-312
            throw std::runtime_error(file + ": the full alignment of " + line + " has lines of different lengths");
    }
}

} // namespace

void TruthLevels::load(const std::string& r1_levels, const std::string& r2_levels) {
    reads.clear(); total = correct = reads_below_90 = 0;
    read_one_file(r1_levels, 0, reads);
    if (!r2_levels.empty()) read_one_file(r2_levels, 1, reads);
}

std::pair<int64_t, int64_t> TruthLevels::evaluate(const std::string& read_id, int n_mates, const int32_t* const* level, const uint8_t* const* schar,
                                                  const int32_t* n_cols, const uint8_t* reverse) {
    auto it = reads.find(read_id);
    if (it == reads.end()) throw std::runtime_error("no true levels for read " + read_id);   // assert(trueLevels.count(alignment.readID)), :40
    std::pair<int64_t, int64_t> r(0, 0);
    std::vector<int32_t> with_base, truth;
    for (int m = 0; m < n_mates; m++) {
        with_base.clear();
        for (int32_t i = 0; i < n_cols[m]; i++) if (schar[m][i] != '_') with_base.push_back(level[m][i]);   // :52-59
        truth = it->second.mate[m].levels;
        if (truth.size() != with_base.size())   // :62
            throw std::runtime_error("read " + read_id + " mate " + std::to_string(m + 1) + ": " + std::to_string(with_base.size()) + " aligned bases, " + std::to_string(truth.size()) + " true levels");
        if (reverse[m]) std::reverse(truth.begin(), truth.end());   // :64-67
        int64_t ok = 0;
        for (size_t i = 0; i < truth.size(); i++) ok += with_base[i] == truth[i];   // :71-80
        r.first += (int64_t)truth.size(); r.second += ok;
        if ((double)ok / (double)truth.size() < 0.9) reads_below_90++;   // :82-84 (the reference prints both alignments for these)
    }
    total += r.first; correct += r.second;   // :190-191
    return r;
}

} // namespace hlala
