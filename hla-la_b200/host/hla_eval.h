// Evaluation of inferred HLA types against known ones: what `HLA-LA --action HLA ... --trueHLA <file>` does after the inference
// (HLA-LA.cpp:801-810). Reference: hla::HLATyper::read_inferred_types (hla/HLATyper.cpp:580-626), read_true_types (:628-688),
// evaluate_HLA_types (:407-529), alleles_compatible (:530-578), removeLocusAndStar (:337-351).
#pragma once
#include <map>
#include <set>
#include <string>
#include <utility>

namespace hlala {

typedef std::map<std::string, std::map<std::string, std::pair<std::set<std::string>, std::set<std::string>>>> InferredTypes;   // sample -> locus -> (alleles chr 1, alleles chr 2)
typedef std::map<std::string, std::map<std::string, std::pair<std::string, std::string>>> TrueTypes;                           // sample -> locus -> (allele 1, allele 2)

void read_inferred_types(const std::string& sample_id, InferredTypes& into, const std::string& bestguess_file);
void read_true_types(TrueTypes& into, const std::string& file);
bool alleles_compatible(std::string inferred, std::string truth);
// locus -> (alleles compared, alleles correct); `summary` receives the text the reference prints on stdout
std::map<std::string, std::pair<int, int>> evaluate_types(const TrueTypes& truth, const InferredTypes& inferred, std::string* summary);

} // namespace hlala
