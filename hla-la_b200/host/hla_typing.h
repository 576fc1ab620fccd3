// HLA typing stage, host side: tables, exon projection, read filters, calls and the hla/*.txt writers. The two hot loops of the
// stage — per-read x allele-cluster log-likelihoods and allele-pair log-likelihood sums — run on the GPU (csrc/typing_kernels.cu)
// behind the TypingDevice interface; this file never computes them (there is no CPU fallback).
//
// Reference (paths relative to the reference tree); short-read paired mode and long-read unpaired mode (HLATyper.cpp:935-946, 1079-1097, 1467-1495,
// 1797-1880, 1918, 3568-3930):
//   gene filter                                   mapper/processBAM.cpp:2427-2446, hla/HLATyper.cpp:259
//   HLATyper::HLATyper (segments, gene bounds)     hla/HLATyper.cpp:36-256, Graph::readGraphLoci Graph/Graph.cpp:2563
//   HLATyper::HLATypeInference                     hla/HLATyper.cpp:933-2810
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace hlala {

struct TypingLocus {
    std::string name;
    std::vector<int32_t> col_level, col_exon, col_exonpos;   // per exon column, exons concatenated (HLATyper.cpp:1211-1260)
    std::vector<int32_t> exon_len;
    int32_t lmin = -1, lmax = -1;
    std::vector<std::vector<std::string>> cluster_members;   // HLA types with identical exon sequence, std::set order (HLATyper.cpp:1323-1372)
    std::vector<std::string> cluster_seq;                    // one string of P symbols per cluster
    int32_t P() const { return (int32_t)col_level.size(); }
    int32_t C() const { return (int32_t)cluster_seq.size(); }
};

struct TypingTables {
    std::string prg_dir;
    std::vector<TypingLocus> loci;                            // the 17 loci of HLATyper.cpp:42, in that order
    std::vector<std::pair<int32_t, int32_t>> gene_bounds;     // [first,last] level of every gene named in segments.txt (gene filter)
    std::map<std::string, std::string> alleles_to_G; std::set<std::string> G_loci; bool g_loaded = false;
    void load(const std::string& prg_graph_dir);
    void load_G(const std::string& dir);                      // hla_nom_g.txt (HLATyper.cpp:4150-4198); the reference reads it from the CWD
};

// Alignments of the read pairs that passed the gene filter, compact: reads 2i / 2i+1 are the mates of entry i.
struct TypingReads {
    std::vector<int64_t> pair_id;        // index of the pair in the batch it came from (names default to "r<id>")
    std::vector<std::string> name;       // per pair (both mates carry the same BAM QNAME)
    std::vector<int64_t> col_off;        // [2n+1] into the column arrays
    std::vector<int32_t> level; std::vector<uint8_t> g, s, mq;
    std::vector<int64_t> base_off;       // [2n+1] into bases / quals (BAM orientation)
    std::vector<uint8_t> bases, quals;
    std::vector<uint8_t> reverse; std::vector<double> mapq;   // per read
    bool long_reads = false;             // long-read mode: every entry is ONE read (reads 2i) followed by an empty mate (no columns, no bases)
    size_t n_pairs() const { return pair_id.size(); }
    void clear();
    void append(const TypingReads& o);
    std::vector<uint8_t> serialize() const;
    void deserialize_append(const uint8_t* p, size_t n);
};

struct LocusDeviceInput {     // what the GPU needs for one locus
    int32_t C = 0, P = 0, R = 0;
    bool long_reads = false;             // indel probabilities 0.075 instead of 0.001 (HLATyper.cpp:935-946)
    const std::vector<std::string>* cluster_seq = nullptr;
    std::vector<int32_t> rec_off;        // [R+1]
    std::vector<int16_t> rec_pos;        // exon column
    std::vector<uint8_t> rec_c0, rec_q0; // first genotype character ('_' for a gap) and its quality character
    std::vector<uint16_t> rec_glen;      // genotype length (1 + inserted bases)
};
struct LocusDeviceOutput {    // reference layouts
    std::vector<double> LL; std::vector<int32_t> mism;              // [C*R], LL[c*R + r]   (HLATyper.cpp:2049-2277)
    std::vector<double> pair_ll, pair_mavg, pair_mmin;              // [C(C+1)/2], c1 <= c2 in loop order (HLATyper.cpp:2280-2364)
};
class TypingDevice {
public:
    virtual ~TypingDevice() {}
    // want_read_ll: also return LL / mism (parity tests); the pair sums are always returned
    virtual void run_locus(const LocusDeviceInput& in, bool want_read_ll, LocusDeviceOutput& out) = 0;
    // A locus failed on this rank before its device stage: with several ranks the per-locus collective must still be entered (the other ranks are
    // waiting in it) and must tell them — the implementation contributes NaNs, which run_locus reports as an error on every rank.
    virtual void abort_locus(int32_t /*C*/) {}
};

struct LocusCall { std::string locus; int32_t C = 0, R = 0; std::string call1, call2; double q1 = 0, q2 = 0; LocusDeviceOutput dev; };

struct TypingOptions { bool keep_read_ll = false; int threads = 0; };     // (long-read mode travels with the reads: TypingReads::long_reads)   // threads: host threads over loci; 0 = one per locus up to the core count, 1 = everything on the calling thread

// Gene filter predicate (processBAM.cpp:2427-2446): either mate's [first,last] level interval overlaps a gene.
bool pair_overlaps_genes(const TypingTables& T, int32_t f1, int32_t l1, int32_t f2, int32_t l2);

// Long-read mode: the reads whose chosen alignment overlaps a gene (processBAM.cpp:2297-2309), packed as TypingReads entries (read + empty mate). Alignment arrays
// are [n_reads, cap] as hlala_align_long_reads returns them; names: one per read or nullptr ("r<index>").
TypingReads long_read_typing_input(const TypingTables& T, int64_t n_reads, const char* const* names, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int32_t cap,
                                   const int32_t* n_cols, const int32_t* level, const uint8_t* g, const uint8_t* s, const uint8_t* mq, const uint8_t* reverse, const double* mapq);

// HLATypeInference on the included pairs; writes out_dir/{R1_bestguess.txt, R1_bestguess_G.txt, R1_PP_<L>_pairs.txt, R1_pileup_<L>.txt,
// R1_readIDs_<L>.txt, R1_columnIncompatibilities_<L>.txt, summaryStatistics.txt, histogram_matchesPerRead.txt, R1_parameters.txt}.
void run_typing(TypingTables& T, const TypingReads& reads, double is_mean, double is_sd, const std::string& out_dir, const std::string& g_dir,
                TypingDevice& dev, const TypingOptions& opt, std::vector<LocusCall>& calls);

} // namespace hlala
