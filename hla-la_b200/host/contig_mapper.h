// FASTQ -> seed batch without bwa: stands where the reference runs `bwa mem -a -M <PRG>/mapping_PRGonly/referenceGenome.fa R1 R2`
// (mapper/bwa/BWAmapper.cpp map(), called at HLA-LA.cpp:742-779) and reads the result back through extractSeeds2. Like bwa it maps every read against
// the PRG's linear contigs (haplotypes and allele sequences of sequences.txt) and reports all placements (-a; here the 64 best-scoring ones of a read at most), the best one as the primary record, the
// others as secondary ones (protoSeeds.cpp:252-314 relies on exactly one primary per mate), each with a CIGAR and a bwa-style score
// (match 1, mismatch 4, gap open 6 + extend 1, clipping 5, minimum score 30: bwa mem's defaults) that plays the part of the AS tag (processBAM.cpp:4314-4336).
// Method: sampled k-mer index of the contigs, votes per (contig, strand, diagonal), banded affine-gap alignment of every candidate with free, penalised
// clipping at either end. It is NOT bwa: there is no oracle for it (bwa is an external program and absent here), placements and scores of hard cases
// may differ. What is measured instead is what the reference's own testPRGMapping measures: bases of simulated reads on their true level after the
// full alignment path (tests/test_zz_fastq_mapper.py). Host code (threads); a CUDA version of the voting + banded alignment is the natural next step.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bam_reader.h"
#include "prg_graph.h"

namespace hlala {

struct MapperParams {
    int k = 19, ref_step = 4, read_step = 1;      // k-mers of every ref_step-th contig position are indexed, every read_step-th read k-mer is looked up
    int max_occ = 4000;                           // k-mers with more placements are skipped (used again only if nothing else votes)
    int min_votes = 2, max_candidates = 64, band = 24;
    int match = 1, mismatch = 4, gap_open = 6, gap_extend = 1, clip = 5, min_score = 30;
    int rescue_window = 800;
};

struct Placement { int32_t contig = 0, pos = 0; bool reverse = false; int32_t score = 0; std::vector<uint32_t> cigar; };   // cigar: BAM packed ops, S for clipped bases

class ContigMapper {
public:
    ContigMapper(const FlatGraph& g, const MapperParams& p = MapperParams());
    // all placements of one read (sequence as sequenced), best first; ties keep (strand, contig, position) order
    std::vector<Placement> map_read(const std::string& seq) const;
    // bwa mem's mate rescue (mem_matesw): a read without a placement of its own is aligned, on the opposite strand, to the stretch of the contig within
    // `rescue_window` bases of its mate's primary placement
    bool rescue(const std::string& seq, const Placement& mate, Placement& out) const;
    const MapperParams& params() const { return p_; }
    int64_t index_entries() const { return (int64_t)keys_.size(); }

private:
    bool align(const uint8_t* read, int len, int32_t contig, int64_t diag, Placement& out, int force_band = 0) const;
    const FlatGraph& g_; MapperParams p_;
    std::vector<uint64_t> keys_; std::vector<uint32_t> pos_, ctg_;   // sorted by key; pos = offset into FlatGraph::contig_seq, ctg = its contig
    std::vector<uint32_t> bucket_; int bucket_shift_ = 0;             // first index entry of every value of the key's top bits
};

// Paired FASTQ files (plain or gzip) -> the batch hlala_bam_read would return for their `bwa mem -a -M` BAM: pairs in byte order of their names, reads 2p / 2p+1
// the first / second mate, SEQ / QUAL in the orientation of the primary placement, chains in (contig, position) order. Pairs with an unplaced mate are dropped and counted.
void map_fastq_pairs(const FlatGraph& g, const std::string& fastq1, const std::string& fastq2, int threads, const MapperParams& p, BamBatch& out);

} // namespace hlala
