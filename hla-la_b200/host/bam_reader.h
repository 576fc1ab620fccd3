// BAM ingest: one streaming pass over a bwa-mem BAM -> flat seed batch (include/hlala_b200.h: hlala_seed_batch_t).
//
// Replaces the BamTools-based seed collection of the reference:
//   processBAM::getReadIDs      mapper/processBAM.cpp:169-246   names with a record inside an interesting interval
//   processBAM::extractSeeds2   mapper/processBAM.cpp:703-862   per 10 000-name segment a rescan of the whole BAM; keeps mapped records with a
//                                                                CIGAR whose [Position, GetEndPosition(false, true)] lies inside the interval
//   protoSeeds::takeAlignment   mapper/reads/protoSeeds.cpp:24  grouped by read name into first-mate / other-mate lists, in file order
//   protoSeeds::isComplete      mapper/reads/protoSeeds.cpp:377 only pairs whose two mates both hold a primary record are aligned
//   (iteration of std::map<std::string, protoSeeds>)            pairs are processed in byte order of the read name
// What differs by design: one pass instead of a rescan per segment; BGZF blocks are inflated by a thread pool; records on contigs that are
// not PRG contigs (sequences.txt) are skipped, which is what the interval test does when reads were remapped against the PRG-only reference.
// Third-party arithmetic restated (BamTools is not in the reference tree; parity unpinned): GetEndPosition(false, true) =
// Position + sum of M,=,X,D,N lengths - 1; IsMapped = !(flag & 0x4); IsPrimaryAlignment = !(flag & 0x100); IsFirstMate = flag & 0x40;
// SEQ decoding "=ACMGRSVTWYHKDBN"; QUAL + 33; the AS tag as any integer aux type.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hlala {

struct BamBatch {
    // read 2p = first mate of pair p, read 2p+1 = the other mate; pairs in byte order of their names
    std::vector<std::string> pair_name;
    std::vector<int64_t> read_off; std::vector<uint8_t> bases, quals;
    std::vector<int32_t> chain_off, chain_contig, chain_pos, chain_as; std::vector<uint16_t> chain_flag;
    std::vector<int32_t> cigar_off; std::vector<uint32_t> cigar;
    // bookkeeping the reference prints (processBAM.cpp:858-859, 2386)
    int64_t records = 0, records_used = 0, names_seen = 0, pairs_incomplete = 0;
    double tlen_mean = 0, tlen_sd = 0; int64_t tlen_n = 0;   // gap between the mates of properly oriented primary pairs (a plain statistic of the BAM; NOT the reference's estimate)
    // The sample processBAM::estimateInsertSize aligns (processBAM.cpp:1071-1181 over extractSeeds(4000), :521-700): records are scanned contig by contig
    // in byte order of the contig NAMES (std::map of intervals), file order inside a contig, until 4000 read names have been seen and 2000 records
    // arrived at a then-complete pair (plus the first record of every later contig: the reference only leaves the current interval); of the pairs complete at that point, in name order, the first primary record of either mate (each with its
    // own SEQ/QUAL). One chain per read. is_loaded_contigs: contigs whose translation the reference has loaded by then (their anchors are the only
    // ones its distance computation sees, :836, :4441-4456).
    struct Sample {
        std::vector<std::string> pair_name;
        std::vector<int64_t> read_off; std::vector<uint8_t> bases, quals;
        std::vector<int32_t> chain_off, chain_contig, chain_pos, chain_as; std::vector<uint16_t> chain_flag;
        std::vector<int32_t> cigar_off; std::vector<uint32_t> cigar;
        std::vector<int32_t> loaded_contigs; int64_t names_seen = 0, incomplete = 0;
    } is_sample;
};

// contig_names: BAM reference name of every PRG contig in sequences.txt order (FlatGraph::contig_bam_name); contig_len: their lengths.
// Throws std::runtime_error on malformed input.
// long_reads: the selection of extractSeeds2 in long-read mode (primary records only, unpaired); the batch keeps the pair layout with an empty second read per name.
void read_bam_seeds(const std::string& path, const std::vector<std::string>& contig_names, const std::vector<int64_t>& contig_len, int threads, BamBatch& out, bool long_reads = false);

} // namespace hlala
