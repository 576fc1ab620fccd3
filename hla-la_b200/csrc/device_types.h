// Device-side views of the flat graph and of a seed batch. Plain pointers into HBM; passed to kernels by value.
#pragma once
#include <cstdint>

namespace hlala {

struct alignas(16) I4 { int32_t x, y, z, w; };

struct DevGraph {
    int32_t n_levels, n_nodes, n_edges, n_contigs;
    const int32_t* level_node_off;   // [n_levels+1]
    const int32_t* level_edge_off;   // [n_levels+1]  (last real entry at n_levels-1; [n_levels] == n_edges)
    const uint32_t* edge_pack;       // from_z | to_z << 8 | emission << 16   (z = rank of the node inside its level)
    const uint32_t* dp_pack;         // edge_pack | long-jump flags of the edge's end nodes (host/dp_pack.h): what the first DP tier scans
    const uint32_t* lvl4;            // [n_levels * 4] per-level record of the first DP tier (extend_lean.h LnLvl): first edge | count << 26, first three dp_pack words
    const int32_t* edge_ord;         // flat edge -> canonical ordinal
    const int32_t* edge_from; const int32_t* edge_to;   // flat edge -> flat nodes (k-mer seeding)
    const int32_t* node_out_off; const int32_t* node_out;
    const int32_t* node_in_off; const int32_t* node_in;
    const int32_t* path_off; const int32_t* path_edges; const int32_t* path_from; const int32_t* path_to;
    const int32_t* jump_fwd_off; const int32_t* jump_fwd_path;
    const int32_t* jump_bwd_off; const int32_t* jump_bwd_path;
    // adjacency packed for the extension DP: one 16-byte record per node / edge / jump, so that a DP step costs one load each
    const I4* adj4;        // [n_nodes+1] {node_out_off, node_in_off, jump_fwd_off, jump_bwd_off}
    const I4* out_adj4;    // [n_edges] in node_out order: {flat edge, target node, edge_pack, 0}
    const I4* in_adj4;     // [n_edges] in node_in order:  {flat edge, source node, edge_pack, 0}
    const I4* jf4;         // [n_paths] in jump_fwd_path order: {path, target node, target z, path length}
    const I4* jb4;         // [n_paths] in jump_bwd_path order: {path, source node, source z, path length}
    const uint8_t* node_gapflags;   // [n_nodes] bit 0: has an outgoing '_' edge, bit 1: has an incoming '_' edge
    const uint8_t* gap_stretch;      // [n_levels-1]
    const int64_t* contig_off;       // [n_contigs+1]
    const uint8_t* contig_seq;
    const int32_t* contig_level;
    const int32_t* contig_prg_id;    // [n_contigs]
    const int32_t* anchor_off;       // [n_levels+1]
    const int32_t* anchor_prg_id; const int32_t* anchor_pos;
};

struct DevBatch {
    int64_t n_reads; int32_t n_chains;
    const int64_t* read_off; const uint8_t* bases; const uint8_t* quals;
    const int32_t* chain_off;
    const int32_t* chain_contig; const int32_t* chain_pos; const uint16_t* chain_flag; const int32_t* chain_as;
    const int32_t* cigar_off; const uint32_t* cigar;
    const int32_t* chain_order;      // slot -> input chain (AS-sorted inside each read, processBAM.cpp:1945)
    const int32_t* read_primary;     // [n_reads] slot of the first primary record (protoSeeds.cpp:252-314), -1 if none
    const int32_t* slot_read;        // [n_chains] read of a slot
};

// Score tables computed once on the host with the same libm calls as the reference, so that every per-column
// term is bit-identical to extensionAligner::scoreOneAlignment (extensionAligner.cpp:52-182):
struct ScoreTables {
    double log_match[256];      // log(pCorrect(q)) with the 0.999 cap and the 0 -> 1e-5 floor
    double log_mismatch[256];   // log((1 - pCorrect(q)) * (1/3))
    double rate_match_mismatch; // log(1 - 2 * 0.001)
    double rate_deletion;       // log(0.001)
    double ins_term;            // log(0.001) + log(1/4)
};

// per-chain status: 0 aligned, 1 skipped (other strand than the primary), negative = include/hlala_b200.h HLALA_E_* code
enum : int32_t { CH_OK = 0, CH_SKIPPED_STRAND = 1 };
#define HLALA_E_ARG_DEV (-1)
#define HLALA_E_CAPACITY_DEV (-4)
#define HLALA_E_INVARIANT_DEV (-5)

} // namespace hlala
