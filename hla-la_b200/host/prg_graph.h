// Flat (structure-of-arrays) population reference graph, as laid out in HBM.
//
// Replaces the reference's pointer graph (Graph/Graph.h:56-153, Node.h:58-89, Edge.h:31-65) and the derived
// tables the aligner builds on top of it:
//   * canonical order. The reference iterates std::set<Node*>/std::set<Edge*> in pointer order; its graph.txt
//     writer emits nodes and edges in exactly that order (Graph.cpp:2236-2310) and its reader allocates them in
//     file order (Graph.cpp:2428-2433, 2515-2524). We therefore define the canonical ("ordinal") order as order of
//     appearance in graph.txt; every "first maximum" / "first taken edge" rule below resolves against it.
//   * z index = rank of a node inside its level in canonical order (alignerBase.cpp:27-37).
//   * nodes are stored sorted by (level, ordinal), edges by (level, ordinal): all edges leaving level l are the
//     contiguous range [level_edge_off[l], level_edge_off[l+1]) and, inside a level, array order == set<Edge*> order.
//   * gap-edge paths and jump lists: Graph::computeGapEdgePaths (Graph.cpp:347-476).
//   * gap stretches: processBAM ctor (processBAM.cpp:91-149).
//   * contigs + translation + level anchors: processBAM ctor / _loadMapping (processBAM.cpp:85-88, 4389-4459).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hlala {

struct FlatGraph {
    int32_t n_levels = 0;   // == NodesPerLevel.size(); edges connect level l -> l+1 for l in [0, n_levels-1)
    int32_t n_nodes = 0, n_edges = 0;
    int32_t max_nodes_per_level = 0, max_edges_per_level = 0;

    std::vector<int32_t> level_node_off;   // n_levels+1
    std::vector<int32_t> node_ord;         // canonical ordinal of flat node i
    std::vector<int32_t> level_edge_off;   // n_levels (+1 sentinel): edges leaving level l
    std::vector<int32_t> edge_from, edge_to;   // flat node indices
    std::vector<uint8_t> edge_emis;        // 'A','C','G','T','N','_','*'
    std::vector<int32_t> edge_ord;         // canonical ordinal of flat edge e (what the parity tests compare)
    std::vector<int32_t> ord_to_edge;      // inverse
    std::vector<int32_t> node_out_off, node_out;   // per node: outgoing flat edges in ordinal order
    std::vector<int32_t> node_in_off, node_in;     // per node: incoming flat edges in ordinal order

    int32_t n_paths = 0;
    std::vector<int32_t> path_off, path_edges;     // completedGapEdgePaths as flat edges
    std::vector<int32_t> path_from, path_to;       // flat nodes
    std::vector<int32_t> jump_fwd_off, jump_fwd_path;   // per node, in std::map<Node*,Edge*> order of the target
    std::vector<int32_t> jump_bwd_off, jump_bwd_path;

    std::vector<uint8_t> gap_stretch;      // n_levels-1

    std::vector<std::string> level_names;  // locus id of each edge level (reads_per_level.txt)

    // contigs (BAM references), in sequences.txt order
    int32_t n_contigs = 0;
    std::vector<int32_t> contig_prg_id;
    std::vector<std::string> contig_bam_name;
    std::vector<int64_t> contig_off;       // n_contigs+1 into contig_seq / contig_level
    std::vector<uint8_t> contig_seq;
    std::vector<int32_t> contig_level;     // translation: contig position -> PRG level
    std::vector<int32_t> contig_tr_len;    // entries the reference's parser would hold (incl. the bogus trailing 0)
    // graphLevel_2_underlyingSequencePositions as CSR over levels; entries sorted by PRG id (std::map<int,int> order)
    std::vector<int32_t> anchor_off;       // n_levels+1
    std::vector<int32_t> anchor_prg_id, anchor_pos;

    int32_t z_of(int32_t node) const;      // needs level lookup; provided for host code/tests
    std::vector<int32_t> node_level;       // level of flat node i
};

// Parse PRG_graph_dir (PRG/graph.txt, sequences.txt, translation/*.txt, mapping_PRGonly/referenceGenome.fa).
// Throws std::runtime_error with the reference's own wording where it has one.
void load_prg_dir(const std::string& dir, FlatGraph& g);

} // namespace hlala
