// k-mer index over the flat PRG: every k-mer spelled by a path of the graph, with the edge path(s) that spell it.
//
// Replaces GraphAndEdgeIndex::Index / queryIndex (Graph/GraphAndEdgeIndex.cpp:428-959, 986-997). The reference slides a window of
// attached (k-1)-mers through the graph level by level (edgeTargetCache / originalGraphAssignedKMers, :494-886) and records, per
// k-mer string, the edge paths in the order its loops meet them. What that procedure enumerates, stated directly:
//   * a k-mer path starts with a non-gap edge at a node that can be reached from the first node of level 0 and whose level is
//     <= n_levels-1-k (:475), traverses exactly k non-gap edges (gap edges in between are part of the path, forwardScanRec
//     :1061-1160) and ends with its k-th non-gap edge;
//   * recording order inside one k-mer string (it decides the order of findChains' results):
//       level 0 (:504-530): only the first node of level 0, paths in depth-first order of forwardScan(N0, k, -1), i.e.
//         lexicographic order of the edge sequence under the canonical edge order;
//       level >= 1 (:567-728): nodes of the level in canonical order; for one start node its attached (k-1)-mers in the order the
//         previous level's cache iteration (:799-806, std::map<Node*, std::map<string, ..>>) pushed them = by canonical order of the
//         LAST node of the (k-1)-mer path, then by the path's edge sequence (pointer strings of equal width compare like the
//         ordinals); for one attached (k-1)-mer its one-character extensions in depth-first order of forwardScan(last, 1, 0).
// k-mers are kept as raw emission bytes (the reference's std::map<std::string, ..> key), sorted like the map.
#pragma once
#include <cstdint>
#include <vector>
#include "prg_graph.h"

namespace hlala {

struct KmerIndex {
    int32_t k = 0;
    int64_t n_kmers = 0, n_pos = 0;
    std::vector<uint8_t> kmer_bytes;     // [n_kmers * k], ascending (std::map<std::string> order)
    std::vector<int64_t> kmer_pos_off;   // [n_kmers + 1] into the position arrays
    std::vector<int64_t> pos_edge_off;   // [n_pos + 1] into pos_edges
    std::vector<int32_t> pos_edges;      // flat edge ids of each path
    std::vector<int32_t> ht;             // open addressing (linear probing): k-mer id + 1, 0 = empty; size = ht_mask + 1
    uint32_t ht_mask = 0;
    int64_t find(const uint8_t* kmer) const;   // k-mer id or -1
};

// Throws std::runtime_error (k out of range, path explosion beyond max_edges).
// max_edges bounds the total length of all recorded paths: in highly polymorphic regions the number of k-mer paths grows exponentially
// with k (the reference's Index() has the same growth); the build fails loudly instead of exhausting memory.
void build_kmer_index(const FlatGraph& g, int k, KmerIndex& out, int64_t max_edges = (int64_t)1 << 30);

} // namespace hlala
