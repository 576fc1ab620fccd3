// Per-edge word of the extension DP's first tier (csrc/extend_lean.h): edge_pack (from_z | to_z << 8 | emission << 16) plus two node flags,
// so that the tier learns from the edges it scans anyway whether a node starts a gap-path jump over >= 2 levels:
//   bit 24: the edge's FROM node has such a forward jump, bit 25: the edge's TO node has such a backward jump.
#pragma once
#include "prg_graph.h"
#include <vector>

namespace hlala {

inline std::vector<uint32_t> make_dp_pack(const FlatGraph& g) {
    std::vector<uint8_t> fwd((size_t)g.n_nodes, 0), bwd((size_t)g.n_nodes, 0);
    for (int32_t p = 0; p < g.n_paths; p++) if (g.path_off[p + 1] - g.path_off[p] >= 2) { fwd[g.path_from[p]] = 1; bwd[g.path_to[p]] = 1; }
    std::vector<uint32_t> pack((size_t)g.n_edges);
    for (int32_t e = 0; e < g.n_edges; e++) {
        const int32_t f = g.edge_from[e], t = g.edge_to[e];
        pack[e] = (uint32_t)(f - g.level_node_off[g.node_level[f]]) | ((uint32_t)(t - g.level_node_off[g.node_level[t]]) << 8) | ((uint32_t)g.edge_emis[e] << 16)
                  | ((uint32_t)fwd[f] << 24) | ((uint32_t)bwd[t] << 25);
    }
    return pack;
}

} // namespace hlala
