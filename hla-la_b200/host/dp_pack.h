// Per-edge word of the extension DP's first tier (csrc/extend_lean.h): edge_pack (from_z | to_z << 8 | emission << 16) plus two node flags,
// so that the tier learns from the edges it scans anyway whether a node starts a gap-path jump over >= 2 levels:
//   bit 24: the edge's FROM node has such a forward jump, bit 25: the edge's TO node has such a backward jump.
#pragma once
#include "prg_graph.h"
#include <stdexcept>
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

// extend_lean.h LnLvl, as four words per level: first edge | min(edge count, 63) << 26, dp_pack of the first three edges; levels [0, n_levels)
inline std::vector<uint32_t> make_lvl4(const FlatGraph& g, const std::vector<uint32_t>& dp_pack) {
    std::vector<uint32_t> r((size_t)g.n_levels * 4, 0u);
    for (int32_t l = 0; l < g.n_levels; l++) {
        const int32_t ea = l < (int32_t)g.level_edge_off.size() ? g.level_edge_off[l] : g.n_edges;
        const int32_t eb = l + 1 < (int32_t)g.level_edge_off.size() ? g.level_edge_off[l + 1] : g.n_edges;
        const int32_t n = eb - ea;
        if (ea >= (1 << 26)) throw std::runtime_error("graph has more than 2^26 edges: the level records of the extension DP hold 26-bit edge offsets");
        r[(size_t)l * 4] = (uint32_t)ea | ((uint32_t)(n > 63 ? 63 : n) << 26);
        for (int k = 0; k < 3 && k < n; k++) r[(size_t)l * 4 + 1 + k] = dp_pack[(size_t)ea + k];
    }
    return r;
}

} // namespace hlala
