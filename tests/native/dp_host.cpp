// Test-only: runs hla-la_b200/csrc/extend_dp.h (the very code the GPU executes) on the host, so the CPU test-suite
// can compare the extension DP with the compiled reference without a GPU.
#include "../../hla-la_b200/host/prg_graph.h"
#include "../../hla-la_b200/csrc/extend_dp.h"
#include "../../hla-la_b200/csrc/extend_lean.h"
#include "../../hla-la_b200/host/dp_pack.h"
#include <memory>
#include <string>
#include <vector>
#include <cstring>

using namespace hlala;

struct DpHost { FlatGraph g; std::vector<uint32_t> pack; DpGraph view; std::vector<unsigned char> scratch;
                std::vector<uint32_t> dp_pack; std::vector<int32_t> leo; LnGraph ln; std::vector<LnRec> rec; std::vector<LnAhead> ahead; std::vector<uint32_t> lvl4; long long ln_steps = 0; };
template <int WORDS> struct HostWords { uint32_t w[WORDS]; uint32_t& operator()(int i) { return w[i]; } uint8_t& b(int word, int byte) { return ((uint8_t*)&w[word])[byte]; } };

extern "C" {
void* dp_host_open(const char* dir) {
    try {
        std::unique_ptr<DpHost> h(new DpHost()); load_prg_dir(dir, h->g);
        const FlatGraph& g = h->g; h->pack.resize(g.n_edges);
        for (int e = 0; e < g.n_edges; e++) { int f = g.edge_from[e], t = g.edge_to[e];
            h->pack[e] = (uint32_t)(f - g.level_node_off[g.node_level[f]]) | ((uint32_t)(t - g.level_node_off[g.node_level[t]]) << 8) | ((uint32_t)g.edge_emis[e] << 16); }
        DpGraph& v = h->view; v.n_levels = g.n_levels; v.level_node_off = g.level_node_off.data(); v.edge_pack = h->pack.data();
        v.node_out_off = g.node_out_off.data(); v.node_out = g.node_out.data(); v.node_in_off = g.node_in_off.data(); v.node_in = g.node_in.data();
        v.path_off = g.path_off.data(); v.path_edges = g.path_edges.data(); v.path_from = g.path_from.data(); v.path_to = g.path_to.data();
        v.jump_fwd_off = g.jump_fwd_off.data(); v.jump_fwd_path = g.jump_fwd_path.data(); v.jump_bwd_off = g.jump_bwd_off.data(); v.jump_bwd_path = g.jump_bwd_path.data();
        h->scratch.resize(dp_scratch_bytes());
        h->dp_pack = make_dp_pack(g); h->leo = g.level_edge_off; h->leo.resize((size_t)g.n_levels + 1, g.n_edges);
        LnGraph& l = h->ln; l.n_levels = g.n_levels; l.level_node_off = g.level_node_off.data(); l.level_edge_off = h->leo.data(); l.dp_pack = h->dp_pack.data();
        l.path_off = g.path_off.data(); l.path_edges = g.path_edges.data(); l.path_from = g.path_from.data(); l.path_to = g.path_to.data();
        l.jump_fwd_off = g.jump_fwd_off.data(); l.jump_fwd_path = g.jump_fwd_path.data(); l.jump_bwd_off = g.jump_bwd_off.data(); l.jump_bwd_path = g.jump_bwd_path.data();
        h->lvl4 = make_lvl4(g, h->dp_pack); l.lvl4 = (const LnLvl*)h->lvl4.data();
        h->rec.resize(LN_CELLS + 1); h->ahead.assign(LN_AHEAD + LN_RING / 8, LnAhead{0xDEADBEEFu, 0xDEADBEEFu});   // the tier clears its table itself
        return h.release();
    } catch (...) { return nullptr; }
}
// seed_edge_ord: canonical ordinal of the seed's first (left, pos=0) or last (right, pos=1) edge.
int dp_host_extend(void* hv, const uint8_t* seq, int seq_len, int start_seq, int seed_edge_ord, int pos, int32_t* out_edge_ord, uint8_t* out_s, int32_t* n_cols, int32_t* far_y, int32_t* applicable) {
    DpHost* h = (DpHost*)hv; const FlatGraph& g = h->g;
    int e = g.ord_to_edge[seed_edge_ord];
    int node = pos ? g.edge_to[e] : g.edge_from[e]; int level = g.node_level[node]; int z = node - g.level_node_off[level];
    *applicable = pos ? (level < g.n_levels - 1) : (level > 0);   // extensionAligner.cpp:224,275
    *n_cols = 0; *far_y = start_seq;
    if (!*applicable) return 0;
    DpScratch S = dp_carve(h->scratch.data());
    std::vector<int32_t> oe(DP_EXT_CAP); DpResult r;
    int rc = dp_extend(h->view, seq, seq_len, start_seq, level, z, pos != 0, S, oe.data(), out_s, r);
    if (rc) return rc;
    for (int i = 0; i < r.n_cols; i++) out_edge_ord[i] = oe[i] >= 0 ? g.edge_ord[oe[i]] : -1;
    *n_cols = r.n_cols; *far_y = r.far_y;
    return 0;
}
}
// the first GPU tier (extend_lean.h) on the host; returns 0, DP_DEFER (-100) or a negative code
template <class CFG> static int run_lean(void* hv, const uint8_t* seq, int seq_len, int start_seq, int seed_edge_ord, int pos, int32_t* out_edge_ord, uint8_t* out_s, int32_t* n_cols, int32_t* far_y, int32_t* applicable) {
    DpHost* h = (DpHost*)hv; const FlatGraph& g = h->g;
    int e = g.ord_to_edge[seed_edge_ord];
    int node = pos ? g.edge_to[e] : g.edge_from[e]; int level = g.node_level[node]; int z = node - g.level_node_off[level];
    *applicable = pos ? (level < g.n_levels - 1) : (level > 0);
    *n_cols = 0; *far_y = start_seq;
    if (!*applicable) return 0;
    typedef HostWords<CFG::WORDS> SM; typedef LnDp<CFG, SM> DP;
    SM S; LnState st;
    int rc = DP::init(h->ln, S, st, h->rec.data(), seq, seq_len, start_seq, level, z, pos != 0);
    if (rc) return rc;
    for (;;) { rc = DP::step(h->ln, S, st, h->rec.data(), h->ahead.data()); h->ln_steps++; if (rc) break; }
    for (int i = 0; i < CFG::TD; i++) if (S(CFG::TK + i) != LN_EMPTY) return -77;   // the touch table must be left empty
    if (rc != 1) return rc;
    std::vector<int32_t> oe(DP_EXT_CAP); DpResult r;
    rc = DP::finish(h->ln, st, h->rec.data(), oe.data(), out_s, r);
    if (rc) return rc;
    for (int i = 0; i < r.n_cols; i++) out_edge_ord[i] = oe[i] >= 0 ? g.edge_ord[oe[i]] : -1;
    *n_cols = r.n_cols; *far_y = r.far_y;
    return 0;
}
extern "C" int dp_host_extend_lean(void* hv, const uint8_t* seq, int seq_len, int start_seq, int seed_edge_ord, int pos, int32_t* out_edge_ord, uint8_t* out_s, int32_t* n_cols, int32_t* far_y, int32_t* applicable) {
    return run_lean<LnStd>(hv, seq, seq_len, start_seq, seed_edge_ord, pos, out_edge_ord, out_s, n_cols, far_y, applicable);
}
extern "C" int dp_host_extend_lean_big(void* hv, const uint8_t* seq, int seq_len, int start_seq, int seed_edge_ord, int pos, int32_t* out_edge_ord, uint8_t* out_s, int32_t* n_cols, int32_t* far_y, int32_t* applicable) {
    return run_lean<LnBig>(hv, seq, seq_len, start_seq, seed_edge_ord, pos, out_edge_ord, out_s, n_cols, far_y, applicable);
}
#ifdef HLALA_LN_REASONS
extern "C" void dp_host_ln_reasons(long long* out) { for (int i = 0; i < 16; i++) { out[i] = ln_reasons()[i]; ln_reasons()[i] = 0; } }
#endif
extern "C" long long dp_host_ln_steps(void* hv) { DpHost* h = (DpHost*)hv; long long n = h->ln_steps; h->ln_steps = 0; return n; }
extern "C" int dp_host_key_less_check(int x1, int z1, int x2, int z2) { return (int)dp_key_less(x1, z1, x2, z2) == (int)ln_key_less(x1, z1, x2, z2); }
#ifdef HLALA_DP_STATS
extern "C" void dp_host_stats(long long* out) { DpStats& s = dp_stats(); out[0] = s.ext; out[1] = s.diags; out[2] = s.touched; out[3] = s.m1; out[4] = s.m2; out[5] = s.cells; out[6] = s.diags_all_at_end; out[7] = s.max_td; out[8] = s.max_m1; s = DpStats(); }
extern "C" long long dp_host_ext_stats(int* out, long long cap) { auto& v = dp_ext_stats(); long long n = (long long)v.size() < cap ? (long long)v.size() : cap; memcpy(out, v.data(), (size_t)n * sizeof(DpExtStat)); v.clear(); return n; }
#endif
