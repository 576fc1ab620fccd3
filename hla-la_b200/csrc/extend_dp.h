// Soft-clip extension: the reference's affine "diagonal" DP over (level x, read position y, node z) with gap-path
// jumps, replayed as an explicit set-wavefront so that every "first maximum" resolves against the same candidate
// order as extensionAligner::fullNeedleman_diagonal_extension_gapJumper (mapper/aligner/extensionAligner.cpp:335-1556).
//
// What is replayed, in the reference's order:
//   * candidates from the m-2 list (diagonal steps, :565-607), then from the m-1 list (graph gap open/extend :621-661,
//     sequence gap open/extend and the non-affine '_' step :664-753, gap-path jumps :757-786);
//   * per touched cell in (x, y, z) order: first maximum of the GraphGap and SequenceGap candidates, the two
//     "from gap matrix" candidates appended to D, first maximum of D (:794-865);
//   * keep if D >= -16; a cell reached again is overwritten per matrix only if strictly better, which also resets the
//     40-diagonal patience (:949-1062); running maximum bookkeeping incl. the scoreDifference rule (:1007-1057);
//   * next wavefront = touched cells within 15 of the best stored D (:1076-1105);
//   * end cell: sequence-complete cells preferred (max D, ties -> first in the lexicographic order of "level/state"
//     keys — the reference draws among them with rand_r; the oracle build pins that draw to index 0), else the first
//     cell that set the running maximum if it is > 0 (:1391-1499); backtrace (:1109-1354).
// Scores are small integers (the reference holds them in doubles); -DBL_MAX becomes DP_NEG and is absorbing.
//
// One thread runs one extension; its working set (cells, hash, wavefront lists) lives in a per-thread slice of a
// global scratch buffer. The function is __host__ __device__ so that the CPU test-suite can run the identical code
// against the compiled reference without a GPU.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace hlala {

constexpr int DP_NEG = -30000;
// Capacities. Inside gene blocks the graph is ~14 nodes wide and every node of the band stays alive, so one extension can
// visit > 10^4 cells; outside it is a few hundred.
constexpr int DP_CELL_CAP = 32768;
constexpr int DP_HASH_CAP = 65536;      // power of two
constexpr int DP_LIST_CAP = 2048;       // wavefront lists / touched cells per diagonal
constexpr int DP_TDHASH_CAP = 4096;     // power of two
constexpr int DP_ACH_CAP = 1024;
constexpr int DP_EXT_CAP = 512;         // output columns per extension
static_assert(DP_CELL_CAP <= 65536 && DP_LIST_CAP <= 4096, "hash entry layouts");

struct DpBT { int32_t edge; int16_t mat; uint16_t srcu; int32_t src; };      // edge: flat id, -1 none, <= -2: gap path (-2 - id); src: cell index, -1 none, -3 self
struct DpCell { int32_t x; int16_t y; int16_t z; int16_t D, GG, SG; int16_t pad; DpBT bD, bGG, bSG; };
struct DpTouch { int32_t x, y, z; int32_t hasD, hasGG, hasSG; int32_t vD, vGG, vSG; DpBT bD, bGG, bSG; };
__host__ __device__ inline DpBT dp_bt(int src, int edge, int mat) { DpBT b; b.edge = edge; b.mat = (int16_t)mat; b.srcu = 0; b.src = src; return b; }

struct DpScratch {
    DpCell* cells; uint32_t* hash; DpTouch* td; uint32_t* tdhash; int32_t* m1; int32_t* m2; int32_t* mt; int32_t* order; int32_t* ach; int32_t* maxima;
    uint32_t* gens;   // [0] generation of the cell hash, [1] generation of the per-diagonal hash (entries of older generations read as empty)
};
__host__ __device__ inline size_t dp_scratch_bytes() {
    return sizeof(DpCell) * DP_CELL_CAP + 4 * DP_HASH_CAP + sizeof(DpTouch) * DP_LIST_CAP + 4 * DP_TDHASH_CAP + 4 * DP_LIST_CAP * 4 + 4 * DP_ACH_CAP * 2 + 4 * DP_LIST_CAP + 16;
}
__host__ __device__ inline DpScratch dp_carve(unsigned char* p) {
    DpScratch s;
    s.cells = (DpCell*)p; p += sizeof(DpCell) * DP_CELL_CAP;
    s.td = (DpTouch*)p; p += sizeof(DpTouch) * DP_LIST_CAP;
    s.hash = (uint32_t*)p; p += 4 * DP_HASH_CAP;
    s.tdhash = (uint32_t*)p; p += 4 * DP_TDHASH_CAP;
    s.m1 = (int32_t*)p; p += 4 * DP_LIST_CAP; s.m2 = (int32_t*)p; p += 4 * DP_LIST_CAP; s.mt = (int32_t*)p; p += 4 * DP_LIST_CAP; s.order = (int32_t*)p; p += 4 * DP_LIST_CAP;
    s.ach = (int32_t*)p; p += 4 * DP_ACH_CAP * 2;
    s.maxima = (int32_t*)p; p += 4 * DP_LIST_CAP;
    s.gens = (uint32_t*)p; p += 16;
    return s;
}

struct DpGraph {
    int32_t n_levels;
    const int32_t* level_node_off; const uint32_t* edge_pack;
    const int32_t* node_out_off; const int32_t* node_out; const int32_t* node_in_off; const int32_t* node_in;
    const int32_t* path_off; const int32_t* path_edges; const int32_t* path_from; const int32_t* path_to;
    const int32_t* jump_fwd_off; const int32_t* jump_fwd_path; const int32_t* jump_bwd_off; const int32_t* jump_bwd_path;
    const void* adj4; const void* out_adj4; const void* in_adj4; const void* jf4; const void* jb4; const uint8_t* node_gapflags;   // packed adjacency (device only)
};

struct DpResult { int32_t n_cols; int32_t n_lvl; int32_t far_y; };   // far_y: read coordinate of the end cell
#ifdef HLALA_DP_STATS   // host-only instrumentation (tools/dp_stats.py): work per extension
struct DpStats { long long ext, diags, touched, m1, m2, cells, diags_all_at_end, cand, max_td, max_m1; };
inline DpStats& dp_stats() { static DpStats s = {}; return s; }
// one record per extension: clip length, diagonals, cells, largest touched set / wavefront, widest level reached (z + 1), widest spread of
// read positions inside one touched set, first diagonal with a gap-path jump over more than one level (0: none), longest such jump, revisits
struct DpExtStat { int clip, diags, cells, max_td, max_m1, max_w, max_vspread, first_long_jump, max_jump_len, revisits, max_deg; };
#include <vector>
inline std::vector<DpExtStat>& dp_ext_stats() { static std::vector<DpExtStat> v; return v; }
#endif

__host__ __device__ inline uint32_t dp_hash3(int x, int y, int z) { uint32_t h = (uint32_t)x * 0x9E3779B1u ^ (uint32_t)y * 0x85EBCA77u ^ (uint32_t)z * 0xC2B2AE3Du; h ^= h >> 15; return h; }

// compares the decimal strings "x1/z1" and "x2/z2" as std::string operator< would ('/' sorts below every digit)
__host__ __device__ inline bool dp_key_less(int x1, int z1, int x2, int z2) {
    char a[24], b[24]; int na = 0, nb = 0;
    auto put = [](char* o, int& n, int v) { char t[12]; int k = 0; if (v == 0) t[k++] = '0'; while (v > 0) { t[k++] = (char)('0' + v % 10); v /= 10; } while (k) o[n++] = t[--k]; };
    put(a, na, x1); a[na++] = '/'; put(a, na, z1);
    put(b, nb, x2); b[nb++] = '/'; put(b, nb, z2);
    int n = na < nb ? na : nb;
    for (int i = 0; i < n; i++) { if (a[i] != b[i]) return (unsigned char)a[i] < (unsigned char)b[i]; }
    return na < nb;
}

// Returns 0 (n_cols may be 0: "no extension"), or a negative HLALA_E_* code (-4 capacity).
__host__ __device__ inline int dp_extend(const DpGraph& G, const uint8_t* seq, int seq_len, int start_seq, int start_level, int start_z, bool pos,
                                         const DpScratch& S, int32_t* out_edge, uint8_t* out_s, DpResult& res) {
    res.n_cols = 0; res.n_lvl = 0; res.far_y = start_seq;
    const int max_level = G.n_levels - 1, max_seq = seq_len, min_level = 0, min_seq = 0;
    const int dir = pos ? 1 : -1;
    const int end_seq = pos ? max_seq : min_seq;
    // hash entries are (generation << 16) | index; a scratch slice is zero-initialised once and then reused across
    // extensions without clearing (generation 0 never matches)
    uint32_t cgen = S.gens[0] + 1;
    if (cgen >= (1u << 15)) { for (int i = 0; i < DP_HASH_CAP; i++) S.hash[i] = 0; cgen = 1; }
    S.gens[0] = cgen;
    uint32_t tgen = S.gens[1];
    int n_cells = 0;
    auto find_cell = [&](int x, int y, int z) -> int {
        uint32_t h = dp_hash3(x, y, z) & (DP_HASH_CAP - 1);
        while ((S.hash[h] >> 16) == cgen) { int idx = (int)(S.hash[h] & 65535u); const DpCell& c = S.cells[idx]; if (c.x == x && c.y == y && c.z == z) return idx; h = (h + 1) & (DP_HASH_CAP - 1); }
        return -1;
    };
    auto add_cell = [&](int x, int y, int z) -> int {
        if (n_cells >= DP_CELL_CAP) return -1;
        uint32_t h = dp_hash3(x, y, z) & (DP_HASH_CAP - 1);
        while ((S.hash[h] >> 16) == cgen) h = (h + 1) & (DP_HASH_CAP - 1);
        S.hash[h] = (cgen << 16) | (uint32_t)n_cells; DpCell& c = S.cells[n_cells]; c.x = x; c.y = (int16_t)y; c.z = (int16_t)z; c.D = c.GG = c.SG = (int16_t)DP_NEG;
        c.bD = c.bGG = c.bSG = dp_bt(-1, -1, -1);
        return n_cells++;
    };
    const int start_cell = add_cell(start_level, start_seq, start_z);
    S.cells[start_cell].D = 0;
    int n_m1 = 1, n_m2 = 0; S.m1[0] = start_cell;
    int32_t* m1 = S.m1; int32_t* m2 = S.m2; int32_t* mt = S.mt;
    int cur_max = 0; int n_maxima = 1; S.maxima[0] = start_cell;
    int last_inc = 0; int n_ach = 0;
    int n_td = 0;
    auto touch = [&](int x, int y, int z) -> int {
        uint32_t h = dp_hash3(x, y, z) & (DP_TDHASH_CAP - 1);
        while ((S.tdhash[h] >> 12) == tgen) { int idx = (int)(S.tdhash[h] & 4095u); const DpTouch& t = S.td[idx]; if (t.x == x && t.y == y && t.z == z) return idx; h = (h + 1) & (DP_TDHASH_CAP - 1); }
        if (n_td >= DP_LIST_CAP) return -1;
        S.tdhash[h] = (tgen << 12) | (uint32_t)n_td; DpTouch& t = S.td[n_td]; t.x = x; t.y = y; t.z = z; t.hasD = t.hasGG = t.hasSG = 0; t.vD = t.vGG = t.vSG = DP_NEG;
        t.bD = t.bGG = t.bSG = dp_bt(-1, -1, -1);
        return n_td++;
    };
    // "push_back then first maximum" == keep the first candidate, replace only on strictly greater
    auto candD = [&](int ti, int v, DpBT b) { DpTouch& t = S.td[ti]; if (!t.hasD || v > t.vD) { t.vD = v; t.bD = b; } t.hasD = 1; };
    auto candGG = [&](int ti, int v, DpBT b) { DpTouch& t = S.td[ti]; if (!t.hasGG || v > t.vGG) { t.vGG = v; t.bGG = b; } t.hasGG = 1; };
    auto candSG = [&](int ti, int v, DpBT b) { DpTouch& t = S.td[ti]; if (!t.hasSG || v > t.vSG) { t.vSG = v; t.bSG = b; } t.hasSG = 1; };
    auto addneg = [](int a, int b) { return a <= DP_NEG ? DP_NEG : a + b; };

    int status = 0;
#ifdef HLALA_DP_STATS
    DpExtStat xs = {}; xs.clip = pos ? seq_len - start_seq : start_seq;
#endif
    for (int diag = 1; ; diag++) {
        if (diag - last_inc > 40) break;
        if (n_m1 == 0 && n_m2 == 0) break;      // nothing can be produced any more; the reference idles until the patience test fires
#ifdef HLALA_DP_STATS
        { DpStats& st = dp_stats(); st.diags++; st.m1 += n_m1; st.m2 += n_m2; if (n_m1 > st.max_m1) st.max_m1 = n_m1; bool all_end = true; for (int i = 0; i < n_m1; i++) all_end &= S.cells[m1[i]].y == end_seq; for (int i = 0; i < n_m2; i++) all_end &= S.cells[m2[i]].y == end_seq; st.diags_all_at_end += all_end; }
#endif
        tgen++;
        if (tgen >= (1u << 19)) { for (int i = 0; i < DP_TDHASH_CAP; i++) S.tdhash[i] = 0; tgen = 1; }
        n_td = 0;
        // ---- from the m-2 list: diagonal steps
        for (int i = 0; i < n_m2 && !status; i++) {
            const DpCell pc = S.cells[m2[i]];
            int nx = pc.x + dir, ny = pc.y + dir;
            if (nx > max_level || ny > max_seq || nx < min_level || ny < min_seq) continue;
            uint8_t sc = pos ? seq[pc.y] : seq[pc.y - 1];
            int node = G.level_node_off[pc.x] + pc.z;
            int k0 = pos ? G.node_out_off[node] : G.node_in_off[node], k1 = pos ? G.node_out_off[node + 1] : G.node_in_off[node + 1];
            for (int k = k0; k < k1; k++) {
                int e = pos ? G.node_out[k] : G.node_in[k]; uint32_t pk = G.edge_pack[e];
                int nz = pos ? (int)((pk >> 8) & 255u) : (int)(pk & 255u); uint8_t em = (uint8_t)(pk >> 16);
                int ti = touch(nx, ny, nz); if (ti < 0) { status = -4; break; }
                candD(ti, pc.D + (em == sc ? 2 : -5), dp_bt(m2[i], e, 0));
            }
        }
        // ---- from the m-1 list
        for (int i = 0; i < n_m1 && !status; i++) {
            const DpCell pc = S.cells[m1[i]];
            // gap in graph: consume a read base, stay on the node
            { int gx = pc.x, gy = pc.y + dir;
              bool ok = pos ? (gx <= max_level && gy <= max_seq) : (gx >= min_level && gy >= min_seq);
              if (ok) { int ti = touch(gx, gy, pc.z); if (ti < 0) { status = -4; break; }
                        candGG(ti, pc.D - 4 - 2, dp_bt(m1[i], -1, 0)); candGG(ti, addneg(pc.GG, -2), dp_bt(m1[i], -1, 1)); } }
            // gap in sequence: follow an edge without consuming a read base
            { int sx = pc.x + dir, sy = pc.y;
              bool ok = pos ? (sx <= max_level && sy <= max_seq) : (sx >= min_level && sy >= min_seq);
              if (ok) {
                  int node = G.level_node_off[pc.x] + pc.z;
                  int k0 = pos ? G.node_out_off[node] : G.node_in_off[node], k1 = pos ? G.node_out_off[node + 1] : G.node_in_off[node + 1];
                  for (int k = k0; k < k1; k++) {
                      int e = pos ? G.node_out[k] : G.node_in[k]; uint32_t pk = G.edge_pack[e];
                      int nz = pos ? (int)((pk >> 8) & 255u) : (int)(pk & 255u); bool gapEdge = ((uint8_t)(pk >> 16) == '_');
                      int ti = touch(sx, sy, nz); if (ti < 0) { status = -4; break; }
                      candSG(ti, gapEdge ? DP_NEG : pc.D - 4 - 2, dp_bt(m1[i], e, 0));
                      candSG(ti, gapEdge ? addneg(pc.SG, 0) : addneg(pc.SG, -2), dp_bt(m1[i], e, 2));
                      if (gapEdge) candD(ti, pc.D + 0, dp_bt(m1[i], e, 0));
                  }
                  if (status) break;
              } }
            // gap-path jumps: D + len * 0 into (jump level, same y, jump node)
            { int node = G.level_node_off[pc.x] + pc.z;
              int k0 = pos ? G.jump_fwd_off[node] : G.jump_bwd_off[node], k1 = pos ? G.jump_fwd_off[node + 1] : G.jump_bwd_off[node + 1];
              for (int k = k0; k < k1; k++) {
                  int p = pos ? G.jump_fwd_path[k] : G.jump_bwd_path[k];
                  int len = G.path_off[p + 1] - G.path_off[p];
                  int jx = pc.x + dir * len; int jy = pc.y;
                  bool ok = pos ? (jx <= max_level && jy <= max_seq) : (jx >= min_level && jy >= min_seq);
                  if (!ok) continue;
                  int tgt = pos ? G.path_to[p] : G.path_from[p];
                  int jz = tgt - G.level_node_off[jx];
                  int ti = touch(jx, jy, jz); if (ti < 0) { status = -4; break; }
#ifdef HLALA_DP_STATS
                  if (len > 1) { if (!xs.first_long_jump) xs.first_long_jump = diag; if (len > xs.max_jump_len) xs.max_jump_len = len; }
#endif
                  candD(ti, pc.D + 0, dp_bt(m1[i], -2 - p, 0));
              } }
        }
        if (status) break;
#ifdef HLALA_DP_STATS
        dp_stats().touched += n_td; if (n_td > dp_stats().max_td) dp_stats().max_td = n_td;
        { xs.diags = diag; if (n_td > xs.max_td) xs.max_td = n_td; if (n_m1 > xs.max_m1) xs.max_m1 = n_m1; int ylo = 1 << 30, yhi = -1;
          for (int i = 0; i < n_td; i++) { const DpTouch& t = S.td[i]; if (t.z + 1 > xs.max_w) xs.max_w = t.z + 1; if (t.y < ylo) ylo = t.y; if (t.y > yhi) yhi = t.y; }
          if (n_td && yhi - ylo + 1 > xs.max_vspread) xs.max_vspread = yhi - ylo + 1; }
#endif
        // ---- finalise touched cells in (x, y, z) order
        for (int i = 0; i < n_td; i++) {
            int j = i; const DpTouch& t = S.td[i];
            while (j > 0) { const DpTouch& u = S.td[S.order[j - 1]];
                bool less = t.x < u.x || (t.x == u.x && (t.y < u.y || (t.y == u.y && t.z < u.z)));
                if (!less) break; S.order[j] = S.order[j - 1]; j--; }
            S.order[j] = i;
        }
        int n_mt = 0;
        for (int oi = 0; oi < n_td; oi++) {
            DpTouch& t = S.td[S.order[oi]];
            int selGG = t.hasGG ? t.vGG : DP_NEG, selSG = t.hasSG ? t.vSG : DP_NEG;
            int ci = find_cell(t.x, t.y, t.z);
            // the two candidates entering D from the gap matrices of this very cell (source = the cell itself)
            // (the cell index is only known once it exists; -3 marks "self")
            { if (!t.hasD || selGG > t.vD) { t.vD = selGG; t.bD = dp_bt(-3, -1, 1); } t.hasD = 1;
              if (selSG > t.vD) { t.vD = selSG; t.bD = dp_bt(-3, -1, 2); } }
            int selD = t.vD;
            if (selD < -16) continue;
            bool isNew = (ci < 0);
#ifdef HLALA_DP_STATS
            if (!isNew) xs.revisits++;
#endif
            if (isNew) { ci = add_cell(t.x, t.y, t.z); if (ci < 0) { status = -4; break; } }
            DpCell& c = S.cells[ci];
            bool overwritten = false;
            if (isNew || c.D < selD) { overwritten = !isNew; c.D = (int16_t)selD; c.bD = t.bD; if (c.bD.src == -3) c.bD.src = ci; }
            if (isNew || c.GG < selGG) { overwritten = !isNew; c.GG = (int16_t)selGG; c.bGG = t.bGG; }
            if (isNew || c.SG < selSG) { overwritten = !isNew; c.SG = (int16_t)selSG; c.bSG = t.bSG; }
            if (t.y == end_seq) {
                bool have = false; for (int a = 0; a < n_ach; a++) if (S.ach[2 * a] == t.x && S.ach[2 * a + 1] == t.z) { have = true; break; }
                if (!have) { if (n_ach >= DP_ACH_CAP) { status = -4; break; } S.ach[2 * n_ach] = t.x; S.ach[2 * n_ach + 1] = t.z; n_ach++; }
            }
            if (n_mt >= DP_LIST_CAP) { status = -4; break; }
            mt[n_mt++] = ci;
            // score of the last real step into this cell, through the STORED backtrace (extensionAligner.cpp:1007-1041)
            DpBT step = c.bD;
            while (step.src == ci) step = (step.mat == 1) ? S.cells[step.src].bGG : S.cells[step.src].bSG;
            int prev = DP_NEG;
            if (step.src >= 0) { const DpCell& pcell = S.cells[step.src]; prev = step.mat == 0 ? pcell.D : (step.mat == 1 ? pcell.GG : pcell.SG); }
            int diff = selD - prev;
            if (selD == cur_max) { if (diff != 0) { if (n_maxima < DP_LIST_CAP) S.maxima[n_maxima++] = ci; last_inc = diag; } }
            else if (selD > cur_max) { cur_max = selD; n_maxima = 0; S.maxima[n_maxima++] = ci; last_inc = diag; }
            if (overwritten) last_inc = diag;
        }
        if (status) break;
        // ---- keep cells within 15 of the best stored D of this wavefront
        if (n_mt > 0) {
            int mx = S.cells[mt[0]].D; for (int i = 1; i < n_mt; i++) { int d = S.cells[mt[i]].D; if (mx < d) mx = d; }
            int w = 0; for (int i = 0; i < n_mt; i++) if (mx - S.cells[mt[i]].D <= 15) mt[w++] = mt[i];
            n_mt = w;
        }
        int32_t* t2 = m2; m2 = m1; n_m2 = n_m1; m1 = mt; n_m1 = n_mt; mt = t2;
    }
    S.gens[1] = tgen;
#ifdef HLALA_DP_STATS
    dp_stats().ext++; dp_stats().cells += n_cells; xs.cells = n_cells; dp_ext_stats().push_back(xs);
#endif
    if (status) return status;

    // ---- end cell
    int end_cell = -1;
    if (n_ach > 0) {
        int bx = -1, bz = -1, bs = 0;
        for (int a = 0; a < n_ach; a++) {
            int x = S.ach[2 * a], z = S.ach[2 * a + 1]; int ci = find_cell(x, end_seq, z); int s = S.cells[ci].D;
            if (bx < 0 || s > bs || (s == bs && dp_key_less(x, z, bx, bz))) { bx = x; bz = z; bs = s; }
        }
        end_cell = find_cell(bx, end_seq, bz);
    } else if (cur_max > 0) {
        end_cell = S.maxima[0];
    }
    if (end_cell < 0) return 0;
    // ---- backtrace (extensionAligner.cpp:1109-1354); columns are produced from the far end towards the origin
    int n = 0, n_lvl = 0;
    auto emit = [&](int32_t e, uint8_t s) -> bool { if (n >= DP_EXT_CAP) return false; out_edge[n] = e; out_s[n] = s; n++; if (e >= 0) n_lvl++; return true; };
    int cur = end_cell, mat = 0; int guard = 0;
    res.far_y = S.cells[end_cell].y;
    while (!(S.cells[cur].x == start_level && S.cells[cur].y == start_seq)) {
        if (++guard > 8 * DP_CELL_CAP) return -5;
        const DpCell& c = S.cells[cur];
        DpBT step = mat == 0 ? c.bD : (mat == 1 ? c.bGG : c.bSG);
        if (step.src < 0) return -5;
        const DpCell& sc = S.cells[step.src];
        if (step.edge > -2) {
            if (pos) {
                if (sc.x == c.x - 1 && sc.y == c.y - 1) { if (!emit(step.edge, seq[c.y - 1])) return -4; }
                else if (sc.x == c.x && sc.y == c.y - 1) { if (!emit(-1, seq[c.y - 1])) return -4; }
                else if (sc.x == c.x - 1 && sc.y == c.y) { if (!emit(step.edge, '_')) return -4; }
            } else {
                if (sc.x == c.x + 1 && sc.y == c.y + 1) { if (!emit(step.edge, seq[c.y])) return -4; }
                else if (sc.x == c.x && sc.y == c.y + 1) { if (!emit(-1, seq[c.y])) return -4; }
                else if (sc.x == c.x + 1 && sc.y == c.y) { if (!emit(step.edge, '_')) return -4; }
            }
        } else {
            int p = -2 - step.edge; int a = G.path_off[p], b = G.path_off[p + 1];
            if (pos) { for (int k = b - 1; k >= a; k--) if (!emit(G.path_edges[k], '_')) return -4; }
            else { for (int k = a; k < b; k++) if (!emit(G.path_edges[k], '_')) return -4; }
        }
        cur = step.src; mat = step.mat;
    }
    if (pos) { for (int i = 0, j = n - 1; i < j; i++, j--) { int32_t te = out_edge[i]; out_edge[i] = out_edge[j]; out_edge[j] = te; uint8_t ts = out_s[i]; out_s[i] = out_s[j]; out_s[j] = ts; } }
    res.n_cols = n; res.n_lvl = n_lvl;
    return 0;
}

} // namespace hlala
