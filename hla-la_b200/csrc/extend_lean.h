// First tier of the extension DP: ONE THREAD per extension, every piece of per-diagonal state in shared memory.
//
// Same algorithm, same candidate order and same tie rules as extend_dp.h (which documents what is replayed from
// extensionAligner::fullNeedleman_diagonal_extension_gapJumper, extensionAligner.cpp:335-1556); what differs is the data layout:
//   * a cell is the 21-bit key K = (x - xbase) << 12 | (y - ybase) << 4 | z whose integer order is the reference's (x, y, z) map order in
//     both directions (bases chosen so that every reachable offset is non-negative), its three scores are one 32-bit word (10 bits each,
//     0 = -infinity), and a wavefront entry is two words (key | cell index, scores);
//   * the cells touched by a diagonal live in a 32-slot open-addressing table in shared memory whose hash is perfect for the common case
//     (one anti-diagonal, <= 8 read positions, <= 4 nodes per level) and ordered like the keys, so "find or insert" is one probe and the
//     (x, y, z) order of finalisation comes out of a scan of the slots (an insertion sort over the slot list fixes what is left);
//   * "push_back, then first maximum" is sequential here: a candidate replaces the kept one only if strictly greater;
//   * finished cells go to HBM as one 16-byte record (key, scores, the three backtrace steps as cell index + step kind + edge rank inside
//     its level) written once; they are read again only by the backtrace, by revisits and by the previous-score rule on revisits;
//   * a cell can be touched on two diagonals only if a gap-path jump over >= 2 levels created it AHEAD of its anti-diagonal
//     (lead = anti-diagonal - diagonal > 0). Only such cells are remembered, in a per-thread ring in HBM indexed by their anti-diagonal (mod 256, at most
//     16 per anti-diagonal): what a diagonal has to look up is one contiguous 128-byte line, and a 256-bit mask of the anti-diagonals that hold ahead cells
//     keeps every other touched cell away from it. A slot of the ring is recycled when the diagonal has passed its anti-diagonal.
//   * jumps over ONE level are not replayed: their candidate equals the '_' edge's D candidate of the same source (pushed earlier, same
//     target, same score), so it never is the first maximum and never creates a cell of its own.
// Anything beyond the small capacities (wide gene blocks, long clips, very long jumps) returns DP_DEFER and is re-run by the next tiers.
//
// The code is __host__ __device__ and templated on the accessor of the per-thread words, so that the CPU test-suite runs the very same
// statements against the scalar DP and the compiled reference (tests/native/dp_host.cpp).
#pragma once
#include "extend_dp.h"

namespace hlala {

#ifndef DP_DEFER_DEFINED
#define DP_DEFER_DEFINED
constexpr int DP_DEFER = -100;
#endif

#ifdef HLALA_LN_REASONS   // host-only instrumentation (tools/dp_stats.py): why the tier deferred
inline long long* ln_reasons() { static long long r[16] = {}; return r; }
#define LN_WHY(i) (ln_reasons()[i]++)
#else
#define LN_WHY(i) ((void)0)
#endif

template <int LIST_, int TD_> struct LnCfg {
    static constexpr int LIST = LIST_, TD = TD_;                       // TD: slots of the touch table (power of two), at most TD - 1 in use (one empty slot ends the probe of a key that is not in the table)
    static constexpr int M_A = 0, M_B = 2 * LIST_, TK = 4 * LIST_, TV = TK + TD_, TB = TV + TD_, PERM = TB + TD_;
    static constexpr int AM = PERM + TD_ / 4;                          // 256-bit mask of the anti-diagonals (mod 256) that hold ahead cells
    static constexpr int WORDS = AM + 8;
};
typedef LnCfg<26, 32> LnStd;      // 864 B per thread
typedef LnCfg<48, 64> LnBig;      // 1600 B per thread: re-runs what LnStd deferred for its list / table capacities
static_assert(LnBig::LIST <= 64, "list positions are 6-bit fields");

constexpr int LN_CELLS = 2047;          // 11-bit cell index
constexpr int LN_RING = 256, LN_RSLOT = 16;     // ahead cells: a ring over the anti-diagonals (mod 256) of LN_RSLOT entries each
constexpr int LN_AHEAD = LN_RING * LN_RSLOT;    // entries per thread; the per-diagonal counts (LN_RING bytes) follow them
constexpr int LN_MAXLEAD = 250;
constexpr int LN_UMAX = 510, LN_VMAX = 255, LN_ZMAX = 15, LN_EMAX = 31;
constexpr uint32_t LN_KMASK = 0x1FFFFFu, LN_EMPTY = 0xFFFFFFFFu;
constexpr int HLALA_DP_E_CAPACITY = -4, HLALA_DP_E_INVARIANT = -5;
enum { LN_DIAG = 0, LN_GAPEDGE = 1, LN_JUMP = 2, LN_SELF_GG = 3, LN_SELF_SG = 4 };

struct LnRec { uint32_t k, v, b0, b1; };
// b0: D source cell (11) | D kind (3) << 11 | D edge rank / jump rank (5) << 14 | GG source (11) << 19 | GG from GG (1) << 30
// b1: SG source (11) | SG from SG (1) << 11 | SG edge rank (5) << 12 | SG edge is '_' (1) << 17

// one 16-byte record per level: first edge (26 bits) | edge count (6 bits; 63 = more than the tier takes) and the dp_pack words of its first three edges, so
// that a cell's edges cost one load (94 % of the levels of a PRG have <= 3 edges); records are prefetched when a cell enters the next wavefront
struct LnLvl { uint32_t ec, pk0, pk1, pk2; };
#if defined(__CUDA_ARCH__)
#define LN_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p))
#else
#define LN_PREFETCH(p) ((void)0)
#endif
struct LnAhead { uint32_t k, v; };   // ahead ring entry: key | cell index << 21, stored scores

struct LnGraph {   // what the tier reads of the graph; dp_pack = edge_pack | (from node has a forward jump over >= 2 levels) << 24 | (to node has a backward one) << 25
    int32_t n_levels; const int32_t* level_node_off; const int32_t* level_edge_off; const uint32_t* dp_pack; const LnLvl* lvl4;
    const int32_t* path_off; const int32_t* path_edges; const int32_t* path_from; const int32_t* path_to;
    const int32_t* jump_fwd_off; const int32_t* jump_fwd_path; const int32_t* jump_bwd_off; const int32_t* jump_bwd_path;
};

struct LnState {
    const uint8_t* seq; int max_seq, max_level, dir, xbase, ybase, kv_end, start_seq;
    int n_m1, n_m2, rot, diag, last_inc, run_max, first_max_cell, n_cells, n_ahead;
    int end_idx, end_f, end_x, end_z;
    int ahead_hi; int ahead_ready;      // ahead_hi: largest anti-diagonal that holds an ahead cell (none pending once the diagonal has passed it)
};

__host__ __device__ inline int ln_dec(uint32_t f) { return f == 0 ? DP_NEG : (int)f - 32; }

// "x1/z1" < "x2/z2" as std::string operator< on the decimal strings ('/' sorts below every digit); equals dp_key_less (checked in the tests)
__host__ __device__ inline bool ln_key_less(int x1, int z1, int x2, int z2) {
    auto digits = [](int v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; };
    auto p10 = [](int n) { int r = 1; while (n-- > 0) r *= 10; return r; };
    if (x1 != x2) {
        const int d1 = digits(x1), d2 = digits(x2);
        if (d1 == d2) return x1 < x2;
        if (d1 < d2) { const int pre = x2 / p10(d2 - d1); return x1 != pre ? x1 < pre : true; }    // equal prefix: '/' against a digit
        const int pre = x1 / p10(d1 - d2); return pre != x2 ? pre < x2 : false;
    }
    if (z1 == z2) return false;
    const int d1 = digits(z1), d2 = digits(z2);
    if (d1 == d2) return z1 < z2;
    if (d1 < d2) { const int pre = z2 / p10(d2 - d1); return z1 != pre ? z1 < pre : true; }           // equal prefix: the shorter string is smaller
    const int pre = z1 / p10(d1 - d2); return pre != z2 ? pre < z2 : false;
}

template <class CFG, class SM> struct LnDp {
    // ---- task set-up; returns 0 or DP_DEFER
    __host__ __device__ static int init(const LnGraph& G, SM& S, LnState& st, LnRec* rec, const uint8_t* seq, int seq_len, int start_seq, int start_level, int start_z, bool pos) {
        st.seq = seq; st.max_seq = seq_len; st.max_level = G.n_levels - 1; st.dir = pos ? 1 : -1; st.start_seq = start_seq;
        const int clip = pos ? seq_len - start_seq : start_seq;
        if (clip > LN_VMAX || start_z > LN_ZMAX) { LN_WHY(9); return DP_DEFER; }
        st.xbase = pos ? start_level : start_level - LN_UMAX; st.ybase = pos ? start_seq : start_seq - LN_VMAX;
        st.kv_end = (pos ? seq_len : 0) - st.ybase;
        st.n_m1 = 1; st.n_m2 = 0; st.rot = 0; st.diag = 0; st.last_inc = 0; st.run_max = 32; st.first_max_cell = 0; st.n_cells = 1; st.n_ahead = 0;
        st.end_idx = -1; st.end_f = 0; st.end_x = 0; st.end_z = 0; st.ahead_hi = 0; st.ahead_ready = 0;
        for (int i = 0; i < 8; i++) S(CFG::AM + i) = 0u;
        const uint32_t K = ((uint32_t)(start_level - st.xbase) << 12) | ((uint32_t)(start_seq - st.ybase) << 4) | (uint32_t)start_z;
        S(CFG::M_A) = K; S(CFG::M_A + 1) = 32u;       // cell 0: D = 0, GG = SG = -infinity
        for (int i = 0; i < CFG::TD; i++) S(CFG::TK + i) = LN_EMPTY;
        LnRec r; r.k = K; r.v = 32u; r.b0 = 0; r.b1 = 0; rec[0] = r;
        return 0;
    }

    // ---- one diagonal; returns 0 to continue, 1 when the extension has ended, DP_DEFER on capacity
    __host__ __device__ static int step(const LnGraph& G, SM& S, LnState& st, LnRec* rec, LnAhead* ahead) {
        const int dir = st.dir; const int diag = ++st.diag;
        if (diag - st.last_inc > 40 || (st.n_m1 == 0 && st.n_m2 == 0)) return 1;
        const int m1o = st.rot ? CFG::M_B : CFG::M_A, m2o = st.rot ? CFG::M_A : CFG::M_B;
        int n_td = 0; bool live = false; bool bad = false; int kv_hi = 0, s_lo = 1 << 20;
        auto touch = [&](uint32_t K) -> int {
            const int kv = (int)((K >> 4) & 255u), s = (int)(K >> 12) + kv;
            int slot = (((255 - kv) << 2) + (int)(K & 15u) + s * 11) & (CFG::TD - 1);
            for (int probe = 0; probe < CFG::TD; probe++) {
                const uint32_t k = S(CFG::TK + slot);
                if (k == LN_EMPTY) {
                    if (n_td >= CFG::TD - 1) return -1;
                    S(CFG::TK + slot) = K; S(CFG::TV + slot) = 0; S(CFG::TB + slot) = 0; n_td++;
                    if (kv > kv_hi) kv_hi = kv; if (s < s_lo) s_lo = s;
                    return slot;
                }
                if ((k & LN_KMASK) == K) return slot;
                slot = (slot + 1) & (CFG::TD - 1);
            }
            return -1;
        };
        auto candD = [&](int slot, uint32_t f, uint32_t src, uint32_t kind, uint32_t rank) {
            const uint32_t tv = S(CFG::TV + slot);
            if (f > (tv & 1023u)) { S(CFG::TV + slot) = (tv & ~1023u) | f; S(CFG::TK + slot) = (S(CFG::TK + slot) & LN_KMASK) | (src << 21); S(CFG::TB + slot) = (S(CFG::TB + slot) & ~255u) | kind | (rank << 3); }
        };
        auto candGG = [&](int slot, uint32_t f, uint32_t pos_in_m1, uint32_t from_gg) {
            const uint32_t tv = S(CFG::TV + slot);
            if (f > ((tv >> 10) & 1023u)) { S(CFG::TV + slot) = (tv & ~(1023u << 10)) | (f << 10); S(CFG::TB + slot) = (S(CFG::TB + slot) & ~(127u << 8)) | (pos_in_m1 << 8) | (from_gg << 14); }
        };
        auto candSG = [&](int slot, uint32_t f, uint32_t pos_in_m1, uint32_t from_sg, uint32_t rank, uint32_t gap) {
            const uint32_t tv = S(CFG::TV + slot);
            if (f > (tv >> 20)) { S(CFG::TV + slot) = (tv & ~(1023u << 20)) | (f << 20); S(CFG::TB + slot) = (S(CFG::TB + slot) & ~(8191u << 15)) | (pos_in_m1 << 15) | (from_sg << 21) | (rank << 22) | (gap << 27); }
        };
        // ---- candidates from the m-2 list: diagonal steps (extensionAligner.cpp:565-607)
        for (int i = 0; i < st.n_m2 && !bad; i++) {
            const uint32_t e0 = S(m2o + 2 * i), fD = S(m2o + 2 * i + 1) & 1023u;
            const uint32_t K = e0 & LN_KMASK, idx = e0 >> 21;
            const int ku = (int)(K >> 12), kv = (int)((K >> 4) & 255u), z = (int)(K & 15u);
            live |= (kv != st.kv_end);
            const int x = st.xbase + ku, y = st.ybase + kv, nx = x + dir, ny = y + dir;
            if (nx > st.max_level || ny > st.max_seq || nx < 0 || ny < 0) continue;
            if (ku + dir < 0 || ku + dir > LN_UMAX) { LN_WHY(0); bad = true; break; }
            const uint8_t sc = dir > 0 ? st.seq[y] : st.seq[y - 1];
            const LnLvl lv = G.lvl4[dir > 0 ? x : x - 1]; const int ea = (int)(lv.ec & 0x3FFFFFFu), ne = (int)(lv.ec >> 26);
            if (ne > LN_EMAX + 1) { LN_WHY(1); bad = true; break; }
            for (int k = 0; k < ne; k++) {
                const uint32_t pk = k == 0 ? lv.pk0 : (k == 1 ? lv.pk1 : (k == 2 ? lv.pk2 : G.dp_pack[ea + k])); const int zf = (int)(pk & 255u), zt = (int)((pk >> 8) & 255u);
                if ((dir > 0 ? zf : zt) != z) continue;
                const int nz = dir > 0 ? zt : zf; if (nz > LN_ZMAX) { LN_WHY(2); bad = true; break; }
                const int slot = touch(((uint32_t)(ku + dir) << 12) | ((uint32_t)(kv + dir) << 4) | (uint32_t)nz); if (slot < 0) { LN_WHY(3); bad = true; break; }
                candD(slot, fD + (((pk >> 16) & 255u) == sc ? 2 : -5), idx, LN_DIAG, (uint32_t)k);
            }
        }
        // ---- candidates from the m-1 list (:621-786)
        for (int i = 0; i < st.n_m1 && !bad; i++) {
            const uint32_t e0 = S(m1o + 2 * i), e1 = S(m1o + 2 * i + 1);
            const uint32_t K = e0 & LN_KMASK, idx = e0 >> 21, fD = e1 & 1023u, fGG = (e1 >> 10) & 1023u, fSG = e1 >> 20;
            const int ku = (int)(K >> 12), kv = (int)((K >> 4) & 255u), z = (int)(K & 15u);
            live |= (kv != st.kv_end);
            const int x = st.xbase + ku, y = st.ybase + kv;
            {   // gap in graph: consume a read base, stay on the node
                const int gy = y + dir;
                if (dir > 0 ? gy <= st.max_seq : gy >= 0) {
                    const int slot = touch(((uint32_t)ku << 12) | ((uint32_t)(kv + dir) << 4) | (uint32_t)z); if (slot < 0) { LN_WHY(3); bad = true; break; }
                    candGG(slot, fD - 6, (uint32_t)i, 0); if (fGG) candGG(slot, fGG - 2, (uint32_t)i, 1);
                }
            }
            const int sx = x + dir;
            if (!(dir > 0 ? sx <= st.max_level : sx >= 0)) continue;
            if (ku + dir < 0 || ku + dir > LN_UMAX) { LN_WHY(0); bad = true; break; }
            const LnLvl lv = G.lvl4[dir > 0 ? x : x - 1]; const int ea = (int)(lv.ec & 0x3FFFFFFu), ne = (int)(lv.ec >> 26);
            if (ne > LN_EMAX + 1) { LN_WHY(1); bad = true; break; }
            bool longjump = false;
            for (int k = 0; k < ne; k++) {   // gap in sequence along every edge; the non-affine '_' step
                const uint32_t pk = k == 0 ? lv.pk0 : (k == 1 ? lv.pk1 : (k == 2 ? lv.pk2 : G.dp_pack[ea + k])); const int zf = (int)(pk & 255u), zt = (int)((pk >> 8) & 255u);
                if ((dir > 0 ? zf : zt) != z) continue;
                const int nz = dir > 0 ? zt : zf; if (nz > LN_ZMAX) { LN_WHY(2); bad = true; break; }
                const bool gap = ((pk >> 16) & 255u) == (uint32_t)'_';
                if (pk & (dir > 0 ? (1u << 24) : (1u << 25))) longjump = true;
                const int slot = touch(((uint32_t)(ku + dir) << 12) | ((uint32_t)kv << 4) | (uint32_t)nz); if (slot < 0) { LN_WHY(3); bad = true; break; }
                if (!gap) candSG(slot, fD - 6, (uint32_t)i, 0, (uint32_t)k, 0);
                if (fSG) candSG(slot, gap ? fSG : fSG - 2, (uint32_t)i, 1, (uint32_t)k, gap ? 1u : 0u);
                if (gap) { live = true; candD(slot, fD, idx, LN_GAPEDGE, (uint32_t)k); }
            }
            if (bad) break;
            if (longjump) {   // gap-path jumps over >= 2 levels: D + 0 into (jump level, same y, jump node)
                live = true;
                const int node = G.level_node_off[x] + z;
                const int k0 = dir > 0 ? G.jump_fwd_off[node] : G.jump_bwd_off[node], k1 = dir > 0 ? G.jump_fwd_off[node + 1] : G.jump_bwd_off[node + 1];
                for (int k = k0; k < k1; k++) {
                    const int p = dir > 0 ? G.jump_fwd_path[k] : G.jump_bwd_path[k];
                    const int len = G.path_off[p + 1] - G.path_off[p];
                    if (len < 2) continue;
                    const int jx = x + dir * len;
                    if (!(dir > 0 ? jx <= st.max_level : jx >= 0)) continue;
                    const int nku = ku + dir * len; if (nku < 0 || nku > LN_UMAX || k - k0 > LN_EMAX) { LN_WHY(4); bad = true; break; }
                    const int jz = (dir > 0 ? G.path_to[p] : G.path_from[p]) - G.level_node_off[jx]; if (jz > LN_ZMAX) { LN_WHY(2); bad = true; break; }
                    const int slot = touch(((uint32_t)nku << 12) | ((uint32_t)kv << 4) | (uint32_t)jz); if (slot < 0) { LN_WHY(3); bad = true; break; }
                    candD(slot, fD, idx, LN_JUMP, (uint32_t)(k - k0));
                }
            }
        }
        if (bad) { clear_touch(S); return DP_DEFER; }
        // Exact early exit (extend_warp.cuh): every live cell has consumed the whole read and none can take a '_' edge or a jump
        if (!live) { clear_touch(S); return 1; }
        // ---- touched slots in (x, y, z) order: the slot order of one anti-diagonal is the key order; an insertion sort fixes the rest
        {
            int n = 0;
            int slot = (((255 - kv_hi) << 2) + s_lo * 11) & (CFG::TD - 1);
            for (int c = 0; c < CFG::TD && n < n_td; c++) {
                if (S(CFG::TK + slot) != LN_EMPTY) {
                    const uint32_t K = S(CFG::TK + slot) & LN_KMASK; int j = n;
                    while (j > 0) { const int ps = getb(S, j - 1); if ((S(CFG::TK + ps) & LN_KMASK) <= K) break; setb(S, j, ps); j--; }
                    setb(S, j, slot); n++;
                }
                slot = (slot + 1) & (CFG::TD - 1);
            }
        }
        if (st.ahead_hi >= diag) {   // ahead cells are pending: start the table look-ups of this diagonal's cells now, they are consumed one by one below
            for (int oi = 0; oi < n_td; oi++) {
                const uint32_t K = S(CFG::TK + getb(S, oi)) & LN_KMASK; const int s = (int)(K >> 12) + (int)((K >> 4) & 255u); const int g = dir > 0 ? s : (LN_UMAX + LN_VMAX) - s;
                if (g <= st.ahead_hi && ((S(CFG::AM + ((g >> 5) & 7)) >> (g & 31)) & 1u)) LN_PREFETCH(ahead + (g & (LN_RING - 1)) * LN_RSLOT);
            }
        }
        // A lower bound of the best D this diagonal will store (a cell stores at least its D candidate): cells more than 15 below it are dropped by the
        // pruning at the end of the step (:1076-1105) whatever the exact maximum turns out to be, so they need no place in the next wavefront list.
        uint32_t mx_lo = 0;
        for (int oi = 0; oi < n_td; oi++) { const uint32_t f = S(CFG::TV + getb(S, oi)) & 1023u; if (f > mx_lo) mx_lo = f; }
        // ---- finalise in that order (:794-1071); the next wavefront is written over the m-2 list
        int n_mt = 0; bool dirty = false; int rc = 0;
        for (int oi = 0; oi < n_td; oi++) {
            const int slot = getb(S, oi);
            const uint32_t tk = S(CFG::TK + slot), tv = S(CFG::TV + slot), tb = S(CFG::TB + slot);
            S(CFG::TK + slot) = LN_EMPTY;
            if (rc) continue;
            const uint32_t K = tk & LN_KMASK;
            uint32_t fD = tv & 1023u; const uint32_t fGG = (tv >> 10) & 1023u, fSG = tv >> 20;
            uint32_t kind = tb & 7u;
            if (fD == 0 || fGG > fD) { fD = fGG; kind = LN_SELF_GG; }     // the two candidates entering D from this cell's own gap matrices
            if (fSG > fD) { fD = fSG; kind = LN_SELF_SG; }
            if (fD < 16) continue;                                        // keep only D >= -16
            const int ku = (int)(K >> 12), kv = (int)((K >> 4) & 255u), z = (int)(K & 15u);
            const int s = ku + kv; const int g = dir > 0 ? s : (LN_UMAX + LN_VMAX) - s; const int lead = g - diag;
            int ci = -1; uint32_t ah = 0, stv = 0;     // ahead slot and stored scores of a revisited cell
            if (g <= st.ahead_hi && ((S(CFG::AM + ((g >> 5) & 7)) >> (g & 31)) & 1u)) {
                const int rg = g & (LN_RING - 1); const int cnt = (int)ring_count(ahead)[rg];
                for (int q = 0; q < cnt; q++) { const LnAhead a = ahead[rg * LN_RSLOT + q]; if ((a.k & LN_KMASK) == K) { ci = (int)(a.k >> 21); stv = a.v; ah = (uint32_t)(rg * LN_RSLOT + q); break; } }
            }
            const bool isNew = ci < 0;
            uint32_t b0n = 0, b1n = 0;     // backtrace steps offered by this diagonal
            {
                const uint32_t ggsrc = fGG ? (S(m1o + 2 * ((tb >> 8) & 63u)) >> 21) : 0u, sgsrc = fSG ? (S(m1o + 2 * ((tb >> 15) & 63u)) >> 21) : 0u;
                b0n = (kind >= LN_SELF_GG ? 0u : (tk >> 21)) | (kind << 11) | (((tb >> 3) & 31u) << 14) | (ggsrc << 19) | (((tb >> 14) & 1u) << 30);
                b1n = sgsrc | (((tb >> 21) & 1u) << 11) | (((tb >> 22) & 31u) << 12) | (((tb >> 27) & 1u) << 17);
            }
            LnRec r; bool overwritten = false;
            if (isNew) {
                if (st.n_cells >= LN_CELLS) { LN_WHY(5); rc = DP_DEFER; continue; }
                ci = st.n_cells++;
                stv = fD | (fGG << 10) | (fSG << 20);
                r.k = K; r.v = stv; r.b0 = b0n; r.b1 = b1n; rec[ci] = r;
                if (lead > 0) {
                    if (lead > LN_MAXLEAD) { LN_WHY(6); rc = DP_DEFER; continue; }
                    if (!st.ahead_ready) { LnRec zero; zero.k = zero.v = zero.b0 = zero.b1 = 0; for (int i = 0; i < LN_RING / 16; i++) reinterpret_cast<LnRec*>(ring_count(ahead))[i] = zero; st.ahead_ready = 1; }
                    const int rg = g & (LN_RING - 1); const int cnt = (int)ring_count(ahead)[rg];
                    if (cnt >= LN_RSLOT) { LN_WHY(7); rc = DP_DEFER; continue; }
                    LnAhead a; a.k = K | ((uint32_t)ci << 21); a.v = stv; ahead[rg * LN_RSLOT + cnt] = a; ring_count(ahead)[rg] = (uint8_t)(cnt + 1); st.n_ahead++; S(CFG::AM + ((g >> 5) & 7)) |= 1u << (g & 31); if (g > st.ahead_hi) st.ahead_hi = g;
                }
            } else {     // revisit: the scores come with the table entry; the record is read only if a matrix improves (or for the tie rule below)
                uint32_t sD = stv & 1023u, sGG = (stv >> 10) & 1023u, sSG = stv >> 20;
                overwritten = sD < fD || sGG < fGG || sSG < fSG;
                if (overwritten) {
                    r = rec[ci];
                    if (sD < fD) { sD = fD; r.b0 = (r.b0 & ~0x7FFFFu) | (b0n & 0x7FFFFu); }
                    if (sGG < fGG) { sGG = fGG; r.b0 = (r.b0 & 0x7FFFFu) | (b0n & ~0x7FFFFu); }
                    if (sSG < fSG) { sSG = fSG; r.b1 = b1n; }
                    stv = sD | (sGG << 10) | (sSG << 20);
                    r.v = stv; rec[ci] = r; dirty = true; ahead[ah].v = stv;
                    // the reference's lists hold coordinates and read the score table when used: a cell improved while it still sits in the
                    // m-1 list must show its new scores when that list is consumed as the m-2 list of the next diagonal
                    for (int i = 0; i < st.n_m1; i++) if ((S(m1o + 2 * i) >> 21) == (uint32_t)ci) S(m1o + 2 * i + 1) = stv;
                }
            }
            const uint32_t stD = stv & 1023u;
            if (kv == st.kv_end) {   // sequence-complete: best end cell so far (max stored D, ties -> lexicographically first "level/state" key, :1391-1478)
                const int x = st.xbase + ku;
                if (st.end_idx < 0 || (int)stD > st.end_f || ((int)stD == st.end_f && ci != st.end_idx && ln_key_less(x, z, st.end_x, st.end_z))) { st.end_idx = ci; st.end_f = (int)stD; st.end_x = x; st.end_z = z; }
            }
            if (mx_lo <= 15u + stD) {     // else: pruned below in any case
                if (n_mt >= CFG::LIST) { LN_WHY(8); rc = DP_DEFER; continue; }
                S(m2o + 2 * n_mt) = K | ((uint32_t)ci << 21); S(m2o + 2 * n_mt + 1) = stv; n_mt++;
                { const int pl = st.xbase + ku - (dir > 0 ? 0 : 1); if (pl >= 0 && pl < G.n_levels) LN_PREFETCH(G.lvl4 + pl); }     // the level record this cell reads on the next two diagonals
            }
            // running maximum / patience (:1007-1062); the step score is 0 exactly for '_' steps, jumps and a sequence gap extended along '_'
            if ((int)fD == st.run_max) {
                bool tie;
                if (isNew && !dirty) tie = !(kind == LN_GAPEDGE || kind == LN_JUMP || (kind == LN_SELF_SG && ((tb >> 21) & 1u) && ((tb >> 27) & 1u)));
                else {   // through the STORED backtrace and the scores stored now
                    if (!isNew && !overwritten) r = rec[ci];
                    const uint32_t sk = (r.b0 >> 11) & 7u; uint32_t src, f;
                    if (sk == LN_SELF_GG) { src = (r.b0 >> 19) & 2047u; const uint32_t v = rec[src].v; f = ((r.b0 >> 30) & 1u) ? ((v >> 10) & 1023u) : (v & 1023u); }
                    else if (sk == LN_SELF_SG) { src = r.b1 & 2047u; const uint32_t v = rec[src].v; f = ((r.b1 >> 11) & 1u) ? (v >> 20) : (v & 1023u); }
                    else { src = r.b0 & 2047u; f = rec[src].v & 1023u; }
                    tie = ln_dec(fD) - ln_dec(f) != 0;
                }
                if (tie) st.last_inc = diag;
            } else if ((int)fD > st.run_max) { st.run_max = (int)fD; st.first_max_cell = ci; st.last_inc = diag; }
            if (overwritten) st.last_inc = diag;
        }
        if (rc) return rc;
        // ---- keep cells within 15 of the best stored D (:1076-1105)
        {
            uint32_t mx = 0; for (int i = 0; i < n_mt; i++) { const uint32_t f = S(m2o + 2 * i + 1) & 1023u; if (f > mx) mx = f; }
            int w = 0;
            for (int i = 0; i < n_mt; i++) { const uint32_t a = S(m2o + 2 * i), b = S(m2o + 2 * i + 1); if (mx - (b & 1023u) <= 15u) { S(m2o + 2 * w) = a; S(m2o + 2 * w + 1) = b; w++; } }
            st.n_m2 = st.n_m1; st.n_m1 = w; st.rot ^= 1;
        }
        if (st.ahead_hi >= diag && ((S(CFG::AM + ((diag >> 5) & 7)) >> (diag & 31)) & 1u)) {     // cells of this anti-diagonal can no longer be touched: its ring slot is free again
            S(CFG::AM + ((diag >> 5) & 7)) &= ~(1u << (diag & 31)); ring_count(ahead)[diag & (LN_RING - 1)] = 0;
        }
        // Exact early exit by bound. No stored cell can be touched again (no ahead cell is pending), so the cells and backtrace steps written so
        // far are final; a cell still to come descends from a live entry and scores at most that entry's D + 2 per read base left (every
        // other move adds <= 0, and GG, SG <= D). If that stays strictly below the best sequence-complete D, no later cell can take over or tie
        // the end cell: what the reference computes during its remaining (up to 40) diagonals does not reach the result.
        if (st.end_idx >= 0 && st.ahead_hi <= diag) {
            int best = 0;
            const int a1 = st.rot ? CFG::M_B : CFG::M_A, a2 = st.rot ? CFG::M_A : CFG::M_B;
            for (int i = 0; i < st.n_m1; i++) { const int kv = (int)((S(a1 + 2 * i) >> 4) & 255u); const int rem = kv > st.kv_end ? kv - st.kv_end : st.kv_end - kv; const int v = (int)(S(a1 + 2 * i + 1) & 1023u) + 2 * rem; if (v > best) best = v; }
            for (int i = 0; i < st.n_m2; i++) { const int kv = (int)((S(a2 + 2 * i) >> 4) & 255u); const int rem = kv > st.kv_end ? kv - st.kv_end : st.kv_end - kv; const int v = (int)(S(a2 + 2 * i + 1) & 1023u) + 2 * rem; if (v > best) best = v; }
            if (best < st.end_f) return 1;
        }
        return 0;
    }

    // ---- end cell + backtrace (:1391-1499, :1109-1354); returns 0 or a negative code
    __host__ __device__ static int finish(const LnGraph& G, const LnState& st, const LnRec* rec, int32_t* out_edge, uint8_t* out_s, DpResult& res) {
        res.n_cols = 0; res.n_lvl = 0; res.far_y = st.start_seq;
        const int end_cell = st.end_idx >= 0 ? st.end_idx : (st.run_max > 32 ? st.first_max_cell : -1);
        if (end_cell < 0) return 0;
        const int dir = st.dir;
        int n = 0, n_lvl = 0; int cur = end_cell, mat = 0, guard = 0;
        LnRec c = rec[cur];
        res.far_y = st.ybase + (int)((c.k >> 4) & 255u);
        while (cur != 0) {
            if (++guard > 8 * LN_CELLS) return HLALA_DP_E_INVARIANT;
            const int x = st.xbase + (int)((c.k & LN_KMASK) >> 12), y = st.ybase + (int)((c.k >> 4) & 255u);
            const uint8_t sc = dir > 0 ? st.seq[y > 0 ? y - 1 : 0] : st.seq[y < st.max_seq ? y : st.max_seq - 1];
            const int lvl = dir > 0 ? x - 1 : x;       // level the step's edge leaves from
            int src, nmat; int32_t e = -1; uint8_t s = '_'; bool emit1 = true;
            if (mat == 0) {
                const uint32_t kind = (c.b0 >> 11) & 7u; src = (int)(c.b0 & 2047u); nmat = 0;
                if (kind == LN_SELF_GG) { mat = 1; continue; }
                if (kind == LN_SELF_SG) { mat = 2; continue; }
                if (kind == LN_JUMP) {
                    const LnRec sr = rec[src];
                    const int sx = st.xbase + (int)((sr.k & LN_KMASK) >> 12), sz = (int)(sr.k & 15u);
                    const int node = G.level_node_off[sx] + sz; const int k = (dir > 0 ? G.jump_fwd_off[node] : G.jump_bwd_off[node]) + (int)((c.b0 >> 14) & 31u);
                    const int p = dir > 0 ? G.jump_fwd_path[k] : G.jump_bwd_path[k]; const int a = G.path_off[p], b = G.path_off[p + 1];
                    if (n + (b - a) > DP_EXT_CAP) return HLALA_DP_E_CAPACITY;
                    if (dir > 0) { for (int q = b - 1; q >= a; q--) { out_edge[n] = G.path_edges[q]; out_s[n] = '_'; n++; n_lvl++; } }
                    else { for (int q = a; q < b; q++) { out_edge[n] = G.path_edges[q]; out_s[n] = '_'; n++; n_lvl++; } }
                    emit1 = false;
                } else { e = (int)(G.lvl4[lvl].ec & 0x3FFFFFFu) + (int)((c.b0 >> 14) & 31u); s = kind == LN_DIAG ? sc : (uint8_t)'_'; }
            } else if (mat == 1) { src = (int)((c.b0 >> 19) & 2047u); nmat = (int)((c.b0 >> 30) & 1u) ? 1 : 0; e = -1; s = sc; }
            else { src = (int)(c.b1 & 2047u); nmat = (int)((c.b1 >> 11) & 1u) ? 2 : 0; e = (int)(G.lvl4[lvl].ec & 0x3FFFFFFu) + (int)((c.b1 >> 12) & 31u); s = '_'; }
            if (emit1) { if (n >= DP_EXT_CAP) return HLALA_DP_E_CAPACITY; out_edge[n] = e; out_s[n] = s; n++; if (e >= 0) n_lvl++; }
            cur = src; mat = nmat; c = rec[cur];
        }
        if (dir > 0) for (int i = 0, j = n - 1; i < j; i++, j--) { const int32_t te = out_edge[i]; out_edge[i] = out_edge[j]; out_edge[j] = te; const uint8_t ts = out_s[i]; out_s[i] = out_s[j]; out_s[j] = ts; }
        res.n_cols = n; res.n_lvl = n_lvl;
        return 0;
    }

    __host__ __device__ static uint8_t* ring_count(LnAhead* ahead) { return reinterpret_cast<uint8_t*>(ahead + LN_AHEAD); }
    __host__ __device__ static int getb(SM& S, int i) { return (int)S.b(CFG::PERM + (i >> 2), i & 3); }
    __host__ __device__ static void setb(SM& S, int i, int v) { S.b(CFG::PERM + (i >> 2), i & 3) = (uint8_t)v; }
    __host__ __device__ static void clear_touch(SM& S) { for (int i = 0; i < CFG::TD; i++) S(CFG::TK + i) = LN_EMPTY; }
};

} // namespace hlala
