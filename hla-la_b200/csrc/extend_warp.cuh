// Warp-cooperative version of the extension DP (extend_dp.h holds the scalar statement of the same algorithm and the
// description of what is replayed from extensionAligner::fullNeedleman_diagonal_extension_gapJumper).
//
// One warp runs one extension. The two live wavefronts (m-1, m-2 lists: cell scores plus the adjacency offsets of the cell's
// node, fetched when the cell is created) and the table of cells touched by the current diagonal sit in shared memory; finished
// cells (scores + the three backtrace steps) go to a per-warp slice of an HBM scratch buffer that is only read again for
// revisits, for the "previous score" rule on ties and for the final backtrace.
//   * every candidate gets a sequence number equal to its position in the reference's push order
//     (list, cell index in the list, edge/jump index, open-before-extend), and a candidate is the packed integer
//     (score + 2048) << 19 | (0x7FFFF - seq): "push_back, then first maximum" is one shared-memory atomicMax per candidate,
//     and the winner's backtrace step is decoded from its sequence number;
//   * cells touched by a diagonal are ranked in (x, y, z) order (the reference iterates a std::map), after which the running
//     maximum / patience bookkeeping, which the reference does cell by cell, reduces to a prefix-max scan over that order;
//   * a cell can be touched on two different diagonals only if a gap-path jump created it ahead of its geometric diagonal,
//     so the (x, y, z) -> cell hash in HBM is built lazily when the first jump candidate appears.
// Extensions that overflow the shared-memory capacities (wide gene blocks) return DP_DEFER and are re-run by the next tier.
#pragma once
#include "extend_dp.h"
#include "device_types.h"
#include <cuda_runtime.h>

namespace hlala {

// Capacities are template parameters: the small configuration serves ordinary extensions, the large one the wide ones inside
// gene blocks (~14 nodes per level keep hundreds of cells alive per diagonal).
template <int LIST_, int TD_, int TDHASH_, int WARPS_> struct WdCfg { static constexpr int LIST = LIST_, TD = TD_, TDHASH = TDHASH_, WARPS = WARPS_; };
typedef WdCfg<48, 64, 128, 4> WdTiny;        // ~8 KB per warp: most extensions, high occupancy
typedef WdCfg<128, 128, 256, 4> WdSmall;
typedef WdCfg<512, 1024, 2048, 2> WdLarge;
constexpr int WD_SUB = 64;         // edges + jumps per node
#ifndef DP_DEFER_DEFINED
#define DP_DEFER_DEFINED
constexpr int DP_DEFER = -100;
#endif

struct WdEntry { int32_t cell; int32_t node; int32_t x; int16_t y; int16_t z; int16_t D, GG, SG; uint16_t deg; int32_t k0; int32_t j0; uint16_t jdeg; uint16_t pad; };

// The table of cells touched by one diagonal is an open-addressing table whose slots ARE the entries: tkey holds the packed
// (x - start_level + 2048, y, z) + 1 (0 = empty), tnode the flat node, kD/kGG/kSG the best candidate keys of the three matrices.
struct WdSlab { WdEntry* l0; WdEntry* l1; WdEntry* l2; uint32_t* tkey; int32_t* tnode; uint32_t* kD; uint32_t* kGG; uint32_t* kSG; uint16_t* slots; uint16_t* order; };
template <class CFG> __host__ __device__ inline size_t wd_slab_bytes() { return sizeof(WdEntry) * CFG::LIST * 3 + 20 * CFG::TDHASH + 2 * CFG::TD * 2 + 32; }
template <class CFG> __device__ inline WdSlab wd_carve(unsigned char* p) {
    WdSlab s; s.l0 = (WdEntry*)p; p += sizeof(WdEntry) * CFG::LIST; s.l1 = (WdEntry*)p; p += sizeof(WdEntry) * CFG::LIST; s.l2 = (WdEntry*)p; p += sizeof(WdEntry) * CFG::LIST;
    s.tkey = (uint32_t*)p; p += 4 * CFG::TDHASH; s.tnode = (int32_t*)p; p += 4 * CFG::TDHASH; s.kD = (uint32_t*)p; p += 4 * CFG::TDHASH; s.kGG = (uint32_t*)p; p += 4 * CFG::TDHASH; s.kSG = (uint32_t*)p; p += 4 * CFG::TDHASH;
    s.slots = (uint16_t*)p; p += 2 * CFG::TD; s.order = (uint16_t*)p; return s;
}
__device__ __forceinline__ uint32_t wd_pack(int xrel, int y, int z) { return (((uint32_t)xrel << 19) | ((uint32_t)y << 8) | (uint32_t)z) + 1u; }

__device__ __forceinline__ uint32_t wd_key(int score, uint32_t seq) { int f = score <= DP_NEG ? 0 : score + 2048; return ((uint32_t)f << 19) | (0x7FFFFu - seq); }
__device__ __forceinline__ int wd_score(uint32_t key) { int f = (int)(key >> 19); return f == 0 ? DP_NEG : f - 2048; }
__device__ __forceinline__ uint32_t wd_seq(uint32_t key) { return 0x7FFFFu - (key & 0x7FFFFu); }
__device__ __forceinline__ uint32_t wd_mkseq(int list, int i, int sub, int type) { return (uint32_t)((((list * 2048 + i) * WD_SUB + sub) << 1) | type); }
__device__ __forceinline__ I4 ld4(const void* base, int idx) { const int4 v = __ldg(reinterpret_cast<const int4*>(base) + idx); I4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }

// find-or-insert the touched cell (x, y, z) in the per-diagonal table; returns its slot or -1 (table full / x out of the packed range)
template <class CFG> __device__ __noinline__ int wd_touch(const WdSlab& S, int start_level, int x, int y, int z, int node) {
    const int xrel = x - start_level + 2048;
    if (xrel < 0 || xrel > 4095) return -1;
    const uint32_t key = wd_pack(xrel, y, z);
    uint32_t h = dp_hash3(x, y, z) & (CFG::TDHASH - 1);
    for (int probe = 0; probe < CFG::TDHASH / 2; probe++) {
        const uint32_t old = atomicCAS(&S.tkey[h], 0u, key);
        if (old == 0u) { S.tnode[h] = node; return (int)h; }
        if (old == key) return (int)h;
        h = (h + 1) & (CFG::TDHASH - 1);
    }
    return -1;
}

struct WdCtx {
    const DpGraph* G; const uint8_t* seq; int seq_len; int start_seq, start_level, start_z, start_node; bool pos;
    DpCell* cells; uint32_t* hash; uint32_t* gens;    // HBM slice of this warp
};

__device__ __noinline__ int wd_find_cell(const WdCtx& C, uint32_t cgen, int x, int y, int z) {
    uint32_t h = dp_hash3(x, y, z) & (DP_HASH_CAP - 1);
    for (;;) {
        uint32_t v = C.hash[h];
        if ((v >> 16) != cgen) return -1;
        int idx = (int)(v & 65535u); const DpCell& c = C.cells[idx];
        if (c.x == x && c.y == y && c.z == z) return idx;
        h = (h + 1) & (DP_HASH_CAP - 1);
    }
}
__device__ __noinline__ void wd_insert_cell(const WdCtx& C, uint32_t cgen, int x, int y, int z, int idx) {
    uint32_t h = dp_hash3(x, y, z) & (DP_HASH_CAP - 1);
    for (;;) {
        uint32_t v = C.hash[h];
        if ((v >> 16) != cgen) { if (atomicCAS(&C.hash[h], v, (cgen << 16) | (uint32_t)idx) == v) return; continue; }
        h = (h + 1) & (DP_HASH_CAP - 1);
    }
}

// adjacency of a node in the direction of the extension: edge list [k0, k0+deg), jump list [j0, j0+jdeg)
__device__ __forceinline__ void wd_adj(const WdCtx& C, int node, int32_t& k0, int& deg, int32_t& j0, int& jdeg) {
    const I4 a = ld4(C.G->adj4, node), b = ld4(C.G->adj4, node + 1);
    if (C.pos) { k0 = a.x; deg = b.x - a.x; j0 = a.z; jdeg = b.z - a.z; } else { k0 = a.y; deg = b.y - a.y; j0 = a.w; jdeg = b.w - a.w; }
}

// decode the backtrace step of a winning candidate
__device__ __noinline__ DpBT wd_decode(const WdCtx& C, const WdEntry* m1, const WdEntry* m2, uint32_t key, int matrix /*0 D,1 GG,2 SG*/) {
    const uint32_t seq = wd_seq(key); const int type = seq & 1; const int sub = (seq >> 1) & (WD_SUB - 1); const int li = seq >> 7; const int list = li / 2048, i = li % 2048;
    const WdEntry& src = list == 0 ? m2[i] : m1[i];
    if (matrix == 1) return dp_bt(src.cell, -1, type == 0 ? 0 : 1);
    const void* el = C.pos ? C.G->out_adj4 : C.G->in_adj4;
    if (matrix == 2) return dp_bt(src.cell, ld4(el, src.k0 + sub).x, type == 0 ? 0 : 2);
    if (sub < (int)src.deg) return dp_bt(src.cell, ld4(el, src.k0 + sub).x, 0);
    const void* jl = C.pos ? C.G->jf4 : C.G->jb4;
    return dp_bt(src.cell, -2 - ld4(jl, src.j0 + (sub - (int)src.deg)).x, 0);
}

// One extension by one warp. Returns 0, DP_DEFER (capacity of the shared-memory structures) or a negative error.
template <class CFG> __device__ int wd_extend(const WdCtx& C, const WdSlab& S, int32_t* out_edge, uint8_t* out_s, DpResult& res, int lane) {
    const DpGraph& G = *C.G;
    res.n_cols = 0; res.n_lvl = 0; res.far_y = C.start_seq;
    const int max_level = G.n_levels - 1, max_seq = C.seq_len;
    const int dir = C.pos ? 1 : -1; const int end_seq = C.pos ? max_seq : 0;
    const void* el = C.pos ? G.out_adj4 : G.in_adj4; const void* jl = C.pos ? G.jf4 : G.jb4;
    uint32_t cgen = 0; bool hashed = false;
    int n_cells = 1;
    if (lane == 0) {
        DpCell& c = C.cells[0]; c.x = C.start_level; c.y = (int16_t)C.start_seq; c.z = (int16_t)C.start_z; c.D = 0; c.GG = c.SG = (int16_t)DP_NEG; c.pad = 0;
        c.bD = c.bGG = c.bSG = dp_bt(-1, -1, -1);
        WdEntry& e = S.l0[0]; e.cell = 0; e.node = C.start_node; e.x = C.start_level; e.y = (int16_t)C.start_seq; e.z = (int16_t)C.start_z; e.D = 0; e.GG = e.SG = (int16_t)DP_NEG;
        int deg, jdeg; wd_adj(C, C.start_node, e.k0, deg, e.j0, jdeg); e.deg = (uint16_t)deg; e.jdeg = (uint16_t)jdeg;
        e.pad = (uint16_t)((G.node_gapflags[C.start_node] >> (C.pos ? 0 : 1)) & 1);
    }
    for (int i = lane; i < CFG::TDHASH; i += 32) { S.tkey[i] = 0; S.kD[i] = 0; S.kGG[i] = 0; S.kSG[i] = 0; }
    __syncwarp();
    WdEntry* m1 = S.l0; WdEntry* m2 = S.l1; WdEntry* mt = S.l2; int n_m1 = 1, n_m2 = 0;
    int cur_max = 0, first_max_cell = 0, last_inc = 0;
    int best_end = DP_NEG;      // best stored D of a sequence-complete cell so far (exit by bound, below)
    int status = 0;
    for (int diag = 1; ; diag++) {
        if (diag - last_inc > 40) break;
        if (n_m1 == 0 && n_m2 == 0) break;
        {   // Exact early exit. Once every live cell has consumed the whole read (y == end) only sequence-gap moves remain; if none of
            // the m-1 cells can take a '_' edge or a gap-path jump (score +0), every cell still to come scores at least 6 below its
            // source, which is itself sequence-complete: the end cell, its score and its stored backtrace can no longer change.
            bool live = false;
            for (int i = lane; i < n_m1; i += 32) { const WdEntry& e = m1[i]; live |= (e.y != end_seq) || e.pad || e.jdeg; }
            for (int i = lane; i < n_m2; i += 32) live |= (m2[i].y != end_seq);
            if (!__any_sync(0xffffffffu, live)) break;
        }
        bool ovf = false, saw_jump = false;
        // ---- candidates from the m-2 list (lane per cell)
        for (int i = lane; i < n_m2; i += 32) {
            const WdEntry pc = m2[i];
            const int nx = pc.x + dir, ny = pc.y + dir;
            if (nx > max_level || ny > max_seq || nx < 0 || ny < 0) continue;
            const uint8_t sc = C.pos ? C.seq[pc.y] : C.seq[pc.y - 1];
            if (pc.deg > WD_SUB) { ovf = true; continue; }
            for (int k = 0; k < (int)pc.deg; k++) {
                const I4 a = ld4(el, pc.k0 + k); const uint32_t pk = (uint32_t)a.z;
                const int nz = C.pos ? (int)((pk >> 8) & 255u) : (int)(pk & 255u);
                const int ti = wd_touch<CFG>(S, C.start_level, nx, ny, nz, a.y); if (ti < 0) { ovf = true; break; }
                atomicMax(&S.kD[ti], wd_key(pc.D + ((uint8_t)(pk >> 16) == sc ? 2 : -5), wd_mkseq(0, i, k, 0)));
            }
        }
        // ---- candidates from the m-1 list
        for (int i = lane; i < n_m1; i += 32) {
            const WdEntry pc = m1[i];
            { const int gy = pc.y + dir; const bool ok = C.pos ? (pc.x <= max_level && gy <= max_seq) : (pc.x >= 0 && gy >= 0);
              if (ok) { const int ti = wd_touch<CFG>(S, C.start_level, pc.x, gy, pc.z, pc.node); if (ti < 0) ovf = true; else {
                  atomicMax(&S.kGG[ti], wd_key(pc.D - 6, wd_mkseq(1, i, 0, 0)));
                  atomicMax(&S.kGG[ti], wd_key(pc.GG <= DP_NEG ? DP_NEG : pc.GG - 2, wd_mkseq(1, i, 0, 1))); } } }
            if ((int)pc.deg + (int)pc.jdeg > WD_SUB) { ovf = true; continue; }
            { const int sx = pc.x + dir; const bool ok = C.pos ? (sx <= max_level && pc.y <= max_seq) : (sx >= 0 && pc.y >= 0);
              if (ok) for (int k = 0; k < (int)pc.deg; k++) {
                  const I4 a = ld4(el, pc.k0 + k); const uint32_t pk = (uint32_t)a.z;
                  const int nz = C.pos ? (int)((pk >> 8) & 255u) : (int)(pk & 255u); const bool gapEdge = ((uint8_t)(pk >> 16) == '_');
                  const int ti = wd_touch<CFG>(S, C.start_level, sx, pc.y, nz, a.y); if (ti < 0) { ovf = true; break; }
                  atomicMax(&S.kSG[ti], wd_key(gapEdge ? DP_NEG : pc.D - 6, wd_mkseq(1, i, k, 0)));
                  atomicMax(&S.kSG[ti], wd_key(pc.SG <= DP_NEG ? DP_NEG : (gapEdge ? pc.SG : pc.SG - 2), wd_mkseq(1, i, k, 1)));
                  if (gapEdge) atomicMax(&S.kD[ti], wd_key(pc.D, wd_mkseq(1, i, k, 0)));
              } }
            for (int j = 0; j < (int)pc.jdeg; j++) {
                const I4 jp = ld4(jl, pc.j0 + j);                       // {path, target node, target z, length}
                const int jx = pc.x + dir * jp.w;
                const bool ok = C.pos ? (jx <= max_level && pc.y <= max_seq) : (jx >= 0 && pc.y >= 0);
                if (!ok) continue;
                const int ti = wd_touch<CFG>(S, C.start_level, jx, pc.y, jp.z, jp.y); if (ti < 0) { ovf = true; break; }
                atomicMax(&S.kD[ti], wd_key(pc.D, wd_mkseq(1, i, (int)pc.deg + j, 0)));
                if (jp.w > 1) saw_jump = true;   // see extend_group.cuh: one-level jumps stay on the anti-diagonal of the '_' edge
            }
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, ovf)) { status = DP_DEFER; break; }
        if (!hashed && __any_sync(0xffffffffu, saw_jump)) {
            // first jump over more than one level in this extension: from now on cells can be revisited, so build the (x, y, z) -> cell map
            if (lane == 0) { cgen = C.gens[0] + 1; if (cgen >= (1u << 15)) cgen = 0; C.gens[0] = cgen ? cgen : 1; }
            cgen = __shfl_sync(0xffffffffu, cgen, 0);
            if (cgen == 0) { for (int i = lane; i < DP_HASH_CAP; i += 32) C.hash[i] = 0; cgen = 1; __syncwarp(); }
            for (int i = lane; i < n_cells; i += 32) { const DpCell& c = C.cells[i]; wd_insert_cell(C, cgen, c.x, c.y, c.z, i); }
            hashed = true;
            __syncwarp();
        }
        // ---- collect the touched slots, then rank them in (x, y, z) order: the packed key is monotone in (x, y, z)
        int n_td = 0;
        for (int base = 0; base < CFG::TDHASH; base += 32) {
            const bool used = S.tkey[base + lane] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, used);
            if (used) { const int d = n_td + __popc(m & ((1u << lane) - 1)); if (d < CFG::TD) S.slots[d] = (uint16_t)(base + lane); }
            n_td += __popc(m);
        }
        if (n_td > CFG::TD) { status = DP_DEFER; break; }
        __syncwarp();
        for (int i = lane; i < n_td; i += 32) {
            const uint32_t k = S.tkey[S.slots[i]]; int rank = 0;
            for (int j = 0; j < n_td; j++) rank += (S.tkey[S.slots[j]] < k) ? 1 : 0;
            S.order[rank] = S.slots[i];
        }
        __syncwarp();
        // ---- finalise in that order, 32 cells at a time; carries implement the sequential bookkeeping
        int n_mt = 0; int run_max = cur_max; int new_first = -1; bool any_inc = false;
        for (int base = 0; base < n_td; base += 32) {
            const int oi = base + lane; const bool in = oi < n_td;
            int selD = DP_NEG, selGG = DP_NEG, selSG = DP_NEG; uint32_t kD = 0, kGG = 0, kSG = 0; int selfmat = 0; int tx = 0, ty = 0, tz = 0, tn = 0;
            if (in) {
                const int ti = S.order[oi]; const uint32_t tk = S.tkey[ti] - 1u; tx = (int)(tk >> 19) - 2048 + C.start_level; ty = (int)((tk >> 8) & 2047u); tz = (int)(tk & 255u); tn = S.tnode[ti];
                kGG = S.kGG[ti]; kSG = S.kSG[ti]; kD = S.kD[ti];
                if (kGG) selGG = wd_score(kGG);
                if (kSG) selSG = wd_score(kSG);
                const bool haveD = kD != 0; if (haveD) selD = wd_score(kD);
                if (!haveD || selGG > selD) { selD = selGG; selfmat = 1; }      // the two candidates entering D from this cell's own gap matrices
                if (selSG > selD) { selD = selSG; selfmat = 2; }
            }
            const bool keep = in && selD >= -16;
            int ci = (keep && hashed) ? wd_find_cell(C, cgen, tx, ty, tz) : -1;
            const bool isNew = keep && ci < 0;
            const unsigned newmask = __ballot_sync(0xffffffffu, isNew);
            if (isNew) ci = n_cells + __popc(newmask & ((1u << lane) - 1));
            n_cells += __popc(newmask);
            if (n_cells > DP_CELL_CAP) { status = -4; break; }
            // running maximum inside this diagonal: exclusive prefix max over kept cells in order
            int incl = keep ? selD : -1000000;
            for (int d = 1; d < 32; d <<= 1) { int t2 = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = max(incl, t2); }
            int excl = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) excl = -1000000;
            const int running = max(run_max, excl);
            bool overwritten = false; int stD = DP_NEG, stGG = DP_NEG, stSG = DP_NEG; bool tie_counts = false;
            WdEntry ne;
            if (keep) {
                DpCell& c = C.cells[ci];
                if (isNew) { c.x = tx; c.y = (int16_t)ty; c.z = (int16_t)tz; c.pad = 0; c.D = c.GG = c.SG = (int16_t)DP_NEG; c.bD = c.bGG = c.bSG = dp_bt(-1, -1, -1); }
                if (isNew || c.D < selD) { overwritten = !isNew; c.D = (int16_t)selD; c.bD = selfmat ? dp_bt(ci, -1, selfmat) : wd_decode(C, m1, m2, kD, 0); }
                if (isNew || c.GG < selGG) { overwritten = !isNew; c.GG = (int16_t)selGG; c.bGG = kGG ? wd_decode(C, m1, m2, kGG, 1) : dp_bt(-1, -1, -1); }
                if (isNew || c.SG < selSG) { overwritten = !isNew; c.SG = (int16_t)selSG; c.bSG = kSG ? wd_decode(C, m1, m2, kSG, 2) : dp_bt(-1, -1, -1); }
                if (ty == end_seq) c.pad = 1;                               // sequence-complete cell (extensionAligner.cpp:982-998)
                stD = c.D; stGG = c.GG; stSG = c.SG;
                // The reference's wavefront lists hold coordinates and read the score table when they are used. A cell that is revisited
                // (only possible after a gap-path jump) and improved while it still sits in the m-1 list must therefore show its new
                // scores when that list is consumed as the m-2 list of the next diagonal: refresh the snapshot.
                if (overwritten) for (int i = 0; i < n_m1; i++) if (m1[i].cell == ci) { m1[i].D = (int16_t)stD; m1[i].GG = (int16_t)stGG; m1[i].SG = (int16_t)stSG; }
                if (isNew && hashed) wd_insert_cell(C, cgen, tx, ty, tz, ci);
                if (selD == running) {
                    // score before the last real step, through the stored backtrace (extensionAligner.cpp:1007-1041)
                    DpBT step = c.bD;
                    if (step.src == ci) step = (step.mat == 1) ? c.bGG : c.bSG;
                    int prev = DP_NEG;
                    if (step.src >= 0) { const DpCell& pc = C.cells[step.src]; prev = step.mat == 0 ? pc.D : (step.mat == 1 ? pc.GG : pc.SG); }
                    tie_counts = (selD - prev) != 0;
                }
                ne.cell = ci; ne.node = tn; ne.x = tx; ne.y = (int16_t)ty; ne.z = (int16_t)tz; ne.D = (int16_t)stD; ne.GG = (int16_t)stGG; ne.SG = (int16_t)stSG; ne.pad = (uint16_t)((__ldg(G.node_gapflags + tn) >> (C.pos ? 0 : 1)) & 1);
                int deg, jdeg; wd_adj(C, tn, ne.k0, deg, ne.j0, jdeg); ne.deg = (uint16_t)min(deg, 65535); ne.jdeg = (uint16_t)min(jdeg, 65535);
            }
            if (__any_sync(0xffffffffu, keep && (selD > running || tie_counts || overwritten))) any_inc = true;
            best_end = max(best_end, __reduce_max_sync(0xffffffffu, (keep && ty == end_seq) ? stD : DP_NEG));
            // the cell that first reaches the largest value becomes maxima[0]
            const int chunk_max = __shfl_sync(0xffffffffu, incl, 31);
            if (chunk_max > run_max) {
                const unsigned m = __ballot_sync(0xffffffffu, keep && selD == chunk_max);
                new_first = __shfl_sync(0xffffffffu, ci, __ffs(m) - 1);
                run_max = chunk_max;
            }
            // append to the next wavefront in order
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (n_mt + __popc(km) > CFG::LIST) { status = DP_DEFER; break; }
            if (keep) mt[n_mt + __popc(km & ((1u << lane) - 1))] = ne;
            n_mt += __popc(km);
            __syncwarp();
        }
        if (status) break;
        for (int i = lane; i < n_td; i += 32) { const int sl = S.slots[i]; S.tkey[sl] = 0; S.kD[sl] = 0; S.kGG[sl] = 0; S.kSG[sl] = 0; }
        if (run_max > cur_max) { cur_max = run_max; first_max_cell = new_first; }
        if (any_inc) last_inc = diag;
        __syncwarp();
        // ---- keep cells within 15 of the best stored D (stable compaction)
        if (n_mt > 0) {
            int mx = -1000000; for (int i = lane; i < n_mt; i += 32) mx = max(mx, (int)mt[i].D);
            for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            int w = 0;
            for (int base = 0; base < n_mt; base += 32) {
                const int i = base + lane; const bool in = i < n_mt; WdEntry e; if (in) e = mt[i];
                const bool k = in && (mx - e.D <= 15);
                const unsigned m = __ballot_sync(0xffffffffu, k);
                __syncwarp();
                if (k) mt[w + __popc(m & ((1u << lane) - 1))] = e;
                w += __popc(m);
                __syncwarp();
            }
            n_mt = w;
        }
        WdEntry* t2 = m2; m2 = m1; n_m2 = n_m1; m1 = mt; n_m1 = n_mt; mt = t2;
        __syncwarp();
        // Exact exit by bound (extend_lean.h has the argument): while no jump over >= 2 levels has been taken no stored cell can be touched again,
        // and a cell still to come scores at most its live ancestor's D + 2 per read base left; below the best sequence-complete D nothing changes.
        if (!hashed && best_end > DP_NEG) {
            int bound = -1000000;
            for (int i = lane; i < n_m1; i += 32) { const int rem = C.pos ? max_seq - m1[i].y : m1[i].y; bound = max(bound, (int)m1[i].D + 2 * rem); }
            for (int i = lane; i < n_m2; i += 32) { const int rem = C.pos ? max_seq - m2[i].y : m2[i].y; bound = max(bound, (int)m2[i].D + 2 * rem); }
            if (__reduce_max_sync(0xffffffffu, bound) < best_end) break;
        }
    }
    if (status) return status;
    // ---- end cell: best sequence-complete cell (ties: lexicographically first "level/state" key), else the first cell of the maximum
    int end_cell = -1;
    {
        int bs = DP_NEG - 1, bx = -1, bz = -1, bc = -1;
        for (int i = lane; i < n_cells; i += 32) {
            const DpCell& c = C.cells[i];
            if (c.pad == 1 && c.y == end_seq) { int s = c.D; if (bc < 0 || s > bs || (s == bs && dp_key_less(c.x, c.z, bx, bz))) { bs = s; bx = c.x; bz = c.z; bc = i; } }
        }
        for (int d = 16; d; d >>= 1) {
            int os = __shfl_xor_sync(0xffffffffu, bs, d), ox = __shfl_xor_sync(0xffffffffu, bx, d), oz = __shfl_xor_sync(0xffffffffu, bz, d), oc = __shfl_xor_sync(0xffffffffu, bc, d);
            if (oc >= 0 && (bc < 0 || os > bs || (os == bs && dp_key_less(ox, oz, bx, bz)))) { bs = os; bx = ox; bz = oz; bc = oc; }
        }
        if (bc >= 0) end_cell = bc; else if (cur_max > 0) end_cell = first_max_cell;
    }
    if (end_cell < 0) return 0;
    // ---- backtrace by one lane (extensionAligner.cpp:1109-1354)
    int rc = 0;
    if (lane == 0) {
        int n = 0, n_lvl = 0; int cur = end_cell, mat = 0, guard = 0;
        res.far_y = C.cells[end_cell].y;
        auto emit = [&](int32_t e, uint8_t s) -> bool { if (n >= DP_EXT_CAP) return false; out_edge[n] = e; out_s[n] = s; n++; if (e >= 0) n_lvl++; return true; };
        while (rc == 0 && !(C.cells[cur].x == C.start_level && C.cells[cur].y == C.start_seq)) {
            if (++guard > 8 * DP_CELL_CAP) { rc = -5; break; }
            const DpCell& c = C.cells[cur];
            const DpBT step = mat == 0 ? c.bD : (mat == 1 ? c.bGG : c.bSG);
            if (step.src < 0) { rc = -5; break; }
            const DpCell& sc = C.cells[step.src];
            bool ok = true;
            if (step.edge > -2) {
                if (C.pos) {
                    if (sc.x == c.x - 1 && sc.y == c.y - 1) ok = emit(step.edge, C.seq[c.y - 1]);
                    else if (sc.x == c.x && sc.y == c.y - 1) ok = emit(-1, C.seq[c.y - 1]);
                    else if (sc.x == c.x - 1 && sc.y == c.y) ok = emit(step.edge, '_');
                } else {
                    if (sc.x == c.x + 1 && sc.y == c.y + 1) ok = emit(step.edge, C.seq[c.y]);
                    else if (sc.x == c.x && sc.y == c.y + 1) ok = emit(-1, C.seq[c.y]);
                    else if (sc.x == c.x + 1 && sc.y == c.y) ok = emit(step.edge, '_');
                }
            } else {
                const int p = -2 - step.edge; const int a = G.path_off[p], b = G.path_off[p + 1];
                if (C.pos) { for (int k = b - 1; k >= a && ok; k--) ok = emit(G.path_edges[k], '_'); }
                else { for (int k = a; k < b && ok; k++) ok = emit(G.path_edges[k], '_'); }
            }
            if (!ok) { rc = -4; break; }
            cur = step.src; mat = step.mat;
        }
        if (rc == 0) {
            if (C.pos) for (int i = 0, j = n - 1; i < j; i++, j--) { int32_t te = out_edge[i]; out_edge[i] = out_edge[j]; out_edge[j] = te; uint8_t ts = out_s[i]; out_s[i] = out_s[j]; out_s[j] = ts; }
            res.n_cols = n; res.n_lvl = n_lvl;
        }
    }
    rc = __shfl_sync(0xffffffffu, rc, 0);
    res.n_cols = __shfl_sync(0xffffffffu, res.n_cols, 0); res.n_lvl = __shfl_sync(0xffffffffu, res.n_lvl, 0); res.far_y = __shfl_sync(0xffffffffu, res.far_y, 0);
    return rc;
}

} // namespace hlala
