// Sub-warp version of the extension DP: GW lanes (8) run one extension, so a warp advances 32/GW extensions with one instruction
// stream. Measured on the bench workload, a diagonal of the DP touches ~6 cells (p99 ~12) and an extension lives for ~100
// diagonals: a full warp per extension leaves most lanes idle and pays the per-diagonal bookkeeping once per extension. Here the
// groups of a warp step through "one diagonal each" in lockstep (a flat loop with one warp-wide vote per iteration keeps them
// converged); a group that finishes its extension backtraces, fetches its next task and rejoins the loop.
//
// The algorithm, the candidate keys and every tie rule are those of extend_warp.cuh / extend_dp.h (which document what is replayed
// from extensionAligner::fullNeedleman_diagonal_extension_gapJumper, extensionAligner.cpp:335-1556); only the data-parallel width and
// the capacities differ. Extensions that overflow the small capacities return DP_DEFER and are re-run by the warp-wide tiers.
#pragma once
#include "extend_warp.cuh"

namespace hlala {

template <int GW_, int LIST_, int TD_, int TDHASH_, int CELLS_, int HASH_, int WARPS_> struct GdCfg {
    static constexpr int GW = GW_, LIST = LIST_, TD = TD_, TDHASH = TDHASH_, CELLS = CELLS_, HASH = HASH_, WARPS = WARPS_, NG = 32 / GW_;
};
typedef GdCfg<8, 24, 32, 64, 2048, 4096, 4> GdOct;      // 3.4 KB of shared memory per group: holds 96 % of the bench workload's extensions (LIST 16 holds 80 %)

// wavefront entry of the group kernel: WdEntry packed into 28 bytes (cell index 16 bit, level relative to the start level, z / degrees 8 bit), so
// that a group's slab is 3.4 KB and four CTAs (64 groups) fit an SM's shared memory instead of three
struct GdEntry { int32_t node; int32_t k0; int32_t j0; uint16_t cell; int16_t xr; int16_t y; int16_t D, GG, SG; uint8_t z, deg, jdeg, pad; };
static_assert(sizeof(GdEntry) == 28, "GdEntry layout");

struct GdSlab { GdEntry* l[3]; uint32_t* tkey; int32_t* tnode; uint32_t* kD; uint32_t* kGG; uint32_t* kSG; uint16_t* slots; uint16_t* order; int32_t* cnt; };
template <class CFG> __host__ __device__ inline size_t gd_slab_bytes() { return (sizeof(GdEntry) * CFG::LIST * 3 + 20 * CFG::TDHASH + 2 * CFG::TD * 2 + 16 + 15) & ~size_t(15); }
template <class CFG> __device__ inline GdSlab gd_carve(unsigned char* p) {
    GdSlab s; for (int i = 0; i < 3; i++) { s.l[i] = (GdEntry*)p; p += sizeof(GdEntry) * CFG::LIST; }
    s.tkey = (uint32_t*)p; p += 4 * CFG::TDHASH; s.tnode = (int32_t*)p; p += 4 * CFG::TDHASH; s.kD = (uint32_t*)p; p += 4 * CFG::TDHASH; s.kGG = (uint32_t*)p; p += 4 * CFG::TDHASH; s.kSG = (uint32_t*)p; p += 4 * CFG::TDHASH;
    s.slots = (uint16_t*)p; p += 2 * CFG::TD; s.order = (uint16_t*)p; p += 2 * CFG::TD; s.cnt = (int32_t*)p; return s;
}
template <class CFG> __host__ __device__ inline size_t gd_hbm_bytes() { return (sizeof(DpCell) * (size_t)CFG::CELLS + 4 * (size_t)CFG::HASH + 16 + 15) & ~size_t(15); }

// group-scoped collectives: lanes [shift, shift + GW) of the warp
template <int GW> struct Grp {
    unsigned mask; int shift; int lane;
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & (GW == 32 ? 0xffffffffu : ((1u << (GW & 31)) - 1u)); }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, GW); }
    template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(mask, v, d, GW); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, int d) const { return __shfl_xor_sync(mask, v, d, GW); }
    __device__ __forceinline__ unsigned below() const { return (1u << lane) - 1u; }
};

template <class CFG> __device__ __noinline__ int gd_touch(const GdSlab& S, int start_level, int x, int y, int z, int node) {
    const int xrel = x - start_level + 2048;
    if (xrel < 0 || xrel > 4095) return -1;
    const uint32_t key = wd_pack(xrel, y, z);
    uint32_t h = dp_hash3(x, y, z) & (CFG::TDHASH - 1);
    for (int probe = 0; probe < CFG::TDHASH / 2; probe++) {
        const uint32_t old = atomicCAS(&S.tkey[h], 0u, key);
        if (old == 0u) { S.tnode[h] = node; const int at = atomicAdd(S.cnt, 1); if (at < CFG::TD) S.slots[at] = (uint16_t)h; return (int)h; }   // the list of touched slots grows as cells appear
        if (old == key) return (int)h;
        h = (h + 1) & (CFG::TDHASH - 1);
    }
    return -1;
}
template <class CFG> __device__ __noinline__ int gd_find_cell(const WdCtx& C, uint32_t cgen, int x, int y, int z) {
    uint32_t h = dp_hash3(x, y, z) & (CFG::HASH - 1);
    for (;;) {
        const uint32_t v = C.hash[h];
        if ((v >> 16) != cgen) return -1;
        const int idx = (int)(v & 65535u); const DpCell& c = C.cells[idx];
        if (c.x == x && c.y == y && c.z == z) return idx;
        h = (h + 1) & (CFG::HASH - 1);
    }
}
template <class CFG> __device__ __noinline__ void gd_insert_cell(const WdCtx& C, uint32_t cgen, int x, int y, int z, int idx) {
    uint32_t h = dp_hash3(x, y, z) & (CFG::HASH - 1);
    for (;;) {
        const uint32_t v = C.hash[h];
        if ((v >> 16) != cgen) { if (atomicCAS(&C.hash[h], v, (cgen << 16) | (uint32_t)idx) == v) return; continue; }
        h = (h + 1) & (CFG::HASH - 1);
    }
}

// decode the backtrace step of a winning candidate (wd_decode for the packed entries)
__device__ __noinline__ DpBT gd_decode(const WdCtx& C, const GdEntry* m1, const GdEntry* m2, uint32_t key, int matrix /*0 D,1 GG,2 SG*/) {
    const uint32_t seq = wd_seq(key); const int type = seq & 1; const int sub = (seq >> 1) & (WD_SUB - 1); const int li = seq >> 7; const int list = li / 2048, i = li % 2048;
    const GdEntry& src = list == 0 ? m2[i] : m1[i];
    if (matrix == 1) return dp_bt(src.cell, -1, type == 0 ? 0 : 1);
    const void* el = C.pos ? C.G->out_adj4 : C.G->in_adj4;
    if (matrix == 2) return dp_bt(src.cell, ld4(el, src.k0 + sub).x, type == 0 ? 0 : 2);
    if (sub < (int)src.deg) return dp_bt(src.cell, ld4(el, src.k0 + sub).x, 0);
    const void* jl = C.pos ? C.G->jf4 : C.G->jb4;
    return dp_bt(src.cell, -2 - ld4(jl, src.j0 + (sub - (int)src.deg)).x, 0);
}

// per-group DP state (uniform across the lanes of a group)
struct GdState { int n_m1, n_m2, rot, diag, last_inc, cur_max, first_max_cell, n_cells; uint32_t cgen; bool hashed;
                 int end_c; };   // index of the first sequence-complete cell (-1: none); cells are numbered in creation order

template <class CFG> __device__ __forceinline__ void gd_init(const WdCtx& C, const GdSlab& S, GdState& st, const Grp<CFG::GW>& g) {
    const DpGraph& G = *C.G;
    if (g.lane == 0) {
        DpCell& c = C.cells[0]; c.x = C.start_level; c.y = (int16_t)C.start_seq; c.z = (int16_t)C.start_z; c.D = 0; c.GG = c.SG = (int16_t)DP_NEG; c.pad = 0;
        c.bD = c.bGG = c.bSG = dp_bt(-1, -1, -1);
        GdEntry& e = S.l[0][0]; e.cell = 0; e.node = C.start_node; e.xr = 0; e.y = (int16_t)C.start_seq; e.z = (uint8_t)C.start_z; e.D = 0; e.GG = e.SG = (int16_t)DP_NEG;
        int deg, jdeg; wd_adj(C, C.start_node, e.k0, deg, e.j0, jdeg); e.deg = (uint8_t)min(deg, 255); e.jdeg = (uint8_t)min(jdeg, 255);
        e.pad = (uint8_t)((G.node_gapflags[C.start_node] >> (C.pos ? 0 : 1)) & 1);
        *S.cnt = 0;
    }
    for (int i = g.lane; i < CFG::TDHASH; i += CFG::GW) { S.tkey[i] = 0; S.kD[i] = 0; S.kGG[i] = 0; S.kSG[i] = 0; }
    st.n_m1 = 1; st.n_m2 = 0; st.rot = 0; st.diag = 0; st.last_inc = 0; st.cur_max = 0; st.first_max_cell = 0; st.n_cells = 1; st.cgen = 0; st.hashed = false; st.end_c = -1;
    g.sync();
}

// Group collectives issued by the WHOLE warp (full member mask, segments of GW lanes). The four groups of a warp only stay on one
// instruction stream if every collective inside the step is executed by all 32 lanes: with per-group member masks the hardware issues a
// vote / shuffle / barrier once per distinct mask, the groups drift apart and each instruction serves one group (ncu: 6.9 active lanes).
template <int GW> struct GrpW {
    int shift; int lane;
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(0xffffffffu, p) >> shift) & ((1u << GW) - 1u); }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(0xffffffffu, v, src, GW); }
    template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(0xffffffffu, v, d, GW); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, int d) const { return __shfl_xor_sync(0xffffffffu, v, d, GW); }
    __device__ __forceinline__ unsigned below() const { return (1u << lane) - 1u; }
};

// One diagonal for every group of the warp at once. ALL 32 lanes call it; `act` says whether the lane's group has a running extension.
// Control flow around collectives is warp-uniform (loops run to the warp-wide maximum trip count, early exits become per-group flags);
// what a group computes is exactly what the one-group-at-a-time version computed.
// Returns, per group: 0 to continue, 1 when the extension has ended (patience / nothing left / exact early exit), DP_DEFER on overflow.
template <class CFG> __device__ __forceinline__ int gd_step(const WdCtx& C, const GdSlab& S, GdState& st, const GrpW<CFG::GW>& g, bool act) {
    constexpr int GW = CFG::GW; constexpr unsigned FULL = 0xffffffffu;
    const DpGraph& G = *C.G;
    const int max_level = G.n_levels - 1, max_seq = C.seq_len;
    const int dir = C.pos ? 1 : -1; const int end_seq = C.pos ? max_seq : 0;
    const void* el = C.pos ? G.out_adj4 : G.in_adj4; const void* jl = C.pos ? G.jf4 : G.jb4;
    GdEntry* m1 = S.l[st.rot]; GdEntry* m2 = S.l[(st.rot + 2) % 3]; GdEntry* mt = S.l[(st.rot + 1) % 3];
    const int lane = g.lane; const int n_m1 = act ? st.n_m1 : 0, n_m2 = act ? st.n_m2 : 0;
    int rc = 0; int diag = st.diag;
    if (act) { diag = ++st.diag; if (diag - st.last_inc > 40 || (n_m1 == 0 && n_m2 == 0)) { rc = 1; act = false; } }
    {   // exact early exit (see extend_warp.cuh)
        bool live = false;
        if (act) { for (int i = lane; i < n_m1; i += GW) { const GdEntry& e = m1[i]; live |= (e.y != end_seq) || e.pad || e.jdeg; }
                   for (int i = lane; i < n_m2; i += GW) live |= (m2[i].y != end_seq); }
        const bool any_live = g.any(live);
        if (act && !any_live) { rc = 1; act = false; }
    }
    bool ovf = false, saw_jump = false;
    if (act) for (int i = lane; i < n_m2; i += GW) {
        const GdEntry pc = m2[i]; const int pcx = C.start_level + pc.xr;
        const int nx = pcx + dir, ny = pc.y + dir;
        if (nx > max_level || ny > max_seq || nx < 0 || ny < 0) continue;
        const uint8_t sc = C.pos ? C.seq[pc.y] : C.seq[pc.y - 1];
        if (pc.deg > WD_SUB) { ovf = true; continue; }
        for (int k = 0; k < (int)pc.deg; k++) {
            const I4 a = ld4(el, pc.k0 + k); const uint32_t pk = (uint32_t)a.z;
            const int nz = C.pos ? (int)((pk >> 8) & 255u) : (int)(pk & 255u);
            const int ti = gd_touch<CFG>(S, C.start_level, nx, ny, nz, a.y); if (ti < 0) { ovf = true; break; }
            atomicMax(&S.kD[ti], wd_key(pc.D + ((uint8_t)(pk >> 16) == sc ? 2 : -5), wd_mkseq(0, i, k, 0)));
        }
    }
    if (act) for (int i = lane; i < n_m1; i += GW) {
        const GdEntry pc = m1[i]; const int pcx = C.start_level + pc.xr;
        { const int gy = pc.y + dir; const bool ok = C.pos ? (pcx <= max_level && gy <= max_seq) : (pcx >= 0 && gy >= 0);
          if (ok) { const int ti = gd_touch<CFG>(S, C.start_level, pcx, gy, pc.z, pc.node); if (ti < 0) ovf = true; else {
              atomicMax(&S.kGG[ti], wd_key(pc.D - 6, wd_mkseq(1, i, 0, 0)));
              atomicMax(&S.kGG[ti], wd_key(pc.GG <= DP_NEG ? DP_NEG : pc.GG - 2, wd_mkseq(1, i, 0, 1))); } } }
        if ((int)pc.deg + (int)pc.jdeg > WD_SUB) { ovf = true; continue; }
        { const int sx = pcx + dir; const bool ok = C.pos ? (sx <= max_level && pc.y <= max_seq) : (sx >= 0 && pc.y >= 0);
          if (ok) for (int k = 0; k < (int)pc.deg; k++) {
              const I4 a = ld4(el, pc.k0 + k); const uint32_t pk = (uint32_t)a.z;
              const int nz = C.pos ? (int)((pk >> 8) & 255u) : (int)(pk & 255u); const bool gapEdge = ((uint8_t)(pk >> 16) == '_');
              const int ti = gd_touch<CFG>(S, C.start_level, sx, pc.y, nz, a.y); if (ti < 0) { ovf = true; break; }
              atomicMax(&S.kSG[ti], wd_key(gapEdge ? DP_NEG : pc.D - 6, wd_mkseq(1, i, k, 0)));
              atomicMax(&S.kSG[ti], wd_key(pc.SG <= DP_NEG ? DP_NEG : (gapEdge ? pc.SG : pc.SG - 2), wd_mkseq(1, i, k, 1)));
              if (gapEdge) atomicMax(&S.kD[ti], wd_key(pc.D, wd_mkseq(1, i, k, 0)));
          } }
        for (int j = 0; j < (int)pc.jdeg; j++) {
            const I4 jp = ld4(jl, pc.j0 + j);
            const int jx = pcx + dir * jp.w;
            const bool ok = C.pos ? (jx <= max_level && pc.y <= max_seq) : (jx >= 0 && pc.y >= 0);
            if (!ok) continue;
            const int ti = gd_touch<CFG>(S, C.start_level, jx, pc.y, jp.z, jp.y); if (ti < 0) { ovf = true; break; }
            atomicMax(&S.kD[ti], wd_key(pc.D, wd_mkseq(1, i, (int)pc.deg + j, 0)));
            if (jp.w > 1) saw_jump = true;   // a jump over one level lands on the same anti-diagonal as the '_' edge itself: still no revisits
        }
    }
    g.sync();
    int n_td = act ? *S.cnt : 0;
    { const bool any_ovf = g.any(ovf); if (act && (any_ovf || n_td > CFG::TD)) { rc = DP_DEFER; act = false; n_td = 0; } }
    {   // first jump over more than one level in this extension: from now on cells can be revisited, so build the (x, y, z) -> cell map
        const bool any_jump = g.any(saw_jump); const bool need = act && !st.hashed && any_jump;
        if (__any_sync(FULL, need)) {
            uint32_t cgen = 0;
            if (need && lane == 0) { cgen = C.gens[0] + 1; if (cgen >= (1u << 15)) cgen = 0; C.gens[0] = cgen ? cgen : 1; }
            cgen = g.shfl(cgen, 0);
            const bool clr = need && cgen == 0;
            if (__any_sync(FULL, clr)) { if (clr) for (int i = lane; i < CFG::HASH; i += GW) C.hash[i] = 0; __syncwarp(); }
            if (clr) cgen = 1;
            if (need) { for (int i = lane; i < st.n_cells; i += GW) { const DpCell& c = C.cells[i]; gd_insert_cell<CFG>(C, cgen, c.x, c.y, c.z, i); } st.cgen = cgen; st.hashed = true; }
            __syncwarp();
        }
    }
    // rank the touched slots in (x, y, z) order (the packed key is monotone in it)
    for (int i = lane; i < n_td; i += GW) {
        const uint32_t k = S.tkey[S.slots[i]]; int rank = 0;
        for (int j = 0; j < n_td; j++) rank += (S.tkey[S.slots[j]] < k) ? 1 : 0;
        S.order[rank] = S.slots[i];
    }
    g.sync();
    int n_mt = 0; int run_max = st.cur_max; int new_first = -1; bool any_inc = false;
    const int n_td_w = __reduce_max_sync(FULL, n_td);
    for (int base = 0; base < n_td_w; base += GW) {
        const int oi = base + lane; const bool in = act && oi < n_td;
        int selD = DP_NEG, selGG = DP_NEG, selSG = DP_NEG; uint32_t kD = 0, kGG = 0, kSG = 0; int selfmat = 0; int tx = 0, ty = 0, tz = 0, tn = 0;
        if (in) {
            const int ti = S.order[oi]; const uint32_t tk = S.tkey[ti] - 1u; tx = (int)(tk >> 19) - 2048 + C.start_level; ty = (int)((tk >> 8) & 2047u); tz = (int)(tk & 255u); tn = S.tnode[ti];
            kGG = S.kGG[ti]; kSG = S.kSG[ti]; kD = S.kD[ti];
            if (kGG) selGG = wd_score(kGG);
            if (kSG) selSG = wd_score(kSG);
            const bool haveD = kD != 0; if (haveD) selD = wd_score(kD);
            if (!haveD || selGG > selD) { selD = selGG; selfmat = 1; }
            if (selSG > selD) { selD = selSG; selfmat = 2; }
        }
        bool keep = in && selD >= -16;
        int ci = (keep && st.hashed) ? gd_find_cell<CFG>(C, st.cgen, tx, ty, tz) : -1;
        bool isNew = keep && ci < 0;
        const unsigned newmask = g.ballot(isNew);
        if (act && st.n_cells + __popc(newmask) > CFG::CELLS) { rc = DP_DEFER; act = false; }   // nothing of this group is stored from here on
        keep = keep && act; isNew = isNew && act;
        if (isNew) ci = st.n_cells + __popc(newmask & g.below());
        if (act) st.n_cells += __popc(newmask);
        int incl = keep ? selD : -1000000;
        for (int d = 1; d < GW; d <<= 1) { const int t2 = g.shfl_up(incl, d); if (lane >= d) incl = max(incl, t2); }
        int excl = g.shfl_up(incl, 1); if (lane == 0) excl = -1000000;
        const int running = max(run_max, excl);
        bool overwritten = false; int stD = DP_NEG, stGG = DP_NEG, stSG = DP_NEG; bool tie_counts = false;
        GdEntry ne;
        if (keep) {
            DpCell& c = C.cells[ci];
            if (isNew) { c.x = tx; c.y = (int16_t)ty; c.z = (int16_t)tz; c.pad = 0; c.D = c.GG = c.SG = (int16_t)DP_NEG; c.bD = c.bGG = c.bSG = dp_bt(-1, -1, -1); }
            if (isNew || c.D < selD) { overwritten = !isNew; c.D = (int16_t)selD; c.bD = selfmat ? dp_bt(ci, -1, selfmat) : gd_decode(C, m1, m2, kD, 0); }
            if (isNew || c.GG < selGG) { overwritten = !isNew; c.GG = (int16_t)selGG; c.bGG = kGG ? gd_decode(C, m1, m2, kGG, 1) : dp_bt(-1, -1, -1); }
            if (isNew || c.SG < selSG) { overwritten = !isNew; c.SG = (int16_t)selSG; c.bSG = kSG ? gd_decode(C, m1, m2, kSG, 2) : dp_bt(-1, -1, -1); }
            if (ty == end_seq) c.pad = 1;
            stD = c.D; stGG = c.GG; stSG = c.SG;
            if (overwritten) for (int i = 0; i < n_m1; i++) if (m1[i].cell == ci) { m1[i].D = (int16_t)stD; m1[i].GG = (int16_t)stGG; m1[i].SG = (int16_t)stSG; }   // see extend_warp.cuh
            if (isNew && st.hashed) gd_insert_cell<CFG>(C, st.cgen, tx, ty, tz, ci);
            if (selD == running) {
                DpBT step = c.bD;
                if (step.src == ci) step = (step.mat == 1) ? c.bGG : c.bSG;
                int prev = DP_NEG;
                if (step.src >= 0) { const DpCell& pc = C.cells[step.src]; prev = step.mat == 0 ? pc.D : (step.mat == 1 ? pc.GG : pc.SG); }
                tie_counts = (selD - prev) != 0;
            }
            ne.cell = (uint16_t)ci; ne.node = tn; ne.xr = (int16_t)(tx - C.start_level); ne.y = (int16_t)ty; ne.z = (uint8_t)tz; ne.D = (int16_t)stD; ne.GG = (int16_t)stGG; ne.SG = (int16_t)stSG; ne.pad = (uint8_t)((__ldg(G.node_gapflags + tn) >> (C.pos ? 0 : 1)) & 1);
            int deg, jdeg; wd_adj(C, tn, ne.k0, deg, ne.j0, jdeg); ne.deg = (uint8_t)min(deg, 255); ne.jdeg = (uint8_t)min(jdeg, 255);
        }
        { const unsigned em = g.ballot(keep && ty == end_seq); const int first_end = g.shfl(ci, em ? __ffs(em) - 1 : 0); if (em && st.end_c < 0) st.end_c = first_end; }   // first sequence-complete cell: the end-cell scan starts there
        if (g.any(keep && (selD > running || tie_counts || overwritten))) any_inc = true;
        const int chunk_max = g.shfl(incl, GW - 1);
        { const unsigned m = g.ballot(keep && selD == chunk_max); const int first_at_max = g.shfl(ci, m ? __ffs(m) - 1 : 0);
          if (act && chunk_max > run_max) { new_first = first_at_max; run_max = chunk_max; } }
        const unsigned km = g.ballot(keep);
        if (act && n_mt + __popc(km) > CFG::LIST) { rc = DP_DEFER; act = false; }
        if (keep && act) mt[n_mt + __popc(km & g.below())] = ne;
        if (act) n_mt += __popc(km);
        g.sync();
    }
    if (act) {
        for (int i = lane; i < n_td; i += GW) { const int sl = S.slots[i]; S.tkey[sl] = 0; S.kD[sl] = 0; S.kGG[sl] = 0; S.kSG[sl] = 0; }
        if (lane == 0) *S.cnt = 0;
        if (run_max > st.cur_max) { st.cur_max = run_max; st.first_max_cell = new_first; }
        if (any_inc) st.last_inc = diag;
    } else n_mt = 0;
    g.sync();
    {   // keep cells within 15 of the best stored D (stable compaction)
        int mx = -1000000; for (int i = lane; i < n_mt; i += GW) mx = max(mx, (int)mt[i].D);
        for (int d = GW / 2; d; d >>= 1) mx = max(mx, g.shfl_xor(mx, d));
        const int n_mt_w = __reduce_max_sync(FULL, n_mt);
        int w = 0;
        for (int base = 0; base < n_mt_w; base += GW) {
            const int i = base + lane; const bool in = i < n_mt; GdEntry e; if (in) e = mt[i];
            const bool k = in && (mx - e.D <= 15);
            const unsigned m = g.ballot(k);
            g.sync();
            if (k) mt[w + __popc(m & g.below())] = e;
            w += __popc(m);
            g.sync();
        }
        if (act) { st.n_m2 = n_m1; st.n_m1 = w; st.rot = (st.rot + 1) % 3; }
    }
    return rc;
}

// end cell + backtrace (lane 0 of the group); returns rc and fills res
template <class CFG> __device__ __noinline__ int gd_finish(const WdCtx& C, const GdState& st, const Grp<CFG::GW>& g, int32_t* out_edge, uint8_t* out_s, DpResult& res) {
    constexpr int GW = CFG::GW;
    const DpGraph& G = *C.G; const int end_seq = C.pos ? C.seq_len : 0; const int lane = g.lane;
    res.n_cols = 0; res.n_lvl = 0; res.far_y = C.start_seq;
    int end_cell = -1;
    {
        int bs = DP_NEG - 1, bx = -1, bz = -1, bc = -1;
        if (st.end_c >= 0) for (int i = st.end_c + lane; i < st.n_cells; i += GW) {
            const DpCell& c = C.cells[i];
            if (c.pad == 1 && c.y == end_seq) { const int s = c.D; if (bc < 0 || s > bs || (s == bs && dp_key_less(c.x, c.z, bx, bz))) { bs = s; bx = c.x; bz = c.z; bc = i; } }
        }
        if (st.end_c >= 0) for (int d = GW / 2; d; d >>= 1) {
            const int os = g.shfl_xor(bs, d), ox = g.shfl_xor(bx, d), oz = g.shfl_xor(bz, d), oc = g.shfl_xor(bc, d);
            if (oc >= 0 && (bc < 0 || os > bs || (os == bs && dp_key_less(ox, oz, bx, bz)))) { bs = os; bx = ox; bz = oz; bc = oc; }
        }
        if (bc >= 0) end_cell = bc; else if (st.cur_max > 0) end_cell = st.first_max_cell;
    }
    if (end_cell < 0) return 0;
    int rc = 0;
    if (lane == 0) {
        int n = 0, n_lvl = 0; int cur = end_cell, mat = 0, guard = 0;
        res.far_y = C.cells[end_cell].y;
        auto emit = [&](int32_t e, uint8_t s) -> bool { if (n >= DP_EXT_CAP) return false; out_edge[n] = e; out_s[n] = s; n++; if (e >= 0) n_lvl++; return true; };
        while (rc == 0 && !(C.cells[cur].x == C.start_level && C.cells[cur].y == C.start_seq)) {
            if (++guard > 8 * CFG::CELLS) { rc = -5; break; }
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
            if (C.pos) for (int i = 0, j = n - 1; i < j; i++, j--) { const int32_t te = out_edge[i]; out_edge[i] = out_edge[j]; out_edge[j] = te; const uint8_t ts = out_s[i]; out_s[i] = out_s[j]; out_s[j] = ts; }
            res.n_cols = n; res.n_lvl = n_lvl;
        }
    }
    rc = g.shfl(rc, 0);
    res.n_cols = g.shfl(res.n_cols, 0); res.n_lvl = g.shfl(res.n_lvl, 0); res.far_y = g.shfl(res.far_y, 0);
    return rc;
}

} // namespace hlala
