// CUDA kernels of the alignment path + their launchers. Compiled for sm_100a only.
#include "chain_kernel.cuh"
#include "extend_dp.h"
#include "extend_warp.cuh"
#include "extend_group.cuh"
#include "extend_lean_kernel.cuh"
#include "pair_kernel.cuh"
#include "align_kernels.h"
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>

namespace hlala {

// ---------------------------------------------------------------------------------------------------------------
// finalize: [left padding][left extension][seed][right extension][right padding] -> global chain record + LL
//   verboseSeedChain::extendWithOtherSeedChain / extendToFullSequenceLength (verboseSeedChain.cpp:23-144)
//   extensionAligner::scoreOneAlignment (extensionAligner.cpp:52-182); the sum runs left to right in one lane so that
//   every rounding step is the reference's.
struct ExtView { const int32_t* edge; const uint8_t* s; int n; int n_lvl; };   // n_lvl = columns that carry a level

__device__ void finalize_chain(const ChainParams& P, const WarpSlab& S, int slot, int64_t rd0, int rdlen,
                               const int32_t* seed_edge, const uint8_t* seed_g, const uint8_t* seed_s, int n_seed,
                               int seq_begin, int seq_end, int l_first, int l_last, ExtView L, ExtView R, int lane) {
    const DevGraph& G = P.g; const DevBatch& B = P.b;
    // after extension: sequence_begin / sequence_end of the chain (extendWithOtherSeedChain)
    int ext_l_bases = 0, ext_r_bases = 0;
    for (int i = lane; i < L.n; i += 32) ext_l_bases += (L.s[i] != '_');
    for (int i = lane; i < R.n; i += 32) ext_r_bases += (R.s[i] != '_');
    for (int d = 16; d; d >>= 1) { ext_l_bases += __shfl_xor_sync(0xffffffffu, ext_l_bases, d); ext_r_bases += __shfl_xor_sync(0xffffffffu, ext_r_bases, d); }
    const int padL = seq_begin - ext_l_bases;
    const int padR = rdlen - 1 - (seq_end + ext_r_bases);
    const int n_total = padL + L.n + n_seed + R.n + padR;
    if (padL < 0 || padR < 0 || n_total > P.maxcol || n_total > P.slab_cols) {
        const bool cap = padL >= 0 && padR >= 0;
        if (lane == 0) {
            if (cap && P.tier == 0 && n_total <= P.maxcol) { P.status[slot] = CH_DEFERRED; int idx = atomicAdd(P.defer_count, 1); P.defer_slots[idx] = slot; }
            else { P.status[slot] = cap ? HLALA_E_CAPACITY_DEV : HLALA_E_INVARIANT_DEV; P.n_cols[slot] = 0; atomicAdd(P.error_count, 1); }
        }
        return;
    }
    const size_t cbase = (size_t)(slot - P.slot_base) * P.maxcol;
    int32_t* oe = P.c_edge + cbase; uint8_t* os = P.c_schar + cbase; uint8_t* of = P.c_fromseed + cbase;
    uint8_t* kind = S.gA; uint8_t* qv = S.sA;      // per-column scoring class and quality, consumed by lane 0 below
    // one pass over output columns; idx of the read base under a column = number of non-gap read characters before it
    int carry = 0;
    for (int base = 0; base < n_total; base += 32) {
        int k = base + lane; bool in = k < n_total;
        int32_t e = -1; uint8_t sc = '_', gc = '_', fs = 0;
        if (in) {
            int j = k;
            if (j < padL) { sc = B.bases[rd0 + j]; }
            else if ((j -= padL) < L.n) { e = L.edge[j]; sc = L.s[j]; gc = e >= 0 ? (uint8_t)(G.edge_pack[e] >> 16) : (uint8_t)'_'; }
            else if ((j -= L.n) < n_seed) { e = seed_edge[j]; sc = seed_s[j]; gc = seed_g[j]; fs = 1; }
            else if ((j -= n_seed) < R.n) { e = R.edge[j]; sc = R.s[j]; gc = e >= 0 ? (uint8_t)(G.edge_pack[e] >> 16) : (uint8_t)'_'; }
            else { j -= R.n; sc = B.bases[rd0 + (rdlen - padR) + j]; }
            oe[k] = e; os[k] = sc; of[k] = fs;
        }
        bool hasb = in && sc != '_';
        unsigned m = __ballot_sync(0xffffffffu, hasb);
        if (in) {
            uint8_t kd = 0, q = 0;
            if (hasb) {
                int idx = carry + __popc(m & ((1u << lane) - 1));
                q = B.quals[rd0 + idx];
                kd = (gc == '_') ? 1 : (sc == gc ? 3 : 4);
            } else kd = (gc == '_') ? 0 : 2;
            kind[k] = kd; qv[k] = q;
        }
        carry += __popc(m);
    }
    __syncwarp();
    if (lane == 0) {
        const ScoreTables& T = P.long_mode ? c_tables_long : c_tables;
        double ll = 0;
        for (int k = 0; k < n_total; k++) {
            switch (kind[k]) {
            case 1: ll += T.ins_term; break;                                                     // insertion (extensionAligner.cpp:119)
            case 2: ll += T.rate_deletion; break;                                                // deletion  (:163)
            case 3: ll += T.rate_match_mismatch; ll += T.log_match[qv[k]]; break;         // (:125,139)
            case 4: ll += T.rate_match_mismatch; ll += T.log_mismatch[qv[k]]; break;      // (:125,146)
            default: break;                                                                             // gap against gap
            }
        }
        P.ll[slot] = ll; P.n_cols[slot] = n_total; P.status[slot] = CH_OK;
        P.first_level[slot] = l_first - L.n_lvl; P.last_level[slot] = l_last + R.n_lvl;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 0: one thread per read. Strand rule (processBAM.cpp:3216), PRG start/stop of every BAM record
// (alignment_get_startstop_PRGcoordinates, processBAM.cpp:3840) and the identical-coordinates rule (:3234). The reference aligns
// a chain and only then discards it as a duplicate; the outcome does not depend on the discarded work, so it is skipped here.
__global__ void k_prepare(ChainParams P) {
    const DevGraph& G = P.g; const DevBatch& B = P.b;
    const int r = P.read_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.read_end) return;
    const int s0 = B.chain_off[r], s1 = B.chain_off[r + 1];
    const int prim = B.read_primary[r];
    for (int s = s0; s < s1; s++) {
        const int c = B.chain_order[s];
        int st = CH_TODO;
        if (prim < 0) st = HLALA_E_INVARIANT_DEV;
        else if ((B.chain_flag[c] ^ B.chain_flag[B.chain_order[prim]]) & 0x10) st = CH_SKIPPED_STRAND;
        else {
            const int contig = B.chain_contig[c]; const int pos = B.chain_pos[c];
            if (contig < 0 || contig >= G.n_contigs) st = HLALA_E_ARG_DEV;
            else {
                int reflen = 0;
                for (int k = B.cigar_off[c]; k < B.cigar_off[c + 1]; k++) { uint32_t cg = B.cigar[k]; int op = cg & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += (int)(cg >> 4); }
                const int64_t cbase = G.contig_off[contig]; const int clen = (int)(G.contig_off[contig + 1] - cbase);
                if (pos < 0 || reflen < 1 || pos + reflen > clen) st = HLALA_E_INVARIANT_DEV;
                else {
                    const int idf = G.contig_level[cbase + pos], idl = G.contig_level[cbase + pos + reflen - 1];
                    P.id_first[s] = idf; P.id_last[s] = idl;
                    if (P.dedup) { for (int t = s0; t < s; t++) if (P.status[t] == CH_TODO && P.id_first[t] == idf && P.id_last[t] == idl) { st = CH_DUPLICATE; break; } }
                }
            }
        }
        P.status[s] = st;
        if (st == CH_TODO) { int idx = atomicAdd(P.todo_count, 1); P.todo_slots[idx] = s; }
        else { P.n_cols[s] = 0; P.ll[s] = 0; if (st < 0) atomicAdd(P.error_count, 1); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K1_WARPS * 32) k_chain_seed(ChainParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSlab S = slab_of(P, smem, warp, K1_WARPS);
    if (lane == 0) { mbar_init(S.mbar, 1); *S.mbar_phase = 0; }
    __syncwarp();
    const DevBatch& B = P.b;
    const int nw = gridDim.x * K1_WARPS;
    const int n_todo = P.tier == 0 ? *P.todo_count : *P.defer_count;
    const int32_t* work = P.tier == 0 ? P.todo_slots : P.defer_slots;
    for (int ti = blockIdx.x * K1_WARPS + warp; ti < n_todo; ti += nw) {
        const int slot = work[ti];
        const int r = B.slot_read[slot];
        const int c = B.chain_order[slot];
        int rc = 0;
        const int64_t rd0 = B.read_off[r]; const int rdlen = (int)(B.read_off[r + 1] - rd0);
        ColBuf A{S.lvlA, S.gA, S.sA}, Bc{S.lvlB, S.gB, S.sB};
        int start_raw = -1, stop_raw = -1, n = 0, idf = -1, idl = -1;
        if (rc == 0) { n = expand_cigar(P, c, rd0, rdlen, A, lane, start_raw, stop_raw, idf, idl); if (n < 0) rc = n; }
        if (rc == 0 && !(start_raw < stop_raw)) rc = HLALA_E_INVARIANT_DEV;     // processBAM.cpp:5245
        if (rc == 0) { n = trim_and_fill(P, A, n, Bc, lane, start_raw, stop_raw); if (n < 0) rc = n; }
        if (rc == 0) n = clean_columns(Bc, n, A, lane);
        if (rc == 0) n = restrict_columns(P, Bc, n, A, lane, start_raw, stop_raw);
        if (rc == 0) rc = P.bt16 ? viterbi_backtrace<true>(P, Bc, n, S, S.lvlA, lane) : viterbi_backtrace<false>(P, Bc, n, S, S.lvlA, lane);
        if (rc != 0) {
            if (lane == 0) {
                if (rc == HLALA_E_CAPACITY_DEV && P.tier == 0) { P.status[slot] = CH_DEFERRED; int idx = atomicAdd(P.defer_count, 1); P.defer_slots[idx] = slot; }
                else { P.status[slot] = rc; P.n_cols[slot] = 0; P.ll[slot] = 0; atomicAdd(P.error_count, 1); }
            }
            __syncwarp();
            continue;
        }
        if (lane == 0) { P.seed_begin[slot] = start_raw; P.seed_end[slot] = stop_raw; }
        const int l_first = Bc.lvl[0], l_last = Bc.lvl[n - 1];
        // extensionAligner::extendSeedChain (extensionAligner.cpp:220-225, 271-276): DP only where the seed leaves read bases uncovered
        const bool need_left = (start_raw != 0) && (l_first > 0);
        const bool need_right = (stop_raw != rdlen - 1) && (l_last + 1 < P.g.n_levels - 1);
        if (P.do_extension && (need_left || need_right)) {
            int32_t* oe = P.c_edge + (size_t)(slot - P.slot_base) * P.maxcol; uint8_t* os = P.c_schar + (size_t)(slot - P.slot_base) * P.maxcol;
            for (int i = lane; i < n; i += 32) { oe[i] = S.lvlA[i]; os[i] = Bc.s[i]; }
            if (lane == 0) {
                P.n_cols[slot] = n; P.first_level[slot] = l_first; P.last_level[slot] = l_last; P.status[slot] = CH_PENDING_EXT;
                int idx = atomicAdd(P.pending_count, 1); P.pending_slots[idx] = slot;
                if (need_left) { const int i = atomicAdd(P.dp_task_count, 1); const int bin = 255 - min(start_raw, 255); P.dp_tasks[i] = 2 * idx; P.dp_task_bin[i] = (uint8_t)bin; P.dp_task_key[i] = l_first; atomicAdd(P.dp_task_hist + bin, 1); }
                if (need_right) { const int i = atomicAdd(P.dp_task_count, 1); const int bin = 255 - min(rdlen - 1 - stop_raw, 255); P.dp_tasks[i] = 2 * idx + 1; P.dp_task_bin[i] = (uint8_t)bin; P.dp_task_key[i] = l_last + 1; atomicAdd(P.dp_task_hist + bin, 1); }
            }
            __syncwarp();
            continue;
        }
        ExtView none{nullptr, nullptr, 0, 0};
        finalize_chain(P, S, slot, rd0, rdlen, S.lvlA, Bc.g, Bc.s, n, start_raw, stop_raw, l_first, l_last, none, none, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2: extension DP, one thread per (pending chain, side). extend_dp.h documents the algorithm.
__global__ void __launch_bounds__(64) k_extend(ExtParams E) {
    const ChainParams& P = E.C; const DevGraph& G = P.g; const DevBatch& B = P.b;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= E.n_dp_threads) return;
    DpScratch S = dp_carve(E.dp_scratch + (size_t)tid * dp_scratch_bytes());
    DpGraph dg; dg.n_levels = G.n_levels; dg.level_node_off = G.level_node_off; dg.edge_pack = G.edge_pack;
    dg.node_out_off = G.node_out_off; dg.node_out = G.node_out; dg.node_in_off = G.node_in_off; dg.node_in = G.node_in;
    dg.path_off = G.path_off; dg.path_edges = G.path_edges; dg.path_from = G.path_from; dg.path_to = G.path_to;
    dg.jump_fwd_off = G.jump_fwd_off; dg.jump_fwd_path = G.jump_fwd_path; dg.jump_bwd_off = G.jump_bwd_off; dg.jump_bwd_path = G.jump_bwd_path;
    for (int t = tid; t < 2 * E.n_pending; t += E.n_dp_threads) {
        const int slot = P.pending_slots[t >> 1]; const int side = t & 1;
        const int r = B.slot_read[slot]; const int64_t rd0 = B.read_off[r]; const int rdlen = (int)(B.read_off[r + 1] - rd0);
        const int n = P.n_cols[slot]; const int sb = P.seed_begin[slot], se = P.seed_end[slot];
        const int l_first = P.first_level[slot], l_last = P.last_level[slot];
        const int32_t* se_edge = P.c_edge + (size_t)(slot - P.slot_base) * P.maxcol;
        if (E.only_deferred && E.ext_rc[t] != DP_DEFER) continue;
        E.ext_n[t] = 0; E.ext_nlvl[t] = 0; E.ext_rc[t] = 0;
        DpResult res; int rc = 0;
        if (side == 0) {
            if (!(sb != 0 && l_first > 0)) continue;
            int z = (int)(G.edge_pack[se_edge[0]] & 255u);                       // From node of the first seed edge
            rc = dp_extend(dg, B.bases + rd0, rdlen, sb, l_first, z, false, S, E.ext_edge + (size_t)t * DP_EXT_CAP, E.ext_s + (size_t)t * DP_EXT_CAP, res);
        } else {
            if (!(se != rdlen - 1 && l_last + 1 < G.n_levels - 1)) continue;
            int z = (int)((G.edge_pack[se_edge[n - 1]] >> 8) & 255u);            // To node of the last seed edge
            rc = dp_extend(dg, B.bases + rd0, rdlen, se + 1, l_last + 1, z, true, S, E.ext_edge + (size_t)t * DP_EXT_CAP, E.ext_s + (size_t)t * DP_EXT_CAP, res);
        }
        E.ext_rc[t] = rc;
        if (rc == 0) { E.ext_n[t] = res.n_cols; E.ext_nlvl[t] = res.n_lvl; }
    }
}

// Kernel 2 (fast path): extension DP, one warp per (pending chain, side); see extend_warp.cuh.
__host__ __device__ inline size_t wd_hbm_bytes() { return sizeof(DpCell) * (size_t)DP_CELL_CAP + 4 * (size_t)DP_HASH_CAP + 16; }
template <class CFG> __global__ void __launch_bounds__(CFG::WARPS * 32) k_extend_warp(ExtParams E) {
    extern __shared__ __align__(16) unsigned char smem[];
    const ChainParams& P = E.C; const DevGraph& G = P.g; const DevBatch& B = P.b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * CFG::WARPS + warp;
    if (gw >= E.n_wd_warps) return;
    WdSlab S = wd_carve<CFG>(smem + (size_t)warp * wd_slab_bytes<CFG>());
    unsigned char* hb = E.wd_scratch + (size_t)gw * wd_hbm_bytes();
    DpGraph dg; dg.n_levels = G.n_levels; dg.level_node_off = G.level_node_off; dg.edge_pack = G.edge_pack;
    dg.node_out_off = G.node_out_off; dg.node_out = G.node_out; dg.node_in_off = G.node_in_off; dg.node_in = G.node_in;
    dg.path_off = G.path_off; dg.path_edges = G.path_edges; dg.path_from = G.path_from; dg.path_to = G.path_to;
    dg.jump_fwd_off = G.jump_fwd_off; dg.jump_fwd_path = G.jump_fwd_path; dg.jump_bwd_off = G.jump_bwd_off; dg.jump_bwd_path = G.jump_bwd_path;
    dg.adj4 = G.adj4; dg.out_adj4 = G.out_adj4; dg.in_adj4 = G.in_adj4; dg.jf4 = G.jf4; dg.jb4 = G.jb4; dg.node_gapflags = G.node_gapflags;
    WdCtx C; C.G = &dg; C.cells = (DpCell*)hb; C.hash = (uint32_t*)(hb + sizeof(DpCell) * (size_t)DP_CELL_CAP); C.gens = C.hash + DP_HASH_CAP;
    const int n_in = E.in_list ? *E.in_count : 2 * E.n_pending;
    for (;;) {
        int ti = 0; if (lane == 0) ti = atomicAdd(E.pop, 1); ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= n_in) break;
        const int t = E.in_list ? E.in_list[ti] : ti;
        const int slot = P.pending_slots[t >> 1]; const int side = t & 1;
        const int r = B.slot_read[slot]; const int64_t rd0 = B.read_off[r]; const int rdlen = (int)(B.read_off[r + 1] - rd0);
        const int n = P.n_cols[slot]; const int sb = P.seed_begin[slot], se = P.seed_end[slot];
        const int l_first = P.first_level[slot], l_last = P.last_level[slot];
        const int32_t* se_edge = P.c_edge + (size_t)(slot - P.slot_base) * P.maxcol;
        DpResult res; res.n_cols = 0; res.n_lvl = 0; res.far_y = 0; int rc = 0; bool run = false;
        C.seq = B.bases + rd0; C.seq_len = rdlen;
        if (side == 0) {
            if (sb != 0 && l_first > 0) { run = true; C.start_seq = sb; C.start_level = l_first; C.start_z = (int)(G.edge_pack[se_edge[0]] & 255u); C.pos = false; C.start_node = G.level_node_off[l_first] + C.start_z; }
        } else {
            if (se != rdlen - 1 && l_last + 1 < G.n_levels - 1) { run = true; C.start_seq = se + 1; C.start_level = l_last + 1; C.start_z = (int)((G.edge_pack[se_edge[n - 1]] >> 8) & 255u); C.pos = true; C.start_node = G.level_node_off[l_last + 1] + C.start_z; }
        }
        if (run) rc = wd_extend<CFG>(C, S, E.ext_edge + (size_t)t * DP_EXT_CAP, E.ext_s + (size_t)t * DP_EXT_CAP, res, lane);
        if (lane == 0) { E.ext_rc[t] = rc; E.ext_n[t] = rc == 0 ? res.n_cols : 0; E.ext_nlvl[t] = rc == 0 ? res.n_lvl : 0; if (rc == DP_DEFER) E.out_list[atomicAdd(E.out_count, 1)] = t; }
        __syncwarp();
    }
}


// Kernel 2 (first tier): extension DP, one 8-lane group per (pending chain, side); the groups of a warp advance one diagonal each per
// iteration of a flat loop (see extend_group.cuh). Tasks are dealt round-robin to the groups of the persistent grid.
template <class CFG> __global__ void __launch_bounds__(CFG::WARPS * 32) k_extend_group(ExtParams E) {
    extern __shared__ __align__(16) unsigned char smem[];
    const ChainParams& P = E.C; const DevGraph& G = P.g; const DevBatch& B = P.b;
    constexpr int GW = CFG::GW, NG = CFG::NG;
    const int gcta = threadIdx.x / GW;                               // group inside the CTA
    const int gg = blockIdx.x * (CFG::WARPS * NG) + gcta;
    Grp<GW> g; g.lane = threadIdx.x % GW; g.shift = ((threadIdx.x & 31) / GW) * GW; g.mask = (GW == 32) ? 0xffffffffu : (((1u << (GW & 31)) - 1u) << g.shift);
    GrpW<GW> gw; gw.lane = g.lane; gw.shift = g.shift;
    GdSlab S = gd_carve<CFG>(smem + (size_t)gcta * gd_slab_bytes<CFG>());
    unsigned char* hb = E.gd_scratch + (size_t)gg * gd_hbm_bytes<CFG>();
    DpGraph dg; dg.n_levels = G.n_levels; dg.level_node_off = G.level_node_off; dg.edge_pack = G.edge_pack;
    dg.node_out_off = G.node_out_off; dg.node_out = G.node_out; dg.node_in_off = G.node_in_off; dg.node_in = G.node_in;
    dg.path_off = G.path_off; dg.path_edges = G.path_edges; dg.path_from = G.path_from; dg.path_to = G.path_to;
    dg.jump_fwd_off = G.jump_fwd_off; dg.jump_fwd_path = G.jump_fwd_path; dg.jump_bwd_off = G.jump_bwd_off; dg.jump_bwd_path = G.jump_bwd_path;
    dg.adj4 = G.adj4; dg.out_adj4 = G.out_adj4; dg.in_adj4 = G.in_adj4; dg.jf4 = G.jf4; dg.jb4 = G.jb4; dg.node_gapflags = G.node_gapflags;
    WdCtx C; C.G = &dg; C.cells = (DpCell*)hb; C.hash = (uint32_t*)(hb + sizeof(DpCell) * (size_t)CFG::CELLS); C.gens = C.hash + CFG::HASH;
    C.seq = nullptr; C.seq_len = 0; C.start_seq = C.start_level = C.start_z = C.start_node = 0; C.pos = false;
    GdState st{}; int phase = (gg < E.n_gd_groups) ? 0 : 2;        // 0: needs a task, 1: running, 2: out of tasks
    int t = 0; const int n_tasks = 2 * E.n_pending;
    for (;;) {
        if (phase == 0) {
            for (;;) {
                if (g.lane == 0) t = atomicAdd(E.pop, 1);
                t = g.shfl(t, 0);
                if (t >= n_tasks) { phase = 2; break; }
                const int slot = P.pending_slots[t >> 1]; const int side = t & 1;
                const int r = B.slot_read[slot]; const int64_t rd0 = B.read_off[r]; const int rdlen = (int)(B.read_off[r + 1] - rd0);
                const int n = P.n_cols[slot]; const int sb = P.seed_begin[slot], se = P.seed_end[slot];
                const int l_first = P.first_level[slot], l_last = P.last_level[slot];
                const int32_t* se_edge = P.c_edge + (size_t)(slot - P.slot_base) * P.maxcol;
                bool run = false;
                C.seq = B.bases + rd0; C.seq_len = rdlen;
                if (side == 0) {
                    if (sb != 0 && l_first > 0) { run = true; C.start_seq = sb; C.start_level = l_first; C.start_z = (int)(G.edge_pack[se_edge[0]] & 255u); C.pos = false; C.start_node = G.level_node_off[l_first] + C.start_z; }
                } else {
                    if (se != rdlen - 1 && l_last + 1 < G.n_levels - 1) { run = true; C.start_seq = se + 1; C.start_level = l_last + 1; C.start_z = (int)((G.edge_pack[se_edge[n - 1]] >> 8) & 255u); C.pos = true; C.start_node = G.level_node_off[l_last + 1] + C.start_z; }
                }
                if (!run) { if (g.lane == 0) { E.ext_rc[t] = 0; E.ext_n[t] = 0; E.ext_nlvl[t] = 0; } continue; }
                gd_init<CFG>(C, S, st, g);
                phase = 1; break;
            }
        }
        if (__all_sync(0xffffffffu, phase == 2)) break;
        {   // every lane of the warp takes the step together (groups without a running extension are predicated off inside)
            int rc = gd_step<CFG>(C, S, st, gw, phase == 1);
            if (phase == 1 && rc != 0) {
                DpResult res; res.n_cols = 0; res.n_lvl = 0; res.far_y = 0;
                if (rc == 1) rc = gd_finish<CFG>(C, st, g, E.ext_edge + (size_t)t * DP_EXT_CAP, E.ext_s + (size_t)t * DP_EXT_CAP, res);
                if (g.lane == 0) { E.ext_rc[t] = rc; E.ext_n[t] = rc == 0 ? res.n_cols : 0; E.ext_nlvl[t] = rc == 0 ? res.n_lvl : 0; if (rc == DP_DEFER) E.out_list[atomicAdd(E.out_count, 1)] = t; }
                g.sync();
                phase = 0;
            }
        }
    }
}

// Kernel 1b: splice extensions into the pending chains, pad, score (one warp per pending chain).
__global__ void __launch_bounds__(K1_WARPS * 32) k_chain_finish(ExtParams E) {
    extern __shared__ __align__(16) unsigned char smem[];
    const ChainParams& P = E.C; const DevGraph& G = P.g; const DevBatch& B = P.b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSlab S = slab_of(P, smem, warp, K1_WARPS);
    const int nw = gridDim.x * K1_WARPS;
    for (int pi = blockIdx.x * K1_WARPS + warp; pi < E.n_pending; pi += nw) {
        const int slot = P.pending_slots[pi];
        const int r = B.slot_read[slot]; const int64_t rd0 = B.read_off[r]; const int rdlen = (int)(B.read_off[r + 1] - rd0);
        const int n = P.n_cols[slot];
        const int32_t* ge = P.c_edge + (size_t)(slot - P.slot_base) * P.maxcol; const uint8_t* gs = P.c_schar + (size_t)(slot - P.slot_base) * P.maxcol;
        for (int i = lane; i < n; i += 32) { int32_t e = ge[i]; S.lvlA[i] = e; S.sB[i] = gs[i]; S.gB[i] = e >= 0 ? (uint8_t)(G.edge_pack[e] >> 16) : (uint8_t)'_'; }
        __syncwarp();
        int rcL = E.ext_rc[2 * pi], rcR = E.ext_rc[2 * pi + 1];
        if (rcL < 0 || rcR < 0) {
            if (lane == 0) { P.status[slot] = rcL < 0 ? rcL : rcR; P.n_cols[slot] = 0; P.ll[slot] = 0; atomicAdd(P.error_count, 1); }
            __syncwarp(); continue;
        }
        ExtView L{E.ext_edge + (size_t)(2 * pi) * DP_EXT_CAP, E.ext_s + (size_t)(2 * pi) * DP_EXT_CAP, E.ext_n[2 * pi], E.ext_nlvl[2 * pi]};
        ExtView R{E.ext_edge + (size_t)(2 * pi + 1) * DP_EXT_CAP, E.ext_s + (size_t)(2 * pi + 1) * DP_EXT_CAP, E.ext_n[2 * pi + 1], E.ext_nlvl[2 * pi + 1]};
        finalize_chain(P, S, slot, rd0, rdlen, S.lvlA, S.gB, S.sB, n, P.seed_begin[slot], P.seed_end[slot], P.first_level[slot], P.last_level[slot], L, R, lane);
    }
}

// edge ids -> canonical ordinals / levels / graph characters for the host-facing chain records
__global__ void k_export_chain_columns(DevGraph G, int n_chains, int maxcol, const int32_t* n_cols, const int32_t* first_level,
                                       const int32_t* c_edge, int32_t* out_level, int32_t* out_edge_ord, uint8_t* out_gchar) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_chains) return;
    const int n = n_cols[warp]; const size_t o = (size_t)warp * maxcol;
    int carry = 0; const int l0 = first_level[warp];
    for (int base = 0; base < n; base += 32) {
        int k = base + lane; bool in = k < n;
        int32_t e = in ? c_edge[o + k] : -1;
        unsigned m = __ballot_sync(0xffffffffu, in && e >= 0);
        if (in) {
            if (e >= 0) { out_level[o + k] = l0 + carry + __popc(m & ((1u << lane) - 1)); out_edge_ord[o + k] = G.edge_ord[e]; out_gchar[o + k] = (uint8_t)(G.edge_pack[e] >> 16); }
            else { out_level[o + k] = -1; out_edge_ord[o + k] = -1; out_gchar[o + k] = '_'; }
        }
        carry += __popc(m);
    }
}

// ---------------------------------------------------------------------------------------------------------------
cudaError_t upload_score_tables(const ScoreTables& t, const ScoreTables& t_long) {
    cudaError_t e = cudaMemcpyToSymbol(c_tables, &t, sizeof(ScoreTables)); if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_tables_long, &t_long, sizeof(ScoreTables));
}

cudaError_t launch_chain_seed(const ChainParams& P, int n_sm, cudaStream_t stream) {
    size_t slab = P.gslab ? k1_gslab_smem_bytes(P.win_cap, P.wcap) : k1_slab_bytes(P.slab_cols, P.pool_cap, P.win_cap, P.wcap);
    size_t smem = slab * K1_WARPS;
    {   // a function attribute belongs to the current device: set it on every launch (several GPUs may be driven from one process)
        cudaError_t e = cudaFuncSetAttribute(k_chain_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_seed, K1_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = ((long long)(P.slot_end - P.slot_base) + K1_WARPS - 1) / K1_WARPS;
    int grid = (int)std::min<long long>(want, (long long)n_sm * per_sm);   // persistent: a multiple of the SM count
    if (grid < 1) grid = 1;
    k_chain_seed<<<grid, K1_WARPS * 32, smem, stream>>>(P);
    return cudaGetLastError();
}

// warps a launch of the chain kernel can have resident with these capacities (ChainParams::gslab: one HBM slice per such warp)
int chain_seed_max_warps(const ChainParams& P, int n_sm) {
    size_t smem = (P.gslab ? k1_gslab_smem_bytes(P.win_cap, P.wcap) : k1_slab_bytes(P.slab_cols, P.pool_cap, P.win_cap, P.wcap)) * K1_WARPS;
    if (cudaFuncSetAttribute(k_chain_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int per_sm = 1; if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_seed, K1_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return n_sm * per_sm * K1_WARPS;
}

cudaError_t launch_prepare(const ChainParams& P, cudaStream_t stream) {
    int n = P.read_end - P.read_begin; if (n <= 0) return cudaSuccess;
    k_prepare<<<(n + 127) / 128, 128, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_extend(const ExtParams& E, cudaStream_t stream) {
    if (E.n_pending <= 0) return cudaSuccess;
    int threads = 64; int blocks = (E.n_dp_threads + threads - 1) / threads;
    k_extend<<<blocks, threads, 0, stream>>>(E);
    return cudaGetLastError();
}

template <class CFG> static int wd_occupancy(int n_sm) {
    size_t smem = wd_slab_bytes<CFG>() * CFG::WARPS; int per_sm = 1;
    if (cudaFuncSetAttribute(k_extend_warp<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_warp<CFG>, CFG::WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return n_sm * per_sm * CFG::WARPS;
}
int wd_warps_for(int n_sm) { return std::max(wd_occupancy<WdTiny>(n_sm), std::max(wd_occupancy<WdSmall>(n_sm), wd_occupancy<WdLarge>(n_sm))); }
size_t wd_warp_scratch_bytes() { return wd_hbm_bytes(); }

template <class CFG> static cudaError_t launch_wd(ExtParams E, int n_sm, cudaStream_t stream) {
    int warps = std::min(wd_occupancy<CFG>(n_sm), E.n_wd_warps);
    if (warps < CFG::WARPS) return cudaErrorInvalidConfiguration;
    E.n_wd_warps = warps;
    long long want = ((long long)2 * E.n_pending + CFG::WARPS - 1) / CFG::WARPS;
    int grid = (int)std::min<long long>(want, (long long)(warps / CFG::WARPS)); if (grid < 1) grid = 1;
    k_extend_warp<CFG><<<grid, CFG::WARPS * 32, wd_slab_bytes<CFG>() * CFG::WARPS, stream>>>(E);
    return cudaGetLastError();
}
// cfg 0 / 1 / 2: tiny / small / large shared-memory configuration; only_deferred: run only the tasks an earlier tier deferred
cudaError_t launch_extend_warp(const ExtParams& E0, int n_sm, int cfg, bool only_deferred, cudaStream_t stream) {
    if (E0.n_pending <= 0) return cudaSuccess;
    ExtParams E = E0; E.only_deferred = only_deferred ? 1 : 0;
    return cfg == 0 ? launch_wd<WdTiny>(E, n_sm, stream) : cfg == 1 ? launch_wd<WdSmall>(E, n_sm, stream) : launch_wd<WdLarge>(E, n_sm, stream);
}

template <class CFG> static int gd_occupancy_groups(int n_sm) {
    size_t smem = gd_slab_bytes<CFG>() * CFG::WARPS * CFG::NG; int per_sm = 1;
    if (cudaFuncSetAttribute(k_extend_group<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_group<CFG>, CFG::WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return n_sm * per_sm * CFG::WARPS * CFG::NG;
}
int gd_groups_for(int n_sm) { return gd_occupancy_groups<GdOct>(n_sm); }
size_t gd_group_scratch_bytes() { return gd_hbm_bytes<GdOct>(); }
cudaError_t launch_extend_group(const ExtParams& E0, int n_sm, cudaStream_t stream) {
    if (E0.n_pending <= 0) return cudaSuccess;
    typedef GdOct CFG;
    ExtParams E = E0;
    const int per_cta = CFG::WARPS * CFG::NG;
    int groups = std::min(gd_occupancy_groups<CFG>(n_sm), E.n_gd_groups);
    if (groups < per_cta) return cudaErrorInvalidConfiguration;
    long long want = ((long long)2 * E.n_pending + per_cta - 1) / per_cta;
    int grid = (int)std::min<long long>(want, (long long)(groups / per_cta)); if (grid < 1) grid = 1;
    E.n_gd_groups = grid * per_cta;
    k_extend_group<CFG><<<grid, CFG::WARPS * 32, gd_slab_bytes<CFG>() * per_cta, stream>>>(E);
    return cudaGetLastError();
}

// Extension tasks ordered by clipped length, longest first (one CTA; counting sort over the 256 bins filled by k_chain_seed): the threads of a
// warp then run extensions of similar length, and the longest extensions of a wave start first instead of forming its tail.
__global__ void __launch_bounds__(1024) k_sort_dp_tasks(const int32_t* tasks, const uint8_t* bin, const int32_t* hist, const int32_t* count, int32_t* sorted) {
    __shared__ int32_t cursor[256];
    if (threadIdx.x < 256) cursor[threadIdx.x] = hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) { int acc = 0; for (int b = 0; b < 256; b++) { const int c = cursor[b]; cursor[b] = acc; acc += c; } }
    __syncthreads();
    const int n = *count;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sorted[atomicAdd(&cursor[bin[i]], 1)] = tasks[i];
}
cudaError_t launch_sort_dp_tasks(const ChainParams& P, int32_t* sorted, cudaStream_t stream) {
    k_sort_dp_tasks<<<1, 1024, 0, stream>>>(P.dp_tasks, P.dp_task_bin, P.dp_task_hist, P.dp_task_count, sorted);
    return cudaGetLastError();
}

// Extension tasks ordered by the level they start from (radix sort of (level, task) pairs; n covers the unused slots, whose keys are large): the threads of a
// warp then walk neighbouring windows of the graph - the same level records, the same bubbles and jumps - instead of 32 unrelated ones.
cudaError_t sort_dp_tasks_by_level(const int32_t* keys_in, int32_t* keys_out, const int32_t* tasks_in, int32_t* tasks_out, int n, void* temp, size_t temp_bytes, size_t* need_bytes, cudaStream_t stream) {
    if (!temp) { size_t need = 0; cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, tasks_in, tasks_out, n, 0, 31, stream); *need_bytes = need; return e; }
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, tasks_in, tasks_out, n, 0, 31, stream);
}

// thread-per-extension tier (extend_lean.h): cfg 0 = LnStd over the task list, cfg 1 = LnBig over what LnStd deferred
int ln_threads_for_any(int n_sm) { return std::max(ln_threads_for<LnStd>(n_sm), ln_threads_for<LnBig>(n_sm)); }
size_t ln_thread_rec_bytes() { return sizeof(LnRec) * (size_t)(LN_CELLS + 1); }
size_t ln_thread_ahead_bytes() { return sizeof(LnAhead) * (size_t)LN_AHEAD + (size_t)LN_RING; }     // ring entries + per-anti-diagonal counts
cudaError_t launch_extend_lean(const ExtParams& E, int n_sm, int cfg, cudaStream_t stream) {
    if (E.n_pending <= 0) return cudaSuccess;
    return cfg == 0 ? launch_ln<LnStd>(E, n_sm, stream) : launch_ln<LnBig>(E, n_sm, stream);
}

// Algorithmic HBM bytes of the extension tasks of one wave (extension DP): per task that runs, per clipped read base: the base
// itself (1) + one level of graph window (adj4 record 16 + 1.1 out/in records of 16) + one output column (edge 4 + read char 1), plus 40 B
// of task scalars. Summed into a 64-bit counter; launched only when the session's timing is on.
__global__ void k_dp_task_bytes(ExtParams E, unsigned long long* out) {
    const ChainParams& P = E.C; const DevGraph& G = P.g; const DevBatch& B = P.b;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long b = 0;
    if (t < 2 * E.n_pending) {
        const int slot = P.pending_slots[t >> 1]; const int side = t & 1;
        const int r = B.slot_read[slot]; const int rdlen = (int)(B.read_off[r + 1] - B.read_off[r]);
        const int sb = P.seed_begin[slot], se = P.seed_end[slot]; const int l_first = P.first_level[slot], l_last = P.last_level[slot];
        int len = 0;
        if (side == 0) { if (sb != 0 && l_first > 0) len = sb; } else { if (se != rdlen - 1 && l_last + 1 < G.n_levels - 1) len = rdlen - 1 - se; }
        if (len > 0) b = 40ull + (unsigned long long)len * (1 + 16 + 18 + 5);
    }
    for (int d = 16; d; d >>= 1) b += __shfl_xor_sync(0xffffffffu, b, d);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, b);
}
cudaError_t launch_dp_task_bytes(const ExtParams& E, unsigned long long* out, cudaStream_t stream) {
    if (E.n_pending <= 0) return cudaSuccess;
    k_dp_task_bytes<<<(2 * E.n_pending + 255) / 256, 256, 0, stream>>>(E, out);
    return cudaGetLastError();
}

cudaError_t launch_chain_finish(const ExtParams& E, int n_sm, cudaStream_t stream) {
    if (E.n_pending <= 0) return cudaSuccess;
    size_t smem = (E.C.gslab ? k1_gslab_smem_bytes(E.C.win_cap, E.C.wcap) : k1_slab_bytes(E.C.slab_cols, E.C.pool_cap, E.C.win_cap, E.C.wcap)) * K1_WARPS;
    { cudaError_t e = cudaFuncSetAttribute(k_chain_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    int per_sm = 1; cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_finish, K1_WARPS * 32, smem);
    if (e != cudaSuccess) return e; if (per_sm < 1) per_sm = 1;
    long long want = ((long long)E.n_pending + K1_WARPS - 1) / K1_WARPS;
    int grid = (int)std::min<long long>(want, (long long)n_sm * per_sm); if (grid < 1) grid = 1;
    k_chain_finish<<<grid, K1_WARPS * 32, smem, stream>>>(E);
    return cudaGetLastError();
}

// shared memory of a warp's pair slab; beyond PAIR_SMEM_SLAB_MAX per warp the slabs live in HBM (PairParams::gslab, long-read mode with thousands of columns)
constexpr size_t PAIR_SMEM_SLAB_MAX = 100 * 1024;
size_t pair_gslab_bytes(int maxcol, int tier) { size_t b = tier == 0 ? pair_slab_bytes<PairTier0>(maxcol) : pair_slab_bytes<PairTier1>(maxcol); return b > PAIR_SMEM_SLAB_MAX ? b : 0; }
template <class CFG, bool FROM_LIST, bool UNPAIRED> static cudaError_t launch_pair_tier(const PairParams& P, int n_sm, long long want_warps, cudaStream_t stream) {
    // the slab of a warp grows with max_columns: as many warps per CTA (<= K3_WARPS) as fit the shared memory of an SM
    const bool in_hbm = P.gslab[FROM_LIST ? 1 : 0] != nullptr;
    int warps = K3_WARPS; while (!in_hbm && warps > 1 && pair_slab_bytes<CFG>(P.maxcol) * (size_t)warps > 200 * 1024) warps /= 2;
    const size_t smem = in_hbm ? 0 : pair_slab_bytes<CFG>(P.maxcol) * (size_t)warps;
    cudaError_t e = cudaFuncSetAttribute(k_pair<CFG, FROM_LIST, UNPAIRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
    int per_sm = 1; e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pair<CFG, FROM_LIST, UNPAIRED>, warps * 32, smem);
    if (e != cudaSuccess) return e; if (per_sm < 1) per_sm = 1;
    long long want = (want_warps + warps - 1) / warps;
    long long cap = (long long)n_sm * per_sm; if (in_hbm) cap = std::min<long long>(cap, (long long)(P.gslab_warps / (unsigned long long)warps));
    int grid = (int)std::min<long long>(want, cap); if (grid < 1) grid = 1;
    k_pair<CFG, FROM_LIST, UNPAIRED><<<grid, warps * 32, smem, stream>>>(P);
    return cudaGetLastError();
}
// tier 0 over all pairs of the wave, then tier 1 (64 chains per read, 4096 combinations) over the pairs tier 0 queued; P.defer_count must be zero on entry
cudaError_t launch_pair(const PairParams& P, int n_sm, cudaStream_t stream) {
    long long n_pairs = P.pair_end - P.pair_begin;
    if (n_pairs <= 0) return cudaSuccess;
    if (P.unpaired) {
        cudaError_t e = launch_pair_tier<PairTier0, false, true>(P, n_sm, n_pairs, stream); if (e != cudaSuccess) return e;
        if (!P.defer_list) return cudaSuccess;
        return launch_pair_tier<PairTier1, true, true>(P, n_sm, (long long)n_sm * K3_WARPS, stream);
    }
    cudaError_t e = launch_pair_tier<PairTier0, false, false>(P, n_sm, n_pairs, stream); if (e != cudaSuccess) return e;
    if (!P.defer_list) return cudaSuccess;
    return launch_pair_tier<PairTier1, true, false>(P, n_sm, (long long)n_sm * K3_WARPS, stream);     // a handful of pairs at most: one CTA per SM pops them
}

size_t dp_thread_scratch_bytes() { return dp_scratch_bytes(); }
int dp_ext_cap() { return DP_EXT_CAP; }

__global__ void k_exp_probe(const double* x, double* y, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = exp_like_host_libm(x[i]); }
cudaError_t launch_exp_probe(const double* x, double* y, int n, cudaStream_t stream) {
    if (n > 0) k_exp_probe<<<(n + 255) / 256, 256, 0, stream>>>(x, y, n);
    return cudaGetLastError();
}

cudaError_t launch_export_chain_columns(const DevGraph& G, int n_chains, int maxcol, const int32_t* n_cols, const int32_t* first_level,
                                        const int32_t* c_edge, int32_t* out_level, int32_t* out_edge_ord, uint8_t* out_gchar, cudaStream_t stream) {
    if (n_chains <= 0) return cudaSuccess;
    int threads = 128; long long blocks = ((long long)n_chains * 32 + threads - 1) / threads;
    k_export_chain_columns<<<(unsigned)blocks, threads, 0, stream>>>(G, n_chains, maxcol, n_cols, first_level, c_edge, out_level, out_edge_ord, out_gchar);
    return cudaGetLastError();
}

} // namespace hlala
