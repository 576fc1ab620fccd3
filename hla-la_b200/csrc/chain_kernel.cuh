// Kernel 1: one warp per chain. BAM record -> graph-space columns -> column Viterbi over the level DAG ->
// backtrace -> (extension) -> padding -> alignment log-likelihood.
//
// Restates, per chain, the reference sequence
//   processBAM::transformBAMreadToInternalAlignment   mapper/processBAM.cpp:4794-5337   (expand_cigar)
//   processBAM::PRGContigAlignment2Seed                mapper/processBAM.cpp:2491-3017   (trim, fill, viterbi, backtrace)
//   processBAM::cleanInitialAlignment                  mapper/processBAM.cpp:4621-4792   (clean_columns)
//   processBAM::restrictInitialAlignmentToNoGapAreas   mapper/processBAM.cpp:4461-4619   (restrict_columns)
//   extensionAligner::extendSeedChain                  mapper/aligner/extensionAligner.cpp:186-333 (extension hook + padding)
//   verboseSeedChain::extendToFullSequenceLength       mapper/reads/verboseSeedChain.cpp:82-144
//   extensionAligner::scoreOneAlignment                mapper/aligner/extensionAligner.cpp:52-182
// as a B200 design: columns live in shared memory (one slab per warp), the edges of the chain's level window (a contiguous range of the
// flat edge array) are brought into shared memory by ONE bulk copy of the TMA engine (cp.async.bulk + mbarrier, issued by lane 0 and
// overlapped with the computation of the per-level offsets), the per-column Viterbi step is a shared-memory
// atomicMax over packed (score, canonical edge rank) keys so that "keep all arg-max incoming edges, take the first in
// set<Edge*> order" becomes a single integer max, and the backtrace never leaves shared memory.
#pragma once
#include "chain_params.h"
#include <cuda_runtime.h>

namespace hlala {

__constant__ ScoreTables c_tables;
__constant__ ScoreTables c_tables_long;     // long-read mode: indel rates 0.075 (extensionAligner.cpp:58-64)

struct WarpSlab {
    int32_t* lvlA; int32_t* lvlB; uint8_t* gA; uint8_t* sA; uint8_t* gB; uint8_t* sB;
    uint32_t* bt; uint32_t* win; uint16_t* weoff; uint16_t* wwid; uint16_t* coloff; uint32_t* cur; uint32_t* nxt;
    unsigned long long* mbar; uint32_t* mbar_phase;     // transaction barrier of the window's bulk copy and its phase bit
};

// ---- TMA bulk copy (1-D, global -> shared, completion on an mbarrier): the PTX the SASS shows as UBLKCP / SYNCS
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy accesses to dst are ordered before the async-proxy writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_addr(bar)), "r"(phase) : "memory");
}

__device__ inline WarpSlab carve_slab(unsigned char* base, int maxcol, int pool_cap, int win_cap, int wcap) {
    WarpSlab s; unsigned char* p = base;
    s.win = (uint32_t*)p; p += (size_t)(win_cap + 4) * 4;      // first: the slab is 16-byte aligned, which the bulk copy needs; 3 words of slack for the aligned start, win_cap % 4 == 0
    s.mbar = (unsigned long long*)p; p += 8; s.mbar_phase = (uint32_t*)p; p += 8;
    s.lvlA = (int32_t*)p; p += (size_t)maxcol * 4;
    s.lvlB = (int32_t*)p; p += (size_t)maxcol * 4;
    s.bt = (uint32_t*)p; p += (size_t)pool_cap * 4;
    s.weoff = (uint16_t*)p; p += (size_t)(maxcol + 4) * 2; s.wwid = (uint16_t*)p; p += (size_t)(maxcol + 4) * 2;
    s.cur = (uint32_t*)p; p += (size_t)wcap * 4;
    s.nxt = (uint32_t*)p; p += (size_t)wcap * 4;
    s.coloff = (uint16_t*)p; p += (size_t)((maxcol + 1) / 2 * 2) * 2;
    s.gA = p; p += maxcol; s.sA = p; p += maxcol; s.gB = p; p += maxcol; s.sB = p; p += maxcol;
    return s;
}

// long chains: window, node-score rows and barrier in shared memory, everything else in this warp's slice of HBM
__device__ inline WarpSlab carve_slab_split(unsigned char* sm, unsigned char* gm, int maxcol, int pool_cap, int win_cap, int wcap) {
    WarpSlab s; unsigned char* p = sm;
    s.win = (uint32_t*)p; p += (size_t)(win_cap + 4) * 4;
    s.mbar = (unsigned long long*)p; p += 8; s.mbar_phase = (uint32_t*)p; p += 8;
    s.cur = (uint32_t*)p; p += (size_t)wcap * 4; s.nxt = (uint32_t*)p; p += (size_t)wcap * 4;
    p = gm;
    s.lvlA = (int32_t*)p; p += (size_t)maxcol * 4;
    s.lvlB = (int32_t*)p; p += (size_t)maxcol * 4;
    s.bt = (uint32_t*)p; p += (size_t)pool_cap * 4;
    s.weoff = (uint16_t*)p; p += (size_t)(maxcol + 4) * 2; s.wwid = (uint16_t*)p; p += (size_t)(maxcol + 4) * 2;
    s.coloff = (uint16_t*)p; p += (size_t)((maxcol + 1) / 2 * 2) * 2;
    s.gA = p; p += maxcol; s.sA = p; p += maxcol; s.gB = p; p += maxcol; s.sB = p; p += maxcol;
    return s;
}
__device__ inline WarpSlab slab_of(const ChainParams& P, unsigned char* smem, int warp, int warps_per_cta) {
    if (P.gslab) {
        const size_t gw = (size_t)blockIdx.x * warps_per_cta + warp;
        return carve_slab_split(smem + (size_t)warp * k1_gslab_smem_bytes(P.win_cap, P.wcap), P.gslab_base + gw * P.gslab_bytes, P.slab_cols, P.pool_cap, P.win_cap, P.wcap);
    }
    return carve_slab(smem + (size_t)warp * k1_slab_bytes(P.slab_cols, P.pool_cap, P.win_cap, P.wcap), P.slab_cols, P.pool_cap, P.win_cap, P.wcap);
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}
__device__ __forceinline__ int warp_incl_maxscan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
    return v;
}

struct ColBuf { int32_t* lvl; uint8_t* g; uint8_t* s; };

// --- transformBAMreadToInternalAlignment: CIGAR -> (level, graph char, read char) columns. Returns column count or <0.
__device__ int expand_cigar(const ChainParams& P, int c, int64_t rd0, int rdlen, ColBuf o, int lane, int& start_raw, int& stop_raw, int& id_first, int& id_last) {
    const DevGraph& G = P.g; const DevBatch& B = P.b;
    int contig = B.chain_contig[c]; int gpos = B.chain_pos[c];
    int cg0 = B.cigar_off[c], ncg = B.cigar_off[c + 1] - cg0;
    if (contig < 0 || contig >= G.n_contigs || ncg < 1) return HLALA_E_ARG_DEV;
    int64_t cbase = G.contig_off[contig]; int clen = (int)(G.contig_off[contig + 1] - cbase);
    int ridx = 0, ncol = 0; start_raw = -1; stop_raw = -1;
    const int pos0 = gpos;
    for (int k = 0; k < ncg; k++) {
        uint32_t cg = B.cigar[cg0 + k]; int op = cg & 15; int len = (int)(cg >> 4);
        if (op == 0 || op == 7 || op == 8 || op == 2) {            // M = X consume both; D consumes the reference only
            bool isD = (op == 2);
            if (gpos < 0 || gpos + len > clen) return HLALA_E_INVARIANT_DEV;
            if (!isD && ridx + len > rdlen) return HLALA_E_INVARIANT_DEV;
            if (ncol + len > P.slab_cols) return HLALA_E_CAPACITY_DEV;
            for (int i = lane; i < len; i += 32) {
                o.lvl[ncol + i] = G.contig_level[cbase + gpos + i];
                o.g[ncol + i] = G.contig_seq[cbase + gpos + i];
                o.s[ncol + i] = isD ? (uint8_t)'_' : B.bases[rd0 + ridx + i];
            }
            if (start_raw < 0) start_raw = ridx;
            ncol += len; gpos += len; if (!isD) ridx += len;
            stop_raw = ridx - 1;
        } else if (op == 1) {                                       // I: read bases against a graph gap, level -1
            if (ridx + len > rdlen) return HLALA_E_INVARIANT_DEV;
            if (ncol + len > P.slab_cols) return HLALA_E_CAPACITY_DEV;
            for (int i = lane; i < len; i += 32) { o.lvl[ncol + i] = -1; o.g[ncol + i] = '_'; o.s[ncol + i] = B.bases[rd0 + ridx + i]; }
            if (start_raw < 0) start_raw = ridx;
            ncol += len; ridx += len; stop_raw = ridx - 1;
        } else if (op == 4) { ridx += len; }                        // S
        else if (op == 5) { if (k == 0) ridx += len; }              // H: a leading hard clip offsets into the primary's SEQ (processBAM.cpp:4868-4874)
        else if (op == 6) { }                                        // P: dropped before the column walk (processBAM.cpp:4817)
        else return HLALA_E_INVARIANT_DEV;                          // N: the reference throws (processBAM.cpp:5167)
    }
    // alignment_get_startstop_PRGcoordinates (processBAM.cpp:3840): levels of Position and GetEndPosition(false, true)
    if (pos0 < 0 || gpos - 1 >= clen || gpos - 1 < pos0) return HLALA_E_INVARIANT_DEV;
    id_first = G.contig_level[cbase + pos0]; id_last = G.contig_level[cbase + gpos - 1];
    __syncwarp();
    return ncol;
}

// --- PRGContigAlignment2Seed part 1: drop leading/trailing insertion columns, insert '_'/'_' columns for skipped levels.
__device__ int trim_and_fill(const ChainParams& P, ColBuf in, int n, ColBuf out, int lane, int& start_raw, int& stop_raw) {
    // first / last column with a level
    int first = n, last = -1;
    for (int i = lane; i < n; i += 32) if (in.lvl[i] != -1) { first = min(first, i); last = max(last, i); }
    for (int d = 16; d; d >>= 1) { first = min(first, __shfl_xor_sync(0xffffffffu, first, d)); last = max(last, __shfl_xor_sync(0xffffffffu, last, d)); }
    if (!(first < last)) return HLALA_E_INVARIANT_DEV;              // assert(firstColumn < lastColumn), processBAM.cpp:2532
    start_raw += first; stop_raw -= (n - 1 - last);
    int carry_prev = -1;      // last level seen so far
    int carry_out = 0;        // output columns emitted so far
    for (int base = first; base <= last; base += 32) {
        int i = base + lane; bool in_range = i <= last;
        int lv = in_range ? in.lvl[i] : -1;
        int seen = warp_incl_maxscan(lv, lane);                     // levels increase along the alignment
        int prev_incl = max(seen, carry_prev);
        int prev_excl = __shfl_up_sync(0xffffffffu, prev_incl, 1); if (lane == 0) prev_excl = carry_prev;
        int fill = 0;
        if (in_range && lv != -1 && prev_excl != -1) { fill = lv - prev_excl - 1; if (fill < 0) fill = -1000000; }
        int bad = __any_sync(0xffffffffu, fill < 0);
        if (bad) return HLALA_E_INVARIANT_DEV;                      // assert((last+1) < level), processBAM.cpp:2563
        int width = in_range ? 1 + fill : 0;
        int incl = warp_incl_scan(width, lane);
        int dst_end = carry_out + incl;                             // one past this column's own slot
        int total = __shfl_sync(0xffffffffu, incl, 31);
        if (carry_out + total > P.slab_cols) return HLALA_E_CAPACITY_DEV;
        if (in_range) {
            int own = dst_end - 1;
            for (int k = 0; k < fill; k++) { int d = own - fill + k; out.lvl[d] = prev_excl + 1 + k; out.g[d] = '_'; out.s[d] = '_'; }
            out.lvl[own] = lv; out.g[own] = in.g[i]; out.s[own] = in.s[i];
        }
        carry_out += total;
        carry_prev = __shfl_sync(0xffffffffu, prev_incl, 31);
    }
    __syncwarp();
    return carry_out;
}

// --- cleanInitialAlignment (processBAM.cpp:4621-4792). Works in place on `c`; `tmp` is scratch. Returns new column count.
__device__ int clean_columns(ColBuf c, int n, ColBuf tmp, int lane) {
    int any_ins = 0;
    for (int i = lane; i < n; i += 32) any_ins |= (c.lvl[i] == -1);
    if (!__any_sync(0xffffffffu, any_ins)) return n;               // no insertion column: no stretch can balance
    int cleaning = 0;
    if (lane == 0) {
        bool inS = false; int sStart = -1, balance = 0;
        for (int p = 0; p < n; p++) {
            bool ins = (c.lvl[p] == -1); bool dgap = (c.g[p] == '_' && c.s[p] == '_');
            if (ins || dgap) {
                if (!inS) { sStart = p; inS = true; }
                if (ins) balance++;
                if (dgap) balance--;
            } else if (inS) {
                int sStop = p - 1;
                if (balance == 0) {
                    int L = sStop - sStart + 1, ni = 0, ng = 0;
                    for (int q = sStart; q <= sStop; q++) { if (c.lvl[q] == -1) tmp.s[sStart + ni++] = c.s[q]; else tmp.lvl[sStart + ng++] = c.lvl[q]; }
                    for (int q = sStart; q <= sStop; q++) {
                        int k = q - sStart;
                        if (k < L / 2) { c.lvl[q] = tmp.lvl[sStart + k]; c.g[q] = '_'; c.s[q] = tmp.s[sStart + k]; }
                        else { c.lvl[q] = -1; c.g[q] = '_'; c.s[q] = '_'; }
                    }
                    cleaning = 1;
                }
                inS = false; sStart = -1; balance = 0;
            }
        }
    }
    cleaning = __shfl_sync(0xffffffffu, cleaning, 0);
    __syncwarp();
    if (!cleaning) return n;
    // delete columns that are (level -1, '_', '_'): stable compaction through tmp
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane; bool keep = i < n && !(c.lvl[i] == -1 && c.g[i] == '_' && c.s[i] == '_');
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) { int d = carry + __popc(m & ((1u << lane) - 1)); tmp.lvl[d] = c.lvl[i]; tmp.g[d] = c.g[i]; tmp.s[d] = c.s[i]; }
        carry += __popc(m);
    }
    __syncwarp();
    for (int i = lane; i < carry; i += 32) { c.lvl[i] = tmp.lvl[i]; c.g[i] = tmp.g[i]; c.s[i] = tmp.s[i]; }
    __syncwarp();
    return carry;
}

// --- restrictInitialAlignmentToNoGapAreas (processBAM.cpp:4461-4619). In place; returns new column count.
__device__ int restrict_columns(const ChainParams& P, ColBuf c, int n, ColBuf tmp, int lane, int& start_raw, int& stop_raw) {
    int any = 0;
    for (int i = lane; i < n; i += 32) { int lv = c.lvl[i]; any |= (lv != -1 && P.g.gap_stretch[lv]); }
    if (!__any_sync(0xffffffffu, any)) return n;                   // one uninterrupted run from column 0: no candidate (processBAM.cpp:4492-4502)
    int selA = -1, selB = -1;
    if (lane == 0) {
        int run = -1; int bestLen = 0;
        auto consider = [&](int a, int b) {
            // trim insertion columns at both ends (processBAM.cpp:4511-4527)
            while (c.lvl[a] == -1) { a++; if (a > b || a > n - 1) break; }
            if (a <= b) { while (c.lvl[b] == -1) { b--; if (b < a || b < 0) break; } }
            if (b >= a) {
                int len = b - a + 1;
                // ascending std::sort by length, then .back(): a longest stretch; among equals the later one
                // (libstdc++ sorts <= 16 elements by insertion sort, which is stable)
                if (len >= bestLen) { bestLen = len; selA = a; selB = b; }
            }
        };
        for (int i = 0; i < n; i++) {
            int lv = c.lvl[i];
            if (lv != -1 && P.g.gap_stretch[lv]) { if (run != -1) { consider(run, i - 1); run = -1; } }
            else if (run == -1) run = i;
        }
        if (run != -1 && run != 0) consider(run, n - 1);
    }
    selA = __shfl_sync(0xffffffffu, selA, 0); selB = __shfl_sync(0xffffffffu, selB, 0);
    if (selA < 0) return n;
    int tot = 0, inside = 0, before = 0, after = 0;
    for (int i = lane; i < n; i += 32) { if (c.s[i] != '_') { tot++; if (i < selA) before++; else if (i > selB) after++; else inside++; } }
    for (int d = 16; d; d >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, d); inside += __shfl_xor_sync(0xffffffffu, inside, d); before += __shfl_xor_sync(0xffffffffu, before, d); after += __shfl_xor_sync(0xffffffffu, after, d); }
    if (!(((double)inside / (double)tot) > 0.3)) return n;          // keep the alignment unchanged (processBAM.cpp:4606-4617)
    int m = selB - selA + 1;
    for (int i = lane; i < m; i += 32) { tmp.lvl[i] = c.lvl[selA + i]; tmp.g[i] = c.g[selA + i]; tmp.s[i] = c.s[selA + i]; }
    __syncwarp();
    for (int i = lane; i < m; i += 32) { c.lvl[i] = tmp.lvl[i]; c.g[i] = tmp.g[i]; c.s[i] = tmp.s[i]; }
    __syncwarp();
    start_raw += before; stop_raw -= after;
    return m;
}

// --- column Viterbi ("sequence" pass, processBAM.cpp:2789-2832) + backtrace (2838-2932).
// On return edge_out[col] holds the flat edge id (or -1 for insertion columns) and c.g[col] the chosen edge's emission.
template <bool BT16>
__device__ int viterbi_backtrace(const ChainParams& P, ColBuf c, int n, const WarpSlab& S, int32_t* edge_out, int lane) {
    const DevGraph& G = P.g;
    if (c.lvl[0] == -1 || c.lvl[n - 1] == -1) return HLALA_E_INVARIANT_DEV;
    const int l_first = c.lvl[0], l_last = c.lvl[n - 1];
    const int nlev = l_last - l_first + 1;
    if (l_last + 1 >= G.n_levels) return HLALA_E_INVARIANT_DEV;
    if (nlev > P.slab_cols) return HLALA_E_CAPACITY_DEV;
    // packed Viterbi keys: (score + 1) << KS | (rank mask - edge rank inside its level); KS = P.key_shift (20 for short reads: scores up to 4094; smaller for
    // chains of thousands of columns, chosen from the graph's widest level). A chain whose score could leave the field is a capacity error, never a wrap-around.
    const int KS = P.key_shift; const uint32_t RM = (1u << KS) - 1u;
    if ((uint32_t)n + 2u >= (1u << (32 - KS))) return HLALA_E_CAPACITY_DEV;
    // stage the window: the packed edges of levels [l_first, l_first + nlev) are one contiguous range of edge_pack. Lane 0 hands the 16-byte
    // aligned superset of that range to the TMA engine; the per-level edge offsets (relative) and node counts are computed while it is in flight.
    const int e_base = G.level_edge_off[l_first];
    const int n_win = G.level_edge_off[l_first + nlev] - e_base;
    const int woff = e_base & 3;
    const bool staged = n_win + woff <= P.win_cap;
    const uint32_t* win = S.win + woff;
    __syncwarp();
    if (staged && lane == 0) bulk_copy_g2s(S.win, G.edge_pack + (e_base - woff), (uint32_t)(((n_win + woff) * 4 + 15) & ~15), S.mbar);
    int bad = 0;
    for (int i = lane; i <= nlev + 1; i += 32) {
        if (i <= nlev) { int rel = G.level_edge_off[l_first + i] - e_base; if (rel > 65535) bad = 1; S.weoff[i] = (uint16_t)rel; }
        int w = G.level_node_off[l_first + i + 1] - G.level_node_off[l_first + i]; if (w > P.wcap) bad = 1; S.wwid[i] = (uint16_t)w;
    }
    if (staged) { const uint32_t ph = *S.mbar_phase; mbar_wait(S.mbar, ph); __syncwarp(); if (lane == 0) *S.mbar_phase = ph ^ 1u; }
    if (__any_sync(0xffffffffu, bad)) return HLALA_E_CAPACITY_DEV;
    __syncwarp();
    // init: every node of the first level, score 0 (processBAM.cpp:2696-2701)
    const int w0 = S.wwid[0];
    uint32_t* cur = S.cur; uint32_t* nxt = S.nxt;
    uint16_t* bt16 = (uint16_t*)S.bt; const int pool_cap = BT16 ? P.pool_cap * 2 : P.pool_cap;
    for (int z = lane; z < w0; z += 32) cur[z] = (1u << KS) | RM;
    __syncwarp();
    int pool = 0; int lev = l_first; int status = 0;
    for (int col = 0; col < n; col++) {
        {   // ---- a run of columns whose levels have ONE node on either side (everything outside bubbles and gene blocks): one column per lane.
            // With a single node per level the step has no choice of node: the edge is the first one that emits the read's character (+1), else - if the
            // linear alignment has a mismatch here - the first edge (+0); the scores are a prefix sum. Anything else (several nodes, a level without a usable
            // edge, a capacity limit) ends the run and is taken by the general step below, which also reports the errors.
            const unsigned full = 0xffffffffu;
            const int j = col + lane; const bool in = j < n;
            const int lv = in ? c.lvl[j] : -2;
            const bool isl = in && lv != -1;
            const unsigned lm = __ballot_sync(full, isl);
            const int pre = __popc(lm & ((1u << lane) - 1u));
            bool ok = in; int rank = 0, inc = 0;
            if (isl) {
                const int li = lev - l_first + pre;
                ok = (lv == lev + pre) && S.wwid[li] == 1 && S.wwid[li + 1] == 1;
                if (ok) {
                    const int e0 = S.weoff[li], ne = (int)S.weoff[li + 1] - e0;
                    ok = ne >= 1 && ne <= 8;
                    if (ok) {
                        const uint8_t sc = c.s[j], gc = c.g[j]; int hit = -1;
                        for (int k = 0; k < ne; k++) { const uint32_t pk = staged ? win[e0 + k] : G.edge_pack[e_base + e0 + k]; if ((uint8_t)(pk >> 16) == sc) { hit = k; break; } }
                        if (hit >= 0) { rank = hit; inc = 1; } else if (sc != gc) { rank = 0; inc = 0; } else ok = false;
                    }
                }
            }
            const unsigned okm = __ballot_sync(full, ok);
            const int m = okm == full ? 32 : __ffs(~okm) - 1;      // leading run of columns this path can take
            const unsigned mm = m >= 32 ? full : ((1u << m) - 1u);
            const int cnt = __popc(lm & mm);
            if (m >= 4 && cnt == 0) { col += m - 1; continue; }     // insertion columns only
            const uint32_t k0 = cur[0];
            if (m >= 4 && k0 != 0 && pool + cnt <= pool_cap && pool + cnt - 1 <= 65535) {
                int sinc = (lane < m && isl) ? inc : 0;
                for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(full, sinc, d); if (lane >= d) sinc += t; }
                if (lane < m && isl) { const int pj = pool + pre; if (BT16) bt16[pj] = (uint16_t)rank; else S.bt[pj] = (uint32_t)rank; S.coloff[j] = (uint16_t)pj; }
                const int lastl = 31 - __clz(lm & mm);
                __syncwarp();
                if (lane == lastl) cur[0] = (((k0 >> KS) + (uint32_t)sinc) << KS) | (RM - (uint32_t)rank);
                pool += cnt; lev += cnt; col += m - 1;
                __syncwarp();
                continue;
            }
        }
        if (c.lvl[col] == -1) continue;
        if (c.lvl[col] != lev) { status = HLALA_E_INVARIANT_DEV; break; }     // level contiguity (processBAM.cpp:2649-2665)
        const int li = lev - l_first;
        const int e0 = S.weoff[li], e1 = S.weoff[li + 1];
        const int wn = S.wwid[li + 1];
        if ((BT16 && (e1 - e0) > 255) || (uint32_t)(e1 - e0) > RM) { status = HLALA_E_CAPACITY_DEV; break; }
        if (pool + wn > pool_cap || pool > 65535) { status = HLALA_E_CAPACITY_DEV; break; }
        const uint8_t sc = c.s[col], gc = c.g[col]; const bool isMatch = (sc == gc);
        if (wn == 1 && S.wwid[li] == 1 && (e1 - e0) <= 32) {
            // one node on either side (the common case outside variant sites): the step is a single warp reduction
            uint32_t key = 0; uint32_t pk = 0; const int e = e0 + lane;
            if (e < e1) {
                pk = staged ? win[e] : G.edge_pack[e_base + e];
                const uint32_t kf = cur[0]; const uint8_t em = (uint8_t)(pk >> 16);
                if (kf != 0 && !(isMatch && em != sc)) key = (((kf >> KS) + (em == sc ? 1u : 0u)) << KS) | (RM - (uint32_t)(e - e0));
            }
            const uint32_t best = __reduce_max_sync(0xffffffffu, key);
            if (lane == 0) {
                nxt[0] = best;
                const uint32_t rank = RM - (best & RM);
                if (BT16) bt16[pool] = (uint16_t)rank; else S.bt[pool] = rank;       // from node is node 0
                S.coloff[col] = (uint16_t)pool;
            }
            pool += 1; lev++;
            uint32_t* t = cur; cur = nxt; nxt = t;
            __syncwarp();
            continue;
        }
        for (int z = lane; z < wn; z += 32) nxt[z] = 0;
        __syncwarp();
        for (int e = e0 + lane; e < e1; e += 32) {
            uint32_t pk = staged ? win[e] : G.edge_pack[e_base + e];
            uint32_t kf = cur[pk & 255u]; uint8_t em = (uint8_t)(pk >> 16);
            if (kf != 0 && !(isMatch && em != sc)) {
                uint32_t key = (((kf >> KS) + (em == sc ? 1u : 0u)) << KS) | (RM - (uint32_t)(e - e0));
                atomicMax(&nxt[(pk >> 8) & 255u], key);
            }
        }
        __syncwarp();
        for (int e = e0 + lane; e < e1; e += 32) {
            uint32_t pk = staged ? win[e] : G.edge_pack[e_base + e];
            uint32_t kf = cur[pk & 255u]; uint8_t em = (uint8_t)(pk >> 16);
            if (kf != 0 && !(isMatch && em != sc)) {
                uint32_t key = (((kf >> KS) + (em == sc ? 1u : 0u)) << KS) | (RM - (uint32_t)(e - e0));
                uint32_t tz = (pk >> 8) & 255u;
                if (nxt[tz] == key) {                                    // the unique winner records (edge rank, from node)
                    if (BT16) bt16[pool + tz] = (uint16_t)((e - e0) | ((pk & 255u) << 8));
                    else S.bt[pool + tz] = (uint32_t)(e - e0) | ((pk & 255u) << 20);
                }
            }
        }
        if (lane == 0) S.coloff[col] = (uint16_t)pool;
        pool += wn; lev++;
        uint32_t* t = cur; cur = nxt; nxt = t;
        __syncwarp();
    }
    if (status) return status;
    // end nodes with maximal score; the first of them in canonical node order (processBAM.cpp:2856-2867)
    const int wl = S.wwid[nlev];
    uint32_t best = 0;
    for (int z = lane; z < wl; z += 32) { uint32_t k = cur[z]; if (k) { uint32_t v = ((k >> KS) << 12) | (uint32_t)(4095 - min(z, 4095)); best = max(best, v); } }
    for (int d = 16; d; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (best == 0) return HLALA_E_INVARIANT_DEV;                    // the reference asserts a non-empty column map
    {   // ---- backtrace (processBAM.cpp:2869-2932), 32 columns at a time from the right. The entry a column reads is addressed by the node it arrives at in the NEXT
        // level: wherever that level has one node the address is known at once, so those columns (80 % on a PRG) are independent; the others wait for the column to
        // their right, a few rounds per block. Blocks with long runs of several-node levels (gene blocks) are walked by one lane as before.
        const unsigned full = 0xffffffffu;
        int carry = 4095 - (int)(best & 4095u);       // node the path arrives at in the level after the rightmost unresolved column
        for (int base = ((n - 1) / 32) * 32; base >= 0; base -= 32) {
            const int j = base + lane; const bool in = j < n;
            const int lv = in ? c.lvl[j] : -1; const bool isl = in && lv != -1;
            if (in && !isl) { edge_out[j] = -1; c.g[j] = '_'; }
            const int li = isl ? lv - l_first : 0;
            const bool known = isl && S.wwid[li + 1] == 1;
            const unsigned lm = __ballot_sync(full, isl);
            if (lm == 0u) continue;
            const int n_unknown = __popc(lm & ~__ballot_sync(full, known));
            auto take = [&](int col, int lcol, int z, int& fz_out) {
                int rank, fz;
                if (BT16) { const uint16_t ent = bt16[S.coloff[col] + z]; rank = ent & 255; fz = ent >> 8; }
                else { const uint32_t ent = S.bt[S.coloff[col] + z]; rank = (int)(ent & KEY_RANK_MASK); fz = (int)(ent >> 20); }
                const int erel = S.weoff[lcol] + rank;
                const uint32_t pk = staged ? win[erel] : G.edge_pack[e_base + erel];
                edge_out[col] = e_base + erel; c.g[col] = (uint8_t)(pk >> 16);
                fz_out = fz;
            };
            if (n_unknown > 6) {      // one lane, right to left
                if (lane == 0) { int z = carry; for (int col = min(base + 31, n - 1); col >= base; col--) { if (c.lvl[col] == -1) continue; int fz; take(col, c.lvl[col] - l_first, z, fz); z = fz; } carry = z; }
                carry = __shfl_sync(full, carry, 0);
                __syncwarp();
                continue;
            }
            const unsigned higher = lane == 31 ? 0u : (lm & ~((2u << lane) - 1u));
            const int nb = higher ? __ffs(higher) - 1 : -1;           // the level column to the right inside this block; -1: the carry
            int z = known ? 0 : -1, fz = 0; bool done = !isl;
            for (;;) {
                if (isl && !done && z >= 0) { take(j, li, z, fz); done = true; }
                if (__all_sync(full, done)) break;
                const int nfz = __shfl_sync(full, fz, nb >= 0 ? nb : 0); const int ndone = __shfl_sync(full, (int)done, nb >= 0 ? nb : 0);
                if (isl && !done && z < 0) { if (nb < 0) z = carry; else if (ndone) z = nfz; }
            }
            carry = __shfl_sync(full, fz, __ffs(lm) - 1);              // the leftmost level column of this block hands its node on
        }
    }
    __syncwarp();
    return 0;
}

} // namespace hlala
