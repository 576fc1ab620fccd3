// Kernel 3: one warp per read pair. Chain de-duplication, chain-pair likelihoods with the insert-size model, first-maximum
// selection, mapping qualities (pair, per chain, per alignment column), per-level coverage.
//
// Restates processBAM::alignOneReadPair (mapper/processBAM.cpp:3188-3550, the part after the per-chain work),
// alignerBase::alignedReadPair_strandsValid / _pairsDistancesUnderlyingSequences (mapper/aligner/alignerBase.cpp:213-329),
// verboseSeedChain::alignment_{begin,end}_originalSequenceAnchors (mapper/reads/verboseSeedChain.h:231-283),
// processBAM::assignMappingQualities (mapper/processBAM.cpp:4062-4312) and the per-level counting of
// alignReads_postSeedExtraction_andStoreInto (mapper/processBAM.cpp:2411-2426).
//
// The reference keys its per-column posterior map with strings "gchar:level:rN:strand:readIndex"; within one read the
// strand and read number are constant and readIndex is a bijection of "how many read bases precede the column", so two
// columns of two chains share a key iff (graph char, level, preceding-base count or "no base") agree. Floating-point
// sums run in the reference's order (combination index ascending) so that they round identically.
#pragma once
#include "chain_params.h"
#include "exp_libm.cuh"
#include <cuda_runtime.h>

namespace hlala {

// Two capacity tiers (the reference has no limit): tier 0 holds 32 kept chains per read and 512 combinations per pair in 8.6 KB of shared memory per
// warp and serves practically every pair; pairs beyond it (reads with tens of distinct-coordinate chains inside HLA paralogs) are queued and re-run
// by tier 1 with 64 chains per read (64-bit membership masks) and 4096 combinations. Beyond tier 1 a pair is an error (HLALA_E_CAPACITY).
template <int KCAP_, int COMBO_, class MASK_> struct PairCfg { static constexpr int KCAP = KCAP_, COMBO = COMBO_; typedef MASK_ Mask; };
typedef PairCfg<K3_KCAP, K3_COMBO_CAP, uint32_t> PairTier0;
typedef PairCfg<64, 4096, unsigned long long> PairTier1;

template <class CFG> struct PairSlab { double* ll; double* Q; typename CFG::Mask* mask; int32_t* blevel; uint8_t* bg; uint8_t* linfo; int32_t* kept1; int32_t* kept2; };
template <class CFG> __host__ __device__ inline size_t pair_slab_bytes(int maxcol) {
    size_t b = (size_t)CFG::COMBO * 8 + (size_t)maxcol * 8 + (size_t)maxcol * sizeof(typename CFG::Mask) + (size_t)maxcol * 4 + (size_t)CFG::KCAP * 4 * 2 + (size_t)maxcol * 2;
    return (b + 15) & ~size_t(15);
}
template <class CFG> __device__ inline PairSlab<CFG> carve_pair_slab(unsigned char* p, int maxcol) {
    PairSlab<CFG> s;
    s.ll = (double*)p; p += (size_t)CFG::COMBO * 8;
    s.Q = (double*)p; p += (size_t)maxcol * 8;
    s.mask = (typename CFG::Mask*)p; p += (size_t)maxcol * sizeof(typename CFG::Mask);
    s.blevel = (int32_t*)p; p += (size_t)maxcol * 4;
    s.kept1 = (int32_t*)p; p += CFG::KCAP * 4; s.kept2 = (int32_t*)p; p += CFG::KCAP * 4;
    s.bg = p; p += maxcol; s.linfo = p; p += maxcol;
    return s;
}

__device__ __forceinline__ double is_value(const PairParams& P, int d) { int k = d - P.is_dmin; return (k >= 0 && k < P.is_n) ? P.is_table[k] : P.is_penalty; }

__device__ inline bool anchor_lookup(const DevGraph& G, int level, int id, int& pos) {
    for (int k = G.anchor_off[level]; k < G.anchor_off[level + 1]; k++) { int a = G.anchor_prg_id[k]; if (a == id) { pos = G.anchor_pos[k]; return true; } if (a > id) break; }
    return false;
}

// max over shared underlying sequences of log N(distance) between the upstream chain's end and the downstream chain's begin
__device__ double insert_size_ll(const PairParams& P, int up_last, int up_nlev, int dn_first, int dn_nlev) {
    const DevGraph& G = P.g;
    bool any = false; double best = 0;
    const int na = up_nlev < 2 ? up_nlev : 2, nb = dn_nlev < 2 ? dn_nlev : 2;
    for (int ai = 0; ai < na; ai++) {
        const int la = up_last - ai;
        for (int k = G.anchor_off[la]; k < G.anchor_off[la + 1]; k++) {
            const int id = G.anchor_prg_id[k]; int posA = G.anchor_pos[k], tmp;
            if (ai == 1 && anchor_lookup(G, up_last, id, tmp)) continue;     // the level nearer to the end already supplied this sequence
            int posB; bool have = anchor_lookup(G, dn_first, id, posB);
            if (!have && nb > 1) have = anchor_lookup(G, dn_first + 1, id, posB);
            if (!have) continue;
            double v = is_value(P, posB - posA - 1);
            if (!any || v > best) best = v;
            any = true;
        }
    }
    return any ? best : P.is_penalty;
}

__device__ __forceinline__ uint8_t phred_char(const PairParams& P, double q) {
    int lo = 0, hi = 222;                       // largest k with q >= thr[k]; thr[0] == 0
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (q >= P.phred_thr[mid]) lo = mid; else hi = mid - 1; }
    return (uint8_t)(33 + lo);
}

// per-read helper: kept chains after the strand and identical-coordinates rules (processBAM.cpp:3216, 3234)
template <int KCAP> __device__ int gather_kept(const PairParams& P, int r, int32_t* kept, int& err) {
    int n = 0;
    for (int s = P.b.chain_off[r]; s < P.b.chain_off[r + 1]; s++) {
        int st = P.status[s];
        if (st == CH_SKIPPED_STRAND || st == CH_DUPLICATE) continue;
        if (st != CH_OK) { err = st < 0 ? st : HLALA_E_INVARIANT_DEV; return n; }
        bool dup = false;
        for (int k = 0; k < n; k++) if (P.id_first[kept[k]] == P.id_first[s] && P.id_last[kept[k]] == P.id_last[s]) { dup = true; break; }
        if (dup) continue;        // chains are AS-sorted, so the registered score of an equal id is always >= this one
        if (n >= KCAP) { err = HLALA_E_CAPACITY_DEV; return n; }
        kept[n++] = s;
    }
    return n;
}

// membership bit `bit` for every column of the chosen chain `cs` against chain `as` (both of the same read)
template <class CFG> __device__ void mark_members(const PairParams& P, const PairSlab<CFG>& S, int cs, int as, int bit, int lane) {
    const int mc = P.maxcol;
    const int na = P.n_cols[as], fa = P.first_level[as], nlev_a = P.last_level[as] - fa + 1;
    const int32_t* ae = P.c_edge + (size_t)(as - P.slot_base) * mc; const uint8_t* asq = P.c_schar + (size_t)(as - P.slot_base) * mc;
    for (int i = lane; i < nlev_a && i < mc; i += 32) S.linfo[i] = 0;
    __syncwarp();
    int cl = 0, cb = 0;
    for (int base = 0; base < na; base += 32) {
        int k = base + lane; bool in = k < na;
        int32_t e = in ? ae[k] : -1; uint8_t sc = in ? asq[k] : (uint8_t)'_';
        unsigned ml = __ballot_sync(0xffffffffu, in && e >= 0), mb = __ballot_sync(0xffffffffu, in && sc != '_');
        if (in) {
            int lvl = e >= 0 ? fa + cl + __popc(ml & ((1u << lane) - 1)) : -1;
            uint8_t g = e >= 0 ? (uint8_t)(P.g.edge_pack[e] >> 16) : (uint8_t)'_';
            if (sc != '_') { int bi = cb + __popc(mb & ((1u << lane) - 1)); S.blevel[bi] = lvl; S.bg[bi] = g; }
            else if (e >= 0) S.linfo[lvl - fa] = g | 0x80;
        }
        cl += __popc(ml); cb += __popc(mb);
    }
    __syncwarp();
    const int nc = P.n_cols[cs], fc = P.first_level[cs];
    const int32_t* ce = P.c_edge + (size_t)(cs - P.slot_base) * mc; const uint8_t* csq = P.c_schar + (size_t)(cs - P.slot_base) * mc;
    cl = 0; cb = 0;
    for (int base = 0; base < nc; base += 32) {
        int k = base + lane; bool in = k < nc;
        int32_t e = in ? ce[k] : -1; uint8_t sc = in ? csq[k] : (uint8_t)'_';
        unsigned ml = __ballot_sync(0xffffffffu, in && e >= 0), mb = __ballot_sync(0xffffffffu, in && sc != '_');
        if (in) {
            int lvl = e >= 0 ? fc + cl + __popc(ml & ((1u << lane) - 1)) : -1;
            uint8_t g = e >= 0 ? (uint8_t)(P.g.edge_pack[e] >> 16) : (uint8_t)'_';
            bool mem;
            if (sc != '_') { int bi = cb + __popc(mb & ((1u << lane) - 1)); mem = (S.blevel[bi] == lvl && S.bg[bi] == g); }
            else { int kk = lvl - fa; mem = (e >= 0 && kk >= 0 && kk < nlev_a && S.linfo[kk] == (uint8_t)(g | 0x80)); }
            if (mem) S.mask[k] |= ((typename CFG::Mask)1 << bit);
        }
        cl += __popc(ml); cb += __popc(mb);
    }
    __syncwarp();
}

// UNPAIRED (long-read mode): the unit is one read (stored as a pair whose second mate is empty): processBAM::alignOneLongRead's first-maximum chain and
// assignMappingQualities_unpaired (processBAM.cpp:3752-3758, 3900-4059) — the posterior over the read's kept chains, no insert-size term, the same per-column
// confidences. It is the paired arithmetic with one (virtual) chain of log-likelihood 0 on the other side.
template <class CFG, bool FROM_LIST, bool UNPAIRED> __global__ void __launch_bounds__(K3_WARPS * 32) k_pair(PairParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename CFG::Mask Mask;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    unsigned char* const gs = P.gslab[FROM_LIST ? 1 : 0];
    if (gs && (unsigned long long)gridDim.x * warps_per_cta > P.gslab_warps) return;     // cannot happen: the launcher sizes the grid to the slabs
    PairSlab<CFG> S = carve_pair_slab<CFG>(gs ? gs + ((size_t)blockIdx.x * warps_per_cta + warp) * pair_slab_bytes<CFG>(P.maxcol) : smem + (size_t)warp * pair_slab_bytes<CFG>(P.maxcol), P.maxcol);
    const DevBatch& B = P.b; const int mc = P.maxcol;
    const int nw = gridDim.x * warps_per_cta;
    const long long n_work = FROM_LIST ? (long long)*P.defer_count : P.pair_end - P.pair_begin;
    for (long long wi = (long long)blockIdx.x * warps_per_cta + warp; wi < n_work; wi += nw) {
        const long long p = FROM_LIST ? (long long)P.defer_list[wi] : P.pair_begin + wi;
        const int r1 = (int)(2 * p), r2 = r1 + 1;
        int n1 = 0, n2 = 0, err = 0;
        if (lane == 0) { n1 = gather_kept<CFG::KCAP>(P, r1, S.kept1, err); if (UNPAIRED) n2 = 1; else if (!err) n2 = gather_kept<CFG::KCAP>(P, r2, S.kept2, err); if (!err && (n1 == 0 || n2 == 0)) err = HLALA_E_INVARIANT_DEV; if (!err && n1 * n2 > CFG::COMBO) err = HLALA_E_CAPACITY_DEV; }
        n1 = __shfl_sync(0xffffffffu, n1, 0); n2 = __shfl_sync(0xffffffffu, n2, 0); err = __shfl_sync(0xffffffffu, err, 0);
        __syncwarp();
        if (!FROM_LIST && err == HLALA_E_CAPACITY_DEV && P.defer_list) {     // re-run by the large tier
            if (lane == 0) P.defer_list[atomicAdd(P.defer_count, 1)] = (int32_t)p;
            continue;
        }
        if (err) {
            if (lane == 0) { P.pair_status[p] = err; P.pair_mapq[p] = -1; P.chosen_slot[r1] = -1; P.chosen_slot[r2] = -1; if (P.out_n_cols) { P.out_n_cols[r1] = 0; P.out_n_cols[r2] = 0; } atomicAdd(P.error_count, 1); }
            continue;
        }
        const bool rev1 = (B.chain_flag[B.chain_order[S.kept1[0]]] & 0x10) != 0, rev2 = UNPAIRED ? false : (B.chain_flag[B.chain_order[S.kept2[0]]] & 0x10) != 0;
        const int nc = n1 * n2;
        // ---- combination likelihoods (processBAM.cpp:3408-3506)
        for (int i = lane; i < nc; i += 32) {
            if (UNPAIRED) { S.ll[i] = P.ll[S.kept1[i]]; continue; }
            const int a = S.kept1[i / n2], b = S.kept2[i % n2];
            double ll = P.ll[a] + P.ll[b];
            const int f1 = P.first_level[a], l1 = P.last_level[a], f2 = P.first_level[b], l2 = P.last_level[b];
            bool valid = false;
            if (f1 != -1 && f2 != -1 && rev1 != rev2) valid = !rev1 ? (f1 < f2) : (l1 > l2);
            double lis = P.is_penalty;
            if (valid) lis = (f1 < f2) ? insert_size_ll(P, l1, l1 - f1 + 1, f2, l2 - f2 + 1) : insert_size_ll(P, l2, l2 - f2 + 1, f1, l1 - f1 + 1);
            ll += lis;
            S.ll[i] = ll;
        }
        __syncwarp();
        // ---- first maximum (Utilities::findVectorMax)
        double best = 0; int besti = -1;
        for (int i = lane; i < nc; i += 32) { double v = S.ll[i]; if (besti < 0 || v > best) { best = v; besti = i; } }
        for (int d = 16; d; d >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, d); int oi = __shfl_xor_sync(0xffffffffu, besti, d);
            if (oi >= 0 && (besti < 0 || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
        }
        const int ia = besti / n2, ib = besti % n2;
        const int sa = S.kept1[ia], sb = UNPAIRED ? -1 : S.kept2[ib];
        double mapq = 1, mq1 = 1, mq2 = 1;
        if (nc > 1) {
            // ---- posteriors (processBAM.cpp:4070-4123)
            for (int i = lane; i < nc; i += 32) S.ll[i] = exp_like_host_libm(S.ll[i] - best);   // the host libm's exp bit for bit (exp_libm.cuh): these terms decide between phred 255 and 190
            __syncwarp();
            double sum = 0;
            if (lane == 0) { for (int i = 0; i < nc; i++) sum += S.ll[i]; }
            sum = __shfl_sync(0xffffffffu, sum, 0);
            for (int i = lane; i < nc; i += 32) S.ll[i] = S.ll[i] / sum;
            __syncwarp();
            if (lane == 0) {
                mapq = S.ll[besti]; mq1 = 0; mq2 = 0;
                for (int i = 0; i < nc; i++) { double pp = S.ll[i]; if (i / n2 == ia) mq1 += pp; if (i % n2 == ib) mq2 += pp; }
                if (mq1 > 1) mq1 = 1; if (mq2 > 1) mq2 = 1;
            }
            mapq = __shfl_sync(0xffffffffu, mapq, 0); mq1 = __shfl_sync(0xffffffffu, mq1, 0); mq2 = __shfl_sync(0xffffffffu, mq2, 0);
        }
        if (lane == 0) {
            P.pair_status[p] = 0; P.pair_mapq[p] = mapq; P.pair_ll[p] = best; P.read_mapq[r1] = mq1; P.read_mapq[r2] = UNPAIRED ? 0.0 : mq2;
            P.read_reverse[r1] = rev1; P.read_reverse[r2] = rev2; P.chosen_slot[r1] = sa; P.chosen_slot[r2] = sb;
            if (UNPAIRED && P.out_n_cols) P.out_n_cols[r2] = 0;
            if (mapq < 1) atomicAdd(P.digest + 2, 1ull);
        }
        // ---- per read: per-column confidences, outputs, coverage
        for (int which = 0; which < (UNPAIRED ? 1 : 2); which++) {
            const int cs = which == 0 ? sa : sb; const int r = which == 0 ? r1 : r2;
            const int ncol = P.n_cols[cs]; const int fl = P.first_level[cs];
            if (nc > 1) {
                for (int k = lane; k < ncol; k += 32) S.mask[k] = 0;
                __syncwarp();
                const int nk = which == 0 ? n1 : n2; const int32_t* kept = which == 0 ? S.kept1 : S.kept2;
                for (int t = 0; t < nk; t++) mark_members<CFG>(P, S, cs, kept[t], t, lane);
                for (int k = lane; k < ncol; k += 32) {
                    const Mask m = S.mask[k]; double q = 0;
                    if (which == 0) { for (int a = 0; a < n1; a++) if ((m >> a) & 1) for (int b = 0; b < n2; b++) q += S.ll[a * n2 + b]; }
                    else { for (int a = 0; a < n1; a++) for (int b = 0; b < n2; b++) if ((m >> b) & 1) q += S.ll[a * n2 + b]; }
                    if (q > 1) q = 1;
                    S.Q[k] = q;
                }
                __syncwarp();
            }
            const size_t cb0 = (size_t)(cs - P.slot_base) * mc;
            const int32_t* ce = P.c_edge + cb0; const uint8_t* csq = P.c_schar + cb0; const uint8_t* cfs = P.c_fromseed + cb0;
            const size_t oo = (size_t)r * mc;
            int cl = 0; unsigned long long dsum = 0;
            for (int base = 0; base < ncol; base += 32) {
                int k = base + lane; bool in = k < ncol;
                int32_t e = in ? ce[k] : -1;
                unsigned ml = __ballot_sync(0xffffffffu, in && e >= 0);
                if (in) {
                    int lvl = e >= 0 ? fl + cl + __popc(ml & ((1u << lane) - 1)) : -1;
                    uint8_t g = e >= 0 ? (uint8_t)(P.g.edge_pack[e] >> 16) : (uint8_t)'_';
                    uint8_t mq = (nc > 1) ? phred_char(P, S.Q[k]) : (uint8_t)255;     // PCorrectToPhred(1) == 255
                    int eo = e >= 0 ? P.g.edge_ord[e] : -1;
                    if (P.out_level) { P.out_level[oo + k] = lvl; P.out_edge[oo + k] = eo; P.out_gchar[oo + k] = g; P.out_schar[oo + k] = csq[k]; P.out_fromseed[oo + k] = cfs[k]; }
                    if (P.out_mapq) P.out_mapq[oo + k] = mq;
                    if (P.bases_per_level && lvl != -1 && g != '_') atomicAdd(P.bases_per_level + lvl, 1);
                    dsum += (unsigned long long)(eo + 1);
                }
                cl += __popc(ml);
            }
            for (int d = 16; d; d >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, d);
            if (lane == 0) { if (P.out_n_cols) P.out_n_cols[r] = ncol; atomicAdd(P.digest + 0, (unsigned long long)ncol); atomicAdd(P.digest + 1, dsum); }
            __syncwarp();
        }
    }
}

} // namespace hlala
