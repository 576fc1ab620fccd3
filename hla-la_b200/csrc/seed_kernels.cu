// k-mer seeding kernels (sm_100a). One warp per read, reads dealt dynamically to a persistent grid.
//
// k_seed_chains replays GraphAndEdgeIndex::findChains (Graph/GraphAndEdgeIndex.cpp:39-356) for every read:
//   * lanes look the read's k-mers up 32 positions at a time (queryIndex, :986): hash of the k bytes -> open-addressing table in HBM ->
//     byte compare against the stored k-mer (exact, no false positives);
//   * the chain bookkeeping is sequential in the read position, as in the reference: per base every running chain is extended by the
//     outgoing paths of its last node that end in a non-gap edge emitting the base (:131-236; the in-function forwardScan lambda
//     :134-187 is replayed pass by pass, back to front, because its order decides which continuation stays in place and which ones
//     are appended as new chains), chains without continuation are archived (:250-279), then the k-mer ending at the base opens a
//     chain for every index position no running chain represents (:307-340). Lanes work on different running chains; events whose
//     order is observable (archiving, branching) are applied in the reference's order (chains back to front).
//   nodes_jumpOverGaps is never filled in the reference (fillEdgeJumper has no caller), so the jump loops (:98-129,:150-155) are dead.
// Running chains keep their edge lists in a per-warp slab of HBM scratch; finished chains are copied to a global pool (atomic bump
// allocation) and put into (read, order) order by k_seed_order / k_seed_gather afterwards.
#include "seed_kernels.h"
#include "kmer_hash.h"

#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace hlala {

namespace {

struct RunEnt { int32_t begin, n, target, slot; };
template <int RC> struct SeedWarpT { RunEnt run[RC]; uint16_t free_slots[RC]; int32_t n_run, n_free; };

__device__ __forceinline__ I4 ldI4(const I4* p) { const int4 v = __ldg(reinterpret_cast<const int4*>(p)); I4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }

__device__ int kmer_lookup(const DevKmerIndex& ix, const uint8_t* s) {
    uint32_t h = (uint32_t)kmer_hash(s, ix.k) & ix.ht_mask;
    for (;;) {
        const int v = __ldg(ix.ht + h);
        if (v == 0) return -1;
        const uint8_t* kb = ix.kmer_bytes + (size_t)(v - 1) * ix.k; bool eq = true;
        for (int i = 0; i < ix.k; i++) if (__ldg(kb + i) != s[i]) { eq = false; break; }
        if (eq) return v - 1;
        h = (h + 1) & ix.ht_mask;
    }
}

// The reference's forwardScan lambda for every outgoing edge of `target`, keeping the paths whose last edge emits c. One lane.
// Compatible paths are written to the blocks behind the work area: [0] length, [1] node the path ends in, [2..] flat edges.
__device__ int scan_slow(const DevGraph& G, int target, uint8_t c, int32_t* ss) {
    int16_t rblk[SEED_SCAN_NB]; int16_t rlen[SEED_SCAN_NB];
    unsigned long long freeb = (SEED_SCAN_NB >= 64) ? ~0ull : ((1ull << SEED_SCAN_NB) - 1ull);
    int ncomp = 0;
    const I4 a = ldI4(G.adj4 + target), b = ldI4(G.adj4 + target + 1);
    for (int j = a.x; j < b.x; j++) {
        int nR = 1; { const int blk = __ffsll((long long)freeb) - 1; freeb &= ~(1ull << blk); rblk[0] = (int16_t)blk; rlen[0] = 1; ss[blk * SEED_SCAN_PLEN] = ldI4(G.out_adj4 + j).x; }
        while (nR > 0) {
            for (int eI = nR - 1; eI >= 0; eI--) {
                const int blk = rblk[eI]; const int len = rlen[eI]; int32_t* w = ss + blk * SEED_SCAN_PLEN;
                const int tip = w[len - 1]; const uint8_t em = (uint8_t)(__ldg(G.edge_pack + tip) >> 16);
                bool erase = false;
                if (em == '_') {
                    const int tn = __ldg(G.edge_to + tip); const I4 a2 = ldI4(G.adj4 + tn); const int d2 = ldI4(G.adj4 + tn + 1).x - a2.x;
                    if (d2 == 0) erase = true;
                    else {
                        if (len + 1 > SEED_SCAN_PLEN - 2) return -1;
                        for (int m = 1; m < d2; m++) {
                            if (!freeb || nR >= SEED_SCAN_NB) return -1;
                            const int nb = __ffsll((long long)freeb) - 1; freeb &= ~(1ull << nb);
                            int32_t* w2 = ss + nb * SEED_SCAN_PLEN; for (int i = 0; i < len; i++) w2[i] = w[i];
                            w2[len] = ldI4(G.out_adj4 + a2.x + m).x; rblk[nR] = (int16_t)nb; rlen[nR] = (int16_t)(len + 1); nR++;
                        }
                        w[len] = ldI4(G.out_adj4 + a2.x).x; rlen[eI] = (int16_t)(len + 1);
                    }
                } else {
                    if (em == c) {
                        if (ncomp >= SEED_SCAN_CB) return -1;
                        int32_t* o = ss + (SEED_SCAN_NB + ncomp) * SEED_SCAN_PLEN; o[0] = len; o[1] = __ldg(G.edge_to + tip); for (int i = 0; i < len; i++) o[2 + i] = w[i];
                        ncomp++;
                    }
                    erase = true;
                }
                if (erase) { freeb |= 1ull << blk; for (int i = eI; i + 1 < nR; i++) { rblk[i] = rblk[i + 1]; rlen[i] = rlen[i + 1]; } nR--; }
            }
        }
    }
    return ncomp;
}

} // namespace

// RC running chains per read; BIG: second tier over the reads the first one queued
template <int RC, bool BIG> __global__ void __launch_bounds__(SEED_WARPS * 32) k_seed_chains(SeedParams P) {
    extern __shared__ __align__(16) unsigned char seed_smem[];
    typedef SeedWarpT<RC> SeedWarp;
    constexpr int SEED_RC = RC;     // (shadows the first tier's constant inside this kernel)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SeedWarp& W = reinterpret_cast<SeedWarp*>(seed_smem)[warp];
    const DevGraph& G = P.g; const DevKmerIndex& ix = P.ix; const int k = ix.k; const int ecap = P.ecap;
    const size_t gw = (size_t)blockIdx.x * SEED_WARPS + warp;
    int32_t* cs = P.chain_scratch + gw * (size_t)SEED_RC * ecap;
    int32_t* ss = P.scan_scratch + gw * (size_t)(SEED_SCAN_NB + SEED_SCAN_CB) * SEED_SCAN_PLEN;
    for (;;) {
        long long r = 0;
        if (BIG) { int qi = 0; if (lane == 0) qi = atomicAdd(P.counters + 3, 1); qi = __shfl_sync(0xffffffffu, qi, 0); if (qi >= *P.defer_count) break; r = P.defer_list[qi]; }
        else { if (lane == 0) r = atomicAdd(P.counters, 1); r = __shfl_sync(0xffffffffu, r, 0); if (r >= P.n_reads) break; }
        const int64_t rd0 = P.read_off[r]; const int L = (int)(P.read_off[r + 1] - rd0); const uint8_t* seq = P.bases + rd0;
        int status = 0, ord = 0;
        for (int i = lane; i < SEED_RC; i += 32) W.free_slots[i] = (uint16_t)(SEED_RC - 1 - i);
        if (lane == 0) { W.n_run = 0; W.n_free = SEED_RC; }
        __syncwarp();
        // warp-cooperative: archive running chain i (forReturn.push_back, :54-56)
        auto archive = [&](int i, int end) {
            const RunEnt e = W.run[i];
            unsigned long long ri = 0, eo = 0;
            if (lane == 0) { ri = atomicAdd(P.rec_count, 1ull); eo = atomicAdd(P.edge_count, (unsigned long long)e.n); }
            ri = __shfl_sync(0xffffffffu, ri, 0); eo = __shfl_sync(0xffffffffu, eo, 0);
            if ((long long)ri < P.rec_cap && (long long)(eo + e.n) <= P.edge_cap) {
                const int32_t* src = cs + (size_t)e.slot * ecap;
                for (int j = lane; j < e.n; j += 32) P.edge_pool[eo + j] = src[j];
                if (lane == 0) { SeedChainRec rec; rec.read = (int32_t)r; rec.ord = ord; rec.begin = e.begin; rec.end = end; rec.edge_off = (long long)eo; rec.n_edges = e.n; rec.pad = BIG ? 1 : 0; P.recs[ri] = rec; }
            } else if (lane == 0) P.counters[2] = 1;
            ord++;
        };
        // warp-cooperative: new running chain from index position p (:62-69, :333-339)
        auto add_from_index = [&](long long p, int begin) {
            const long long eo = ix.pos_edge_off[p]; const int n = (int)(ix.pos_edge_off[p + 1] - eo);
            if (W.n_run >= SEED_RC || n > ecap) { status = HLALA_E_CAPACITY_DEV; return; }
            const int slot = W.free_slots[W.n_free - 1];
            int32_t* dst = cs + (size_t)slot * ecap;
            for (int j = lane; j < n; j += 32) dst[j] = __ldg(ix.pos_edges + eo + j);
            __syncwarp();
            if (lane == 0) { RunEnt e; e.begin = begin; e.n = n; e.target = __ldg(ix.pos_to + p); e.slot = slot; W.run[W.n_run] = e; W.n_run++; W.n_free--; }
            __syncwarp();
        };
        if (L >= k) {
            const int n_kmers = L - k + 1;
            int qbase = 0; int ids = (lane < n_kmers) ? kmer_lookup(ix, seq + lane) : -1;
            { const int id0 = __shfl_sync(0xffffffffu, ids, 0);
              if (id0 >= 0) for (long long p = ix.kmer_pos_off[id0]; p < ix.kmer_pos_off[id0 + 1] && !status; p++) add_from_index(p, 0); }
            for (int seqI = k; seqI < L && !status; seqI++) {
                const uint8_t c = seq[seqI];
                // ---- extend the running chains, back to front (:80-304)
                const int n0 = W.n_run;
                for (int hi = n0; hi > 0 && !status; hi -= 32) {
                    const int idx = hi - 1 - lane; const bool active = idx >= 0;
                    int ncomp = 0; bool slow = false; int cm_e[4], cm_t[4]; RunEnt ent; ent.n = 0; ent.slot = 0; ent.target = 0; ent.begin = 0;
                    if (active) {
                        ent = W.run[idx];
                        const I4 a = ldI4(G.adj4 + ent.target), b = ldI4(G.adj4 + ent.target + 1);
                        for (int j = a.x; j < b.x; j++) {
                            const I4 rec = ldI4(G.out_adj4 + j); const uint8_t em = (uint8_t)((uint32_t)rec.z >> 16);
                            if (em == '_') { slow = true; break; }
                            if (em == c) { if (ncomp < 4) { cm_e[ncomp] = rec.x; cm_t[ncomp] = rec.y; } ncomp++; }
                        }
                        if (ncomp > 4) slow = true;
                        if (!slow && ncomp == 1) {   // the common case: one continuation, one edge (:281-285)
                            if (ent.n >= ecap) slow = true;   // reported below as a capacity error
                            else { cs[(size_t)ent.slot * ecap + ent.n] = cm_e[0]; W.run[idx].n = ent.n + 1; W.run[idx].target = cm_t[0]; }
                        }
                    }
                    __syncwarp();
                    unsigned exc = __ballot_sync(0xffffffffu, active && (slow || ncomp != 1));
                    while (exc && !status) {
                        const int l = __ffs(exc) - 1; exc &= exc - 1;
                        const int j = hi - 1 - l;
                        int nc = 0;
                        if (lane == l) {
                            if (slow) nc = (ent.n >= ecap) ? -1 : scan_slow(G, ent.target, c, ss);
                            else { nc = ncomp; for (int i = 0; i < ncomp; i++) { int32_t* o = ss + (SEED_SCAN_NB + i) * SEED_SCAN_PLEN; o[0] = 1; o[1] = (i == 0 ? cm_t[0] : i == 1 ? cm_t[1] : i == 2 ? cm_t[2] : cm_t[3]); o[2] = (i == 0 ? cm_e[0] : i == 1 ? cm_e[1] : i == 2 ? cm_e[2] : cm_e[3]); } }
                        }
                        nc = __shfl_sync(0xffffffffu, nc, l);
                        __syncwarp();
                        if (nc < 0) { status = HLALA_E_CAPACITY_DEV; break; }
                        if (nc == 0) { archive(j, seqI - 1); if (lane == 0) W.run[j].n = -1; __syncwarp(); continue; }
                        const RunEnt pe = W.run[j];
                        // continuations 1.. become new chains appended to the running list in order (:287-303)
                        for (int i = 1; i < nc && !status; i++) {
                            const int32_t* o = ss + (SEED_SCAN_NB + i) * SEED_SCAN_PLEN; const int pl = o[0];
                            if (W.n_run >= SEED_RC || pe.n + pl > ecap) { status = HLALA_E_CAPACITY_DEV; break; }
                            const int slot = W.free_slots[W.n_free - 1];
                            int32_t* dst = cs + (size_t)slot * ecap; const int32_t* src = cs + (size_t)pe.slot * ecap;
                            for (int t = lane; t < pe.n; t += 32) dst[t] = src[t];
                            for (int t = lane; t < pl; t += 32) dst[pe.n + t] = o[2 + t];
                            __syncwarp();
                            if (lane == 0) { RunEnt e; e.begin = pe.begin; e.n = pe.n + pl; e.target = o[1]; e.slot = slot; W.run[W.n_run] = e; W.n_run++; W.n_free--; }
                            __syncwarp();
                        }
                        if (status) break;
                        {   // continuation 0 extends the chain in place
                            const int32_t* o = ss + (SEED_SCAN_NB + 0) * SEED_SCAN_PLEN; const int pl = o[0];
                            if (pe.n + pl > ecap) { status = HLALA_E_CAPACITY_DEV; break; }
                            int32_t* dst = cs + (size_t)pe.slot * ecap;
                            for (int t = lane; t < pl; t += 32) dst[pe.n + t] = o[2 + t];
                            __syncwarp();
                            if (lane == 0) { W.run[j].n = pe.n + pl; W.run[j].target = o[1]; }
                            __syncwarp();
                        }
                    }
                    __syncwarp();
                }
                if (status) break;
                {   // drop the archived chains, keeping the order of the others (runningChains.erase, :264)
                    const int n1 = W.n_run; int w = 0; int nf = W.n_free;
                    __syncwarp();
                    for (int base = 0; base < n1; base += 32) {
                        const int i = base + lane; const bool in = i < n1; RunEnt e; e.n = 0; e.slot = 0; e.begin = 0; e.target = 0; if (in) e = W.run[i];
                        const bool keep = in && e.n >= 0, dead = in && e.n < 0;
                        const unsigned km = __ballot_sync(0xffffffffu, keep), dm = __ballot_sync(0xffffffffu, dead);
                        __syncwarp();
                        if (keep) W.run[w + __popc(km & ((1u << lane) - 1u))] = e;
                        if (dead) W.free_slots[nf + __popc(dm & ((1u << lane) - 1u))] = (uint16_t)e.slot;
                        w += __popc(km); nf += __popc(dm);
                        __syncwarp();
                    }
                    if (lane == 0) { W.n_run = w; W.n_free = nf; }
                    __syncwarp();
                }
                // ---- k-mer ending at this base (:306-341)
                const int q = seqI - k + 1;
                if (q >= qbase + 32) { qbase += 32; ids = (qbase + lane < n_kmers) ? kmer_lookup(ix, seq + qbase + lane) : -1; }
                const int id = __shfl_sync(0xffffffffu, ids, q - qbase);
                if (id >= 0) {
                    for (long long p = ix.kmer_pos_off[id]; p < ix.kmer_pos_off[id + 1] && !status; p++) {
                        const int hn = (int)(ix.pos_edge_off[p + 1] - ix.pos_edge_off[p]); const int hto = __ldg(ix.pos_to + p), hfrom = __ldg(ix.pos_from + p);
                        bool rep = false; const int n1 = W.n_run;
                        for (int base = 0; base < n1 && !rep; base += 32) {
                            const int i = base + lane; bool mine = false;
                            if (i < n1) { const RunEnt e = W.run[i]; if (e.target == hto && e.n >= hn) mine = __ldg(G.edge_from + cs[(size_t)e.slot * ecap + e.n - hn]) == hfrom; }
                            rep = __any_sync(0xffffffffu, mine);
                        }
                        if (!rep) add_from_index(p, q);
                    }
                }
            }
            // ---- remaining chains in list order (:344-348)
            if (!status) { const int n1 = W.n_run; for (int i = 0; i < n1; i++) archive(i, L - 1); }
        }
        if (lane == 0) {
            P.read_n_chains[r] = status ? 0 : ord; P.read_status[r] = status;
            if (!BIG && status == HLALA_E_CAPACITY_DEV && P.defer_list) { P.read_tier[r] = 1; P.defer_list[atomicAdd(P.defer_count, 1)] = (int32_t)r; }     // the records archived so far are void: the second tier starts over
            else { P.read_tier[r] = BIG ? 1 : 0; if (status) atomicAdd(P.counters + 1, 1); }
        }
        __syncwarp();
    }
}

__global__ void k_seed_order(const SeedChainRec* recs, long long n, const long long* chain_off, const int32_t* read_status, const uint8_t* read_tier, SeedChainRec* out, int32_t* out_n_edges) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SeedChainRec r = recs[i];
    if (read_status[r.read] != 0 || (int)read_tier[r.read] != r.pad) return;   // chains archived before the read hit a capacity error are dropped with it (and replaced by the next tier's)
    const long long o = chain_off[r.read] + r.ord;
    out[o] = r; out_n_edges[o] = r.n_edges;
}

__global__ void k_seed_gather(DevGraph G, const SeedChainRec* ordered, long long n, const long long* edge_off, const int32_t* pool, int32_t* out_begin, int32_t* out_end, int32_t* out_edges) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const SeedChainRec r = ordered[w];
    if (lane == 0) { out_begin[w] = r.begin; out_end[w] = r.end; }
    const long long o = edge_off[w];
    for (int j = lane; j < r.n_edges; j += 32) out_edges[o + j] = G.edge_ord[pool[r.edge_off + j]];
}

static size_t seed_smem_bytes(bool big) { return (big ? sizeof(SeedWarpT<SEED_RC_BIG>) : sizeof(SeedWarpT<SEED_RC>)) * SEED_WARPS; }
int seed_warps_for(int n_sm) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seed_chains<SEED_RC, false>, SEED_WARPS * 32, seed_smem_bytes(false)) != cudaSuccess || per_sm < 1) per_sm = 1;
    return n_sm * per_sm * SEED_WARPS;
}
int seed_big_warps_for(int n_sm) { return std::max(1, n_sm / 2) * SEED_WARPS; }     // a handful of reads: half a wave of CTAs keeps the chain scratch (1024 chains per warp) small
size_t seed_scan_scratch_ints() { return (size_t)(SEED_SCAN_NB + SEED_SCAN_CB) * SEED_SCAN_PLEN; }

cudaError_t launch_seed_chains(const SeedParams& P, int n_sm, cudaStream_t st) {
    if (P.n_reads <= 0) return cudaSuccess;
    const int warps = seed_warps_for(n_sm);
    long long want = (P.n_reads + SEED_WARPS - 1) / SEED_WARPS;
    int grid = (int)std::min<long long>(want, warps / SEED_WARPS); if (grid < 1) grid = 1;
    k_seed_chains<SEED_RC, false><<<grid, SEED_WARPS * 32, seed_smem_bytes(false), st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_seed_chains_big(const SeedParams& P, int n_sm, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_seed_chains<SEED_RC_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem_bytes(true)); if (e != cudaSuccess) return e;
    k_seed_chains<SEED_RC_BIG, true><<<seed_big_warps_for(n_sm) / SEED_WARPS, SEED_WARPS * 32, seed_smem_bytes(true), st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_seed_order(const SeedChainRec* recs, long long n_recs, const long long* chain_off, const int32_t* read_status, const uint8_t* read_tier, SeedChainRec* out, int32_t* out_n_edges, cudaStream_t st) {
    if (n_recs <= 0) return cudaSuccess;
    k_seed_order<<<(unsigned)((n_recs + 255) / 256), 256, 0, st>>>(recs, n_recs, chain_off, read_status, read_tier, out, out_n_edges);
    return cudaGetLastError();
}

cudaError_t launch_seed_gather(const DevGraph& g, const SeedChainRec* ordered, long long n_recs, const long long* edge_off, const int32_t* edge_pool,
                               int32_t* out_begin, int32_t* out_end, int32_t* out_edges, cudaStream_t st) {
    if (n_recs <= 0) return cudaSuccess;
    k_seed_gather<<<(unsigned)((n_recs * 32 + 255) / 256), 256, 0, st>>>(g, ordered, n_recs, edge_off, edge_pool, out_begin, out_end, out_edges);
    return cudaGetLastError();
}

namespace { struct ToI64 { __host__ __device__ long long operator()(int32_t v) const { return (long long)v; } }; }
cudaError_t seed_exclusive_scan_i32_to_i64(const int32_t* in, long long n, long long* out, void* temp, size_t temp_bytes, size_t* need_bytes, cudaStream_t st) {
    auto it = thrust::make_transform_iterator(in, ToI64());
    if (!temp) { size_t need = 0; cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, need, it, out + 1, (int)n, st); *need_bytes = need; return e; }
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(long long), st); if (e != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    return cub::DeviceScan::InclusiveSum(temp, temp_bytes, it, out + 1, (int)n, st);
}

} // namespace hlala
