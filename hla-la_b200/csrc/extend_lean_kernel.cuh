// Kernel around extend_lean.h: one THREAD per extension task, persistent grid.
//
// The per-thread words of LnDp (two wavefront lists, the touch table, the slot order) sit in shared memory, interleaved by lane
// (word i of lane l at slab[i * 32 + l]): whatever index a thread uses it stays in its own bank, so the 32 unrelated extensions of a warp
// never conflict. Finished cells go to a per-thread slab of 16-byte records in HBM (written once, L2-resident while the extension runs).
//
// Every thread is a small state machine (needs a task / running / ended, waiting for its backtrace / out of tasks). The backtrace and the
// task set-up are chains of dependent HBM loads; they run when LN_BATCH threads of the warp are waiting (or nothing else is left to do), so
// that their latencies overlap instead of stalling 31 running extensions once per finished task.
#pragma once
#include "extend_lean.h"
#include "chain_params.h"
#include <cuda_runtime.h>

namespace hlala {

constexpr int LN_BLOCK = 64;       // threads per CTA: one warp, so that the resident warps per SM follow the shared-memory budget in steps of one
constexpr int LN_BATCH = 16;       // default of ExtParams::ln_batch: waiting threads of a warp that trigger the backtrace / fetch phase

struct LnSmem {
    uint32_t* base;
    __device__ __forceinline__ uint32_t& operator()(int i) const { return base[i * 32]; }
    __device__ __forceinline__ uint8_t& b(int word, int byte) const { return reinterpret_cast<uint8_t*>(base + word * 32)[byte]; }
};

template <class CFG> __global__ void __launch_bounds__(LN_BLOCK) k_extend_lean(ExtParams E) {
    extern __shared__ __align__(16) uint32_t ln_smem[];
    const ChainParams& P = E.C; const DevGraph& DG = P.g; const DevBatch& B = P.b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LnSmem S; S.base = ln_smem + (size_t)warp * CFG::WORDS * 32 + lane;
    const int gt = blockIdx.x * LN_BLOCK + threadIdx.x;
    LnRec* rec = E.ln_rec + (size_t)gt * (LN_CELLS + 1);
    LnAhead* ahead = reinterpret_cast<LnAhead*>(reinterpret_cast<unsigned char*>(E.ln_ahead) + (size_t)gt * (sizeof(LnAhead) * (size_t)LN_AHEAD + (size_t)LN_RING));
    LnGraph G; G.n_levels = DG.n_levels; G.level_node_off = DG.level_node_off; G.level_edge_off = DG.level_edge_off; G.dp_pack = DG.dp_pack; G.lvl4 = reinterpret_cast<const LnLvl*>(DG.lvl4);
    G.path_off = DG.path_off; G.path_edges = DG.path_edges; G.path_from = DG.path_from; G.path_to = DG.path_to;
    G.jump_fwd_off = DG.jump_fwd_off; G.jump_fwd_path = DG.jump_fwd_path; G.jump_bwd_off = DG.jump_bwd_off; G.jump_bwd_path = DG.jump_bwd_path;
    typedef LnDp<CFG, LnSmem> DP;
    const int n_in = *E.in_count;
    LnState st{}; int phase = (gt < E.n_ln_threads) ? 0 : 3;      // 0: needs a task, 1: running, 2: ended (backtrace pending), 3: out of tasks
    int t = 0;
    for (;;) {
        const unsigned waiting = __ballot_sync(0xffffffffu, phase == 0 || phase == 2);
        const unsigned running = __ballot_sync(0xffffffffu, phase == 1);
        if (waiting == 0u && running == 0u) break;
        if (waiting != 0u && (__popc(waiting) >= E.ln_batch || running == 0u)) {
            if (phase == 2) {
                DpResult res; const int rc = DP::finish(G, st, rec, E.ext_edge + (size_t)t * DP_EXT_CAP, E.ext_s + (size_t)t * DP_EXT_CAP, res);
                E.ext_rc[t] = rc; E.ext_n[t] = rc == 0 ? res.n_cols : 0; E.ext_nlvl[t] = rc == 0 ? res.n_lvl : 0;
                phase = 0;
            }
            if (phase == 0) {
                const int ti = atomicAdd(E.pop, 1);
                if (ti >= n_in) phase = 3;
                else {
                    t = E.in_list[ti];
                    const int slot = P.pending_slots[t >> 1]; const int side = t & 1;
                    const int r = B.slot_read[slot]; const int64_t rd0 = B.read_off[r]; const int rdlen = (int)(B.read_off[r + 1] - rd0);
                    const int32_t* se_edge = P.c_edge + (size_t)(slot - P.slot_base) * P.maxcol;
                    int start_seq, start_level, start_z;
                    if (side == 0) { start_seq = P.seed_begin[slot]; start_level = P.first_level[slot]; start_z = (int)(DG.edge_pack[se_edge[0]] & 255u); }
                    else { start_seq = P.seed_end[slot] + 1; start_level = P.last_level[slot] + 1; start_z = (int)((DG.edge_pack[se_edge[P.n_cols[slot] - 1]] >> 8) & 255u); }
                    const int rc = DP::init(G, S, st, rec, B.bases + rd0, rdlen, start_seq, start_level, start_z, side == 1);
                    if (rc == 0) phase = 1;
                    else { E.ext_rc[t] = rc; if (rc == DP_DEFER) E.out_list[atomicAdd(E.out_count, 1)] = t; }      // stays in phase 0: next batch
                }
            }
        }
        if (phase == 1) {
            const int rc = DP::step(G, S, st, rec, ahead);
            if (rc == 1) phase = 2;
            else if (rc != 0) { E.ext_rc[t] = rc; if (rc == DP_DEFER) E.out_list[atomicAdd(E.out_count, 1)] = t; phase = 0; }
        }
    }
}

template <class CFG> static int ln_threads_for(int n_sm) {
    const size_t smem = (size_t)CFG::WORDS * 4 * LN_BLOCK; int per_sm = 1;
    if (cudaFuncSetAttribute(k_extend_lean<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_lean<CFG>, LN_BLOCK, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return n_sm * per_sm * LN_BLOCK;
}

template <class CFG> static cudaError_t launch_ln(ExtParams E, int n_sm, cudaStream_t stream) {
    const int threads = std::min(ln_threads_for<CFG>(n_sm), E.n_ln_threads);
    if (threads < LN_BLOCK) return cudaErrorInvalidConfiguration;
    const int grid = threads / LN_BLOCK;
    E.n_ln_threads = grid * LN_BLOCK;
    k_extend_lean<CFG><<<grid, LN_BLOCK, (size_t)CFG::WORDS * 4 * LN_BLOCK, stream>>>(E);
    return cudaGetLastError();
}

} // namespace hlala
