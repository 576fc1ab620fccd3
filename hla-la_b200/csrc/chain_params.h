// Parameters and capacities shared by the launch code (host) and the chain kernels (device).
#pragma once
#include "device_types.h"
#include <cstddef>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace hlala {

constexpr int K1_WARPS = 4;           // warps per CTA
constexpr int K1_WCAP = 256;          // max nodes per level (edge_pack holds z in 8 bits)
constexpr uint32_t KEY_RANK_MASK = 0xFFFFFu;   // backtrack pool entries: 20 bits of edge rank inside a level, the from node above

struct ChainParams {
    DevGraph g; DevBatch b;
    int32_t maxcol;        // stride of the per-chain column records in HBM (= the caller's max_columns)
    int32_t slab_cols;     // column capacity of the shared-memory slab of this launch (tier 0: small slab, high occupancy; tier 1: maxcol)
    int32_t wcap;          // node-per-level capacity of the slab's Viterbi state
    int32_t tier;          // 0: chains from todo_slots, capacity overflow defers the chain; 1: chains from defer_slots
    int32_t pool_cap;      // backtrack pool entries per warp
    int32_t win_cap;       // staged edge window entries per warp
    int32_t bt16;          // 1: backtrack entries are 16 bit (graph has <= 255 edges per level and <= 256 nodes per level)
    int32_t do_extension;  // 1: run the soft-clip extension DP
    int32_t slot_base, slot_end;   // this launch handles slots [slot_base, slot_end); column scratch is indexed by slot - slot_base
    // per-slot outputs
    int32_t* status; int32_t* n_cols; int32_t* seed_begin; int32_t* seed_end; double* ll;
    int32_t* first_level; int32_t* last_level;
    int32_t* c_edge;       // [n_chains * maxcol] flat edge id or -1
    uint8_t* c_schar;      // [n_chains * maxcol]
    uint8_t* c_fromseed;   // [n_chains * maxcol]
    int32_t* error_count;  // global counter of chains that ended with status < 0
    int32_t* id_first; int32_t* id_last;   // PRG levels of the BAM record's first/last reference base (processBAM.cpp:3840): de-dup key
    int32_t* pending_slots; int32_t* pending_count;   // chains whose seed needs the extension DP
    int32_t* dp_tasks; int32_t* dp_task_count;        // the extension tasks that run: 2 * (index in pending_slots) + side, in no particular order
    uint8_t* dp_task_bin; int32_t* dp_task_hist;      // per task: 255 - min(clipped bases, 255); histogram of the bins (k_sort_dp_tasks: longest first)
    int32_t* dp_task_key;                             // per task: the level the extension starts from (tasks sorted by it run neighbouring graph windows in one warp); unused slots keep a large key
    int32_t* todo_slots; int32_t* todo_count;         // chains the chain kernel has to align (written by k_prepare)
    int32_t* defer_slots; int32_t* defer_count;       // chains that did not fit the tier-0 slab
    int32_t read_begin, read_end;                     // reads of this wave
    int32_t dedup;                                    // 1: chains that the pair stage would discard as duplicates are not aligned at all
    // long-read mode (processBAM::alignOneLongRead, processBAM.cpp:3618-3838): no extension DP (the seed is padded to the read), indel rates 0.075
    // (extensionAligner.cpp:58-64). Chains of thousands of columns do not fit a shared-memory slab: with gslab the per-warp column buffers,
    // backtrack pool and level tables live in a per-warp slice of HBM (gslab_base, gslab_bytes each) and only the staged edge window, the two
    // node-score rows and the copy barrier stay in shared memory.
    int32_t long_mode, gslab; unsigned char* gslab_base; unsigned long long gslab_bytes;
    int32_t key_shift;       // bit position of the score in the packed Viterbi keys (chain_kernel.cuh): 20 unless max_columns needs more score bits
};

constexpr int32_t CH_TODO = -100;
constexpr int32_t CH_DEFERRED = -101;   // transient: did not fit the small slab, handled by the tier-1 launch
constexpr int32_t CH_DUPLICATE = 3;   // same PRG start/stop as an earlier (better or equal AS) chain of the read (processBAM.cpp:3234)

constexpr int32_t CH_PENDING_EXT = 2;

// extension stage (one thread per (pending chain, side))
struct LnRec;
struct ExtParams {
    ChainParams C;
    int32_t n_pending;
    int32_t* ext_edge; uint8_t* ext_s;      // [2 * n_pending * DP_EXT_CAP]
    int32_t* ext_n; int32_t* ext_nlvl; int32_t* ext_rc;   // [2 * n_pending]
    unsigned char* dp_scratch; int32_t n_dp_threads;   // scalar kernel: one slice per thread
    unsigned char* wd_scratch; int32_t n_wd_warps;     // warp kernel: one slice per warp
    unsigned char* gd_scratch; int32_t n_gd_groups;    // group kernel: one slice per 8-lane group
    LnRec* ln_rec; uint32_t* ln_ahead; int32_t n_ln_threads;   // thread-per-extension tier: cell records and ahead table, one slice per thread
    int32_t ln_batch;                                   // thread-per-extension tier: waiting threads of a warp that trigger the backtrace / fetch phase
    int32_t only_deferred;                              // scalar kernel: run only the tasks the warp kernel deferred
    // dynamic task queues of the cascade: a tier pops task indices from `pop`; its tasks are in_list[0 .. *in_count) (in_list == nullptr: all
    // 2 * n_pending tasks) and the tasks it defers for capacity are appended to out_list
    const int32_t* in_list; const int32_t* in_count; int32_t* pop; int32_t* out_list; int32_t* out_count;
};

constexpr int K3_WARPS = 4;
constexpr int K3_KCAP = 32;        // kept chains per read
constexpr int K3_COMBO_CAP = 512;  // chain combinations per pair

struct PairParams {
    DevGraph g; DevBatch b; int32_t maxcol;
    int32_t slot_base; long long pair_begin, pair_end;   // wave: pairs [pair_begin, pair_end), column scratch indexed by slot - slot_base
    // chain records (outputs of the chain stage)
    const int32_t* status; const int32_t* n_cols; const double* ll; const int32_t* first_level; const int32_t* last_level;
    const int32_t* id_first; const int32_t* id_last; const int32_t* c_edge; const uint8_t* c_schar; const uint8_t* c_fromseed;
    // insert-size model: log N(d; mean, sd) for integer d in [is_dmin, is_dmin + is_n), is_penalty elsewhere (processBAM.cpp:3446-3472)
    const double* is_table; int32_t is_dmin; int32_t is_n; double is_penalty;
    const double* phred_thr;       // [223]: smallest Q mapped to phred char 33+k (Utilities::PCorrectToPhred, Utilities.cpp:178-203)
    // per pair / per read outputs
    double* pair_mapq; double* read_mapq; uint8_t* read_reverse; int32_t* chosen_slot; double* pair_ll; int32_t* pair_status;
    int32_t* out_n_cols; int32_t* out_level; int32_t* out_edge; uint8_t* out_gchar; uint8_t* out_schar; uint8_t* out_fromseed; uint8_t* out_mapq;   // [n_reads * maxcol], may be null
    int32_t* bases_per_level;      // [n_levels-1] += , may be null
    int32_t* error_count;
    unsigned long long* digest;    // [4]: sum n_cols, sum (edge ordinal + 1), pairs with mapQ < 1, -
    int32_t* defer_list; int32_t* defer_count;   // pairs beyond the small tier's capacities (kept chains per read, combinations), re-run by the large tier
    int32_t unpaired;                            // long-read mode: every "pair" is one read followed by an empty mate (see k_pair)
    unsigned char* gslab[2]; unsigned long long gslab_warps;   // max_columns beyond what shared memory holds: per-warp slabs of tier 0 / tier 1 in HBM (gslab_warps each)
};

// the shared-memory part of a slab whose large arrays live in HBM (ChainParams::gslab)
__host__ __device__ inline size_t k1_gslab_bytes(int maxcol, int pool_cap) {
    size_t b = (size_t)maxcol * 4 * 2 + (size_t)pool_cap * 4 + (size_t)(maxcol + 4) * 2 * 2 + (size_t)((maxcol + 1) / 2 * 2) * 2 + (size_t)maxcol * 4;
    return (b + 255) & ~size_t(255);
}
__host__ __device__ inline size_t k1_gslab_smem_bytes(int win_cap, int wcap) { return ((size_t)(win_cap + 4) * 4 + 16 + (size_t)wcap * 4 * 2 + 15) & ~size_t(15); }

__host__ __device__ inline size_t k1_slab_bytes(int maxcol, int pool_cap, int win_cap, int wcap) {
    size_t b = 0;
    b += (size_t)maxcol * 4 * 2;          // lvlA, lvlB
    b += (size_t)maxcol * 4;              // gA,sA,gB,sB
    b += (size_t)pool_cap * 4;            // bt (half of it used when entries are 16 bit: pool_cap then counts 16-bit entries * 2)
    b += (size_t)(win_cap + 4) * 4 + 16;  // win (+ slack for the 16-byte aligned bulk copy), mbarrier + phase
    b += (size_t)(maxcol + 4) * 2 * 2;    // weoff, wwid (u16)
    b += (size_t)((maxcol + 1) / 2 * 2) * 2;   // coloff (u16)
    b += (size_t)wcap * 4 * 2;            // cur, nxt
    return (b + 15) & ~size_t(15);
}


} // namespace hlala
