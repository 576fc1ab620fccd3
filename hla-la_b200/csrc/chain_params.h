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
constexpr uint32_t KEY_RANK_MASK = 0xFFFFFu;   // 20 bits of edge rank inside a level, 11 bits of score+1 above

struct ChainParams {
    DevGraph g; DevBatch b;
    int32_t maxcol;        // capacity of a column slab (and of the output stride)
    int32_t pool_cap;      // backtrack pool entries per warp
    int32_t win_cap;       // staged edge window entries per warp
    int32_t do_extension;  // 1: run the soft-clip extension DP
    // per-slot outputs
    int32_t* status; int32_t* n_cols; int32_t* seed_begin; int32_t* seed_end; double* ll;
    int32_t* first_level; int32_t* last_level;
    int32_t* c_edge;       // [n_chains * maxcol] flat edge id or -1
    uint8_t* c_schar;      // [n_chains * maxcol]
    uint8_t* c_fromseed;   // [n_chains * maxcol]
    int32_t* error_count;  // global counter of chains that ended with status < 0
};

__host__ __device__ inline size_t k1_slab_bytes(int maxcol, int pool_cap, int win_cap) {
    size_t b = 0;
    b += (size_t)maxcol * 4 * 2;          // lvlA, lvlB
    b += (size_t)maxcol * 4;              // gA,sA,gB,sB
    b += (size_t)pool_cap * 4;            // bt
    b += (size_t)win_cap * 4;             // win
    b += (size_t)(maxcol + 2) * 4;        // weoff
    b += (size_t)((maxcol + 1) / 2 * 2) * 2;   // coloff (u16)
    b += (size_t)K1_WCAP * 4 * 2;         // cur, nxt
    return (b + 15) & ~size_t(15);
}


} // namespace hlala
