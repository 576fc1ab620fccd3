// k-mer seeding on the GPU: index lookup + exact chain extension (GraphAndEdgeIndex::queryIndex / findChains,
// Graph/GraphAndEdgeIndex.cpp:986, 39-356). Launchers and device views.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "device_types.h"

namespace hlala {

struct DevKmerIndex {
    int32_t k; uint32_t ht_mask; int64_t n_kmers, n_pos;
    const int32_t* ht;             // k-mer id + 1, 0 = empty
    const uint8_t* kmer_bytes;     // [n_kmers * k]
    const int64_t* kmer_pos_off;   // [n_kmers + 1]
    const int64_t* pos_edge_off;   // [n_pos + 1]
    const int32_t* pos_edges;      // flat edges
    const int32_t* pos_from;       // [n_pos] flat node the path starts at
    const int32_t* pos_to;         // [n_pos] flat node the path ends in
};

struct alignas(16) SeedChainRec { int32_t read, ord, begin, end; long long edge_off; int32_t n_edges, pad; };

// capacities of the per-warp working set
constexpr int SEED_RC = 128;        // running chains per read, first tier (practically every read)
constexpr int SEED_RC_BIG = 1024;   // second tier: reads the first tier queued (repeats / reads crossing many alleles keep hundreds of chains running; the reference has no limit)
constexpr int SEED_WARPS = 4;       // warps per CTA
constexpr int SEED_SCAN_PLEN = 256; // edges per path in the gap-scan work area
constexpr int SEED_SCAN_NB = 48;    // paths in flight in the gap-scan work area
constexpr int SEED_SCAN_CB = 16;    // compatible continuations of one chain at one base

struct SeedParams {
    DevGraph g; DevKmerIndex ix;
    long long n_reads; const int64_t* read_off; const uint8_t* bases;
    int32_t ecap;                    // edges per running chain (scratch stride)
    int32_t* chain_scratch;          // [n_warps * SEED_RC * ecap]
    int32_t* scan_scratch;           // [n_warps * (SEED_SCAN_NB + SEED_SCAN_CB) * SEED_SCAN_PLEN]
    int32_t* read_n_chains;          // [n_reads]
    int32_t* read_status;            // [n_reads] 0 ok, HLALA_E_CAPACITY_DEV when a working-set capacity was exceeded
    SeedChainRec* recs; long long rec_cap; unsigned long long* rec_count;
    int32_t* edge_pool; long long edge_cap; unsigned long long* edge_count;
    int32_t* counters;               // [0] next read (dynamic scheduling), [1] reads with errors, [2] output overflow flag, [3] next queued read (second tier)
    int32_t* defer_list; int32_t* defer_count;   // reads beyond a first-tier capacity, re-run by the second tier
    uint8_t* read_tier;              // [n_reads] tier whose chain records count for the read (records carry their tier in SeedChainRec::pad)
};

int seed_warps_for(int n_sm);        // warps of the persistent grid (first tier)
int seed_big_warps_for(int n_sm);    // warps of the second tier's grid
size_t seed_scan_scratch_ints();     // per warp
cudaError_t launch_seed_chains(const SeedParams& P, int n_sm, cudaStream_t st);
cudaError_t launch_seed_chains_big(const SeedParams& P, int n_sm, cudaStream_t st);     // P.chain_scratch sized for seed_big_warps_for() x SEED_RC_BIG chains
// records -> (read, ord) order; chain_off = exclusive scan of read_n_chains (computed by the caller)
cudaError_t launch_seed_order(const SeedChainRec* recs, long long n_recs, const long long* chain_off, const int32_t* read_status, const uint8_t* read_tier, SeedChainRec* out, int32_t* out_n_edges, cudaStream_t st);
// edges of the ordered chains, contiguous, as canonical ordinals; edge_off = exclusive scan of the ordered n_edges
cudaError_t launch_seed_gather(const DevGraph& g, const SeedChainRec* ordered, long long n_recs, const long long* edge_off, const int32_t* edge_pool,
                               int32_t* out_begin, int32_t* out_end, int32_t* out_edges, cudaStream_t st);
cudaError_t seed_exclusive_scan_i32_to_i64(const int32_t* in, long long n, long long* out /* [n+1] */, void* temp, size_t temp_bytes, size_t* need_bytes, cudaStream_t st);

} // namespace hlala
