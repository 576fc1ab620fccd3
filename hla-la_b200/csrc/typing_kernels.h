// Launchers of the HLA typing kernels (typing_kernels.cu). Host-callable, no CUDA types beyond cudaStream_t / cudaError_t.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hlala {

// Terms of the per-read x cluster log-likelihood (HLATyper.cpp:2049-2277), computed once on the host with the reference's own libm
// expressions; the kernels only select and add them (rounding step by rounding step like the reference).
struct TypingScoreTables {
    double log_pc[256];        // log(pCorrect(q)) with the 0.999 cap and the 0 -> 0.001 floor (HLATyper.cpp:2189-2200)
    double log_pi[256];        // log((1 - pCorrect(q)) * (1/3))
    double ll_ins_actual;      // log(0.001) + log(1/4)
    double ll_del;             // log(0.001)
    double ll_mm;              // log(1 - 0.001 - 0.001)
    double log_half;           // log(0.5)      (Utilities::logAvg, Utilities.cpp:1368)
    double log_two;            // log(1 + exp(0)): both log-likelihoods equal
};
cudaError_t upload_typing_tables(const TypingScoreTables& t);

constexpr int TY_MAX_GENES = 64;
struct GeneBounds { int32_t n; int32_t first[TY_MAX_GENES]; int32_t last[TY_MAX_GENES]; };

// a19: flag[p] = 1 if either mate's [first level, last level] overlaps a gene (processBAM.cpp:2427-2446)
cudaError_t launch_gene_filter(const GeneBounds& gb, int64_t n_pairs, int32_t maxcol, const int32_t* n_cols, const int32_t* level, uint8_t* flag, cudaStream_t st);
// compact the columns / bases / qualities of the selected reads: src_read[i] -> positions col_off[i].., base_off[i]..
cudaError_t launch_gather_reads(int64_t n_sel, const int64_t* src_read, const int64_t* col_off, const int64_t* base_off, int32_t maxcol, const int32_t* n_cols,
                                const int32_t* level, const uint8_t* g, const uint8_t* s, const uint8_t* mq, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals,
                                int32_t* o_level, uint8_t* o_g, uint8_t* o_s, uint8_t* o_mq, uint8_t* o_bases, uint8_t* o_quals, cudaStream_t st);

// a25: LLt[r][c], mmT[r][c] for reads r in [r0, r1), clusters c < C; clusterT is [P][Cpad] (cluster symbols, position-major)
cudaError_t launch_read_cluster_ll(const uint8_t* clusterT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, int32_t max_rec, const int32_t* rec_off, const int16_t* rec_pos,
                                   const uint8_t* rec_c0, const uint8_t* rec_q0, const uint16_t* rec_glen, double* LLt, int32_t* mmT, cudaStream_t st);
// a26: for every cluster pair c1 <= c2 (reference loop order), sums over reads r in [r0, r1) in ascending r
cudaError_t launch_allele_pair_ll(const double* LLt, const int32_t* mmT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, double* pair_ll, double* pair_mavg, double* pair_mmin, cudaStream_t st);
// the same sums, max-shifted (one exp per (cluster, read)); Et [R * Cpad], rowmax [R], colsum [C], msum [1] are scratch
cudaError_t launch_allele_pair_prod(const double* LLt, const int32_t* mmT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, double* Et, double* rowmax, double* colsum, double* msum,
                                    double* pair_ll, double* pair_mavg, double* pair_mmin, cudaStream_t st);

} // namespace hlala
