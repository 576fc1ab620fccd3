// Launchers implemented in align_kernels.cu (kept free of kernel syntax so that host-only TUs can include it).
#pragma once
#include "device_types.h"
#include <cuda_runtime.h>

namespace hlala {

struct ChainParams; struct ExtParams; struct PairParams;
cudaError_t upload_score_tables(const ScoreTables& t, const ScoreTables& t_long);
cudaError_t launch_prepare(const ChainParams& P, cudaStream_t stream);
cudaError_t launch_chain_seed(const ChainParams& P, int n_sm, cudaStream_t stream);
int chain_seed_max_warps(const ChainParams& P, int n_sm);
cudaError_t sort_dp_tasks_by_level(const int32_t* keys_in, int32_t* keys_out, const int32_t* tasks_in, int32_t* tasks_out, int n, void* temp, size_t temp_bytes, size_t* need_bytes, cudaStream_t stream);
size_t pair_gslab_bytes(int maxcol, int tier);      // per-warp bytes of the pair kernel's HBM slab for this max_columns (0: the slab fits shared memory)
cudaError_t launch_extend(const ExtParams& E, cudaStream_t stream);
cudaError_t launch_extend_warp(const ExtParams& E, int n_sm, int cfg, bool only_deferred, cudaStream_t stream);
cudaError_t launch_extend_group(const ExtParams& E, int n_sm, cudaStream_t stream);
cudaError_t launch_extend_lean(const ExtParams& E, int n_sm, int cfg, cudaStream_t stream);
cudaError_t launch_sort_dp_tasks(const ChainParams& P, int32_t* sorted, cudaStream_t stream);
int ln_threads_for_any(int n_sm);
size_t ln_thread_rec_bytes();
size_t ln_thread_ahead_bytes();
int gd_groups_for(int n_sm);
size_t gd_group_scratch_bytes();
int wd_warps_for(int n_sm);
size_t wd_warp_scratch_bytes();
cudaError_t launch_chain_finish(const ExtParams& E, int n_sm, cudaStream_t stream);
cudaError_t launch_dp_task_bytes(const ExtParams& E, unsigned long long* out, cudaStream_t stream);
cudaError_t launch_pair(const PairParams& P, int n_sm, cudaStream_t stream);
size_t dp_thread_scratch_bytes();
int dp_ext_cap();
cudaError_t launch_export_chain_columns(const DevGraph& G, int n_chains, int maxcol, const int32_t* n_cols, const int32_t* first_level,
                                        const int32_t* c_edge, int32_t* out_level, int32_t* out_edge_ord, uint8_t* out_gchar, cudaStream_t stream);

// parity hook: y[i] = exp_like_host_libm(x[i]) (exp_libm.cuh), the exp of the mapping-quality posteriors
cudaError_t launch_exp_probe(const double* x, double* y, int n, cudaStream_t stream);

} // namespace hlala
