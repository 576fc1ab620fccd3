// HLA typing kernels (sm_100a): gene filter + compaction of the selected alignments, per-read x allele-cluster log-likelihoods,
// allele-pair log-likelihood sums. Reference: hla/HLATyper.cpp:2049-2364, Utilities::logAvg (Utilities.cpp:1368),
// mapper/processBAM.cpp:2427-2446. FP64 adds are issued in the reference's order with explicit round-to-nearest intrinsics (no
// FMA contraction), so per-read log-likelihoods are bit-identical; the pair sums use the device exp/log and agree to ~1e-15 relative.
#include "typing_kernels.h"

namespace hlala {

__constant__ TypingScoreTables c_ty;
cudaError_t upload_typing_tables(const TypingScoreTables& t) { return cudaMemcpyToSymbol(c_ty, &t, sizeof(TypingScoreTables)); }

// ---------------------------------------------------------------------------------------------------------------
__global__ void k_gene_filter(GeneBounds gb, int64_t n_pairs, int32_t maxcol, const int32_t* __restrict__ n_cols, const int32_t* __restrict__ level, uint8_t* __restrict__ flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    bool inc = false;
    for (int m = 0; m < 2; m++) {
        const int64_t r = 2 * p + m; const int n = n_cols[r]; const int32_t* lv = level + (size_t)r * maxcol;
        int f = -1, l = -1;
        for (int c = 0; c < n; c++) if (lv[c] != -1) { f = lv[c]; break; }
        if (f == -1) continue;
        for (int c = n - 1; c >= 0; c--) if (lv[c] != -1) { l = lv[c]; break; }
        for (int g = 0; g < gb.n; g++) inc |= (gb.last[g] >= f && gb.first[g] <= l);
    }
    flag[p] = inc ? 1 : 0;
}
cudaError_t launch_gene_filter(const GeneBounds& gb, int64_t n_pairs, int32_t maxcol, const int32_t* n_cols, const int32_t* level, uint8_t* flag, cudaStream_t st) {
    if (n_pairs <= 0) return cudaSuccess;
    k_gene_filter<<<(unsigned)((n_pairs + 127) / 128), 128, 0, st>>>(gb, n_pairs, maxcol, n_cols, level, flag);
    return cudaGetLastError();
}

__global__ void k_gather_reads(int64_t n_sel, const int64_t* __restrict__ src_read, const int64_t* __restrict__ col_off, const int64_t* __restrict__ base_off, int32_t maxcol,
                               const int32_t* __restrict__ n_cols, const int32_t* __restrict__ level, const uint8_t* __restrict__ g, const uint8_t* __restrict__ s, const uint8_t* __restrict__ mq,
                               const int64_t* __restrict__ read_off, const uint8_t* __restrict__ bases, const uint8_t* __restrict__ quals,
                               int32_t* __restrict__ o_level, uint8_t* __restrict__ o_g, uint8_t* __restrict__ o_s, uint8_t* __restrict__ o_mq, uint8_t* __restrict__ o_bases, uint8_t* __restrict__ o_quals) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    if (w >= n_sel) return;
    const int64_t r = src_read[w]; const int n = n_cols[r]; const size_t src = (size_t)r * maxcol; const int64_t dst = col_off[w];
    for (int c = lane; c < n; c += 32) { o_level[dst + c] = level[src + c]; o_g[dst + c] = g[src + c]; o_s[dst + c] = s[src + c]; o_mq[dst + c] = mq[src + c]; }
    const int64_t b0 = read_off[r]; const int len = (int)(read_off[r + 1] - b0); const int64_t bd = base_off[w];
    for (int i = lane; i < len; i += 32) { o_bases[bd + i] = bases[b0 + i]; o_quals[bd + i] = quals[b0 + i]; }
}
cudaError_t launch_gather_reads(int64_t n_sel, const int64_t* src_read, const int64_t* col_off, const int64_t* base_off, int32_t maxcol, const int32_t* n_cols,
                                const int32_t* level, const uint8_t* g, const uint8_t* s, const uint8_t* mq, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals,
                                int32_t* o_level, uint8_t* o_g, uint8_t* o_s, uint8_t* o_mq, uint8_t* o_bases, uint8_t* o_quals, cudaStream_t st) {
    if (n_sel <= 0) return cudaSuccess;
    k_gather_reads<<<(unsigned)((n_sel * 32 + 127) / 128), 128, 0, st>>>(n_sel, src_read, col_off, base_off, maxcol, n_cols, level, g, s, mq, read_off, bases, quals, o_level, o_g, o_s, o_mq, o_bases, o_quals);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 4: one CTA = one read pair x 256 clusters. The read's exon observations are turned into three candidate terms each
// (cluster has a gap here / cluster base equals the read's first base / differs) and staged in shared memory; a thread then walks
// its cluster's symbols (position-major table: consecutive threads read consecutive bytes) and adds the selected terms in order.
constexpr int K4_THREADS = 256;
struct K4Rec { double v_gap, v_eq, v_ne; };
__global__ void __launch_bounds__(K4_THREADS) k_read_cluster_ll(const uint8_t* __restrict__ clusterT, int32_t C, int32_t Cpad, int32_t r0, const int32_t* __restrict__ rec_off,
                                                                const int16_t* __restrict__ rec_pos, const uint8_t* __restrict__ rec_c0, const uint8_t* __restrict__ rec_q0, const uint16_t* __restrict__ rec_glen,
                                                                double* __restrict__ LLt, int32_t* __restrict__ mmT) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int r = r0 + blockIdx.y; const int k0 = rec_off[r], n = rec_off[r + 1] - k0;
    K4Rec* rec = reinterpret_cast<K4Rec*>(smem); int16_t* pos = reinterpret_cast<int16_t*>(rec + n); uint8_t* c0 = reinterpret_cast<uint8_t*>(pos + n); uint8_t* mmf = c0 + n;
    for (int k = threadIdx.x; k < n; k += K4_THREADS) {
        const uint8_t ch = rec_c0[k0 + k]; const unsigned glen = rec_glen[k0 + k]; const uint8_t q = rec_q0[k0 + k];
        K4Rec v; uint8_t f;
        if (ch == '_') { v.v_gap = 0.0; v.v_eq = v.v_ne = c_ty.ll_del; f = 0; }      // read has a gap: cluster gap -> 0, cluster base -> deletion (HLATyper.cpp:2160-2180)
        else {
            const double ins = __dmul_rn(c_ty.ll_ins_actual, (double)(glen - 1u));   // (len-1) inserted bases behind the first
            v.v_gap = __dmul_rn(c_ty.ll_ins_actual, (double)glen);                  // cluster gap: every base of the genotype is an insertion
            v.v_eq = __dadd_rn(__dadd_rn(c_ty.ll_mm, c_ty.log_pc[q]), ins);
            v.v_ne = __dadd_rn(__dadd_rn(c_ty.ll_mm, c_ty.log_pi[q]), ins);
            f = (uint8_t)(1 | (glen > 1 ? 2 : 0) | 4);                               // mismatch counted if: cluster gap | equal first base but longer genotype | different base
        }
        rec[k] = v; pos[k] = rec_pos[k0 + k]; c0[k] = ch; mmf[k] = f;
    }
    __syncthreads();
    const int c = blockIdx.x * K4_THREADS + threadIdx.x;
    if (c >= Cpad) return;
    double ll = 0.0; int mm = 0;
    if (c < C) {
        for (int k = 0; k < n; k++) {
            const uint8_t e = __ldg(clusterT + (size_t)pos[k] * Cpad + c);
            const int sel = (e == '_') ? 0 : (e == c0[k] ? 1 : 2);
            const double v = sel == 0 ? rec[k].v_gap : (sel == 1 ? rec[k].v_eq : rec[k].v_ne);
            ll = __dadd_rn(ll, v);
            mm += (mmf[k] >> sel) & 1;
        }
    }
    LLt[(size_t)r * Cpad + c] = ll; mmT[(size_t)r * Cpad + c] = mm;
}
cudaError_t launch_read_cluster_ll(const uint8_t* clusterT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, int32_t max_rec, const int32_t* rec_off, const int16_t* rec_pos,
                                   const uint8_t* rec_c0, const uint8_t* rec_q0, const uint16_t* rec_glen, double* LLt, int32_t* mmT, cudaStream_t st) {
    if (r1 <= r0 || C <= 0) return cudaSuccess;
    const size_t smem = (size_t)std::max(max_rec, 1) * (sizeof(K4Rec) + 2 + 1 + 1) + 16;
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k_read_cluster_ll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    dim3 grid((unsigned)((Cpad + K4_THREADS - 1) / K4_THREADS), (unsigned)(r1 - r0));
    k_read_cluster_ll<<<grid, K4_THREADS, smem, st>>>(clusterT, C, Cpad, r0, rec_off, rec_pos, rec_c0, rec_q0, rec_glen, LLt, mmT);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 5: allele-pair sums. One thread owns one (c1, c2) pair of a 16 x 16 tile and walks the reads in ascending order (the
// reference's summation order); the two 16-cluster strips of LLt / mmT are staged through shared memory 32 reads at a time.
constexpr int K5_T = 16, K5_R = 32;
__device__ __forceinline__ double log_avg_dev(double a, double b) {     // Utilities::logAvg
    const double hi = a > b ? a : b, d = a > b ? b - a : a - b;          // d <= 0
    if (d == 0.0) return __dadd_rn(c_ty.log_half, __dadd_rn(c_ty.log_two, hi));
    if (d < -40.0) return __dadd_rn(c_ty.log_half, __dadd_rn(0.0, hi));  // 1 + exp(d) rounds to 1, log(1) = 0
    return __dadd_rn(c_ty.log_half, __dadd_rn(log(__dadd_rn(1.0, exp(d))), hi));
}
__global__ void __launch_bounds__(K5_T * K5_T) k_allele_pair_ll(const double* __restrict__ LLt, const int32_t* __restrict__ mmT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1,
                                                                double* __restrict__ pair_ll, double* __restrict__ pair_mavg, double* __restrict__ pair_mmin) {
    if (blockIdx.x < blockIdx.y) return;                                  // tiles with c2-tile >= c1-tile only
    __shared__ double sA[K5_R][K5_T], sB[K5_R][K5_T]; __shared__ int32_t mA[K5_R][K5_T], mB[K5_R][K5_T];
    const int tx = threadIdx.x & (K5_T - 1), ty = threadIdx.x / K5_T;
    const int c1 = blockIdx.y * K5_T + ty, c2 = blockIdx.x * K5_T + tx;
    double pl = 0.0, sa = 0.0, sm = 0.0;
    for (int rb = r0; rb < r1; rb += K5_R) {
        const int nr = min(K5_R, r1 - rb);
        for (int i = threadIdx.x; i < K5_R * K5_T; i += K5_T * K5_T) {
            const int rr = i / K5_T, cc = i % K5_T;
            if (rr < nr) { const size_t row = (size_t)(rb + rr) * Cpad; const int a = blockIdx.y * K5_T + cc, b = blockIdx.x * K5_T + cc;
                sA[rr][cc] = a < Cpad ? LLt[row + a] : 0.0; mA[rr][cc] = a < Cpad ? mmT[row + a] : 0; sB[rr][cc] = b < Cpad ? LLt[row + b] : 0.0; mB[rr][cc] = b < Cpad ? mmT[row + b] : 0; }
        }
        __syncthreads();
        if (c1 < C && c2 < C && c2 >= c1) {
            for (int rr = 0; rr < nr; rr++) {
                const int m1 = mA[rr][ty], m2 = mB[rr][tx];
                pl = __dadd_rn(pl, log_avg_dev(sA[rr][ty], sB[rr][tx]));
                sa = __dadd_rn(sa, (double)(m1 + m2) / 2.0);
                sm = __dadd_rn(sm, (double)min(m1, m2));
            }
        }
        __syncthreads();
    }
    if (c1 < C && c2 < C && c2 >= c1) {
        const size_t idx = (size_t)c1 * C - (size_t)c1 * (c1 - 1) / 2 + (size_t)(c2 - c1);     // position in the reference's c1 <= c2 loop order
        pair_ll[idx] = pl; pair_mavg[idx] = sa; pair_mmin[idx] = sm;
    }
}
cudaError_t launch_allele_pair_ll(const double* LLt, const int32_t* mmT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, double* pair_ll, double* pair_mavg, double* pair_mmin, cudaStream_t st) {
    if (C <= 0) return cudaSuccess;
    const unsigned nt = (unsigned)((C + K5_T - 1) / K5_T);
    k_allele_pair_ll<<<dim3(nt, nt), K5_T * K5_T, 0, st>>>(LLt, mmT, C, Cpad, r0, r1, pair_ll, pair_mavg, pair_mmin);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 5, max-shifted: the same sums with ONE exp per (cluster, read) instead of one exp + one log per (cluster pair, read).
//   sum_r logAvg(a_r, b_r) = sum_r m_r + n log(1/2) + log prod_r (e^{a_r - m_r} + e^{b_r - m_r}),   m_r = max_c LL[c][r]
// The product is carried as a mantissa in [1, 2) and an integer exponent (one multiply, one add and a few integer operations per
// term), so it can neither overflow nor underflow; a term whose two exponentials both vanish (both clusters > 690 nats below the
// read's best cluster: two alleles with a deletion where the read has bases) is computed term-wise from the log-likelihoods themselves
// and added to a side sum. Deviation from the term-by-term sums: ~1e-15 relative.
// The mismatch sums are integers: the average is (S[c1] + S[c2]) / 2 with S the per-cluster totals, the minimum an integer sum.
__global__ void k_read_exp_shift(const double* __restrict__ LLt, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, double* __restrict__ Et, double* __restrict__ rowmax) {
    const int r = r0 + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); const int lane = threadIdx.x & 31;
    if (r >= r1) return;
    const double* row = LLt + (size_t)r * Cpad;
    double m = -1.0e300;
    for (int c = lane; c < C; c += 32) m = fmax(m, row[c]);
    for (int d = 16; d; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    for (int c = lane; c < Cpad; c += 32) Et[(size_t)r * Cpad + c] = c < C ? exp(row[c] - m) : 0.0;
    if (lane == 0) rowmax[r] = m;
}
__global__ void k_typing_col_sums(const int32_t* __restrict__ mmT, const double* __restrict__ rowmax, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, double* __restrict__ colsum, double* __restrict__ msum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { long long s = 0; for (int r = r0; r < r1; r++) s += mmT[(size_t)r * Cpad + c]; colsum[c] = (double)s; }
    if (c == 0) { double t = 0.0; for (int r = r0; r < r1; r++) t = __dadd_rn(t, rowmax[r]); *msum = t; }     // one thread, ascending reads: a fixed order
}
__global__ void __launch_bounds__(K5_T * K5_T) k_allele_pair_prod(const double* __restrict__ LLt, const double* __restrict__ Et, const int32_t* __restrict__ mmT, const double* __restrict__ colsum,
                                                                  const double* __restrict__ msum, const double* __restrict__ rowmax, int32_t C, int32_t Cpad, int32_t r0, int32_t r1,
                                                                  double* __restrict__ pair_ll, double* __restrict__ pair_mavg, double* __restrict__ pair_mmin) {
    if (blockIdx.x < blockIdx.y) return;
    __shared__ double sA[K5_R][K5_T], sB[K5_R][K5_T]; __shared__ int32_t mA[K5_R][K5_T], mB[K5_R][K5_T];
    const int tx = threadIdx.x & (K5_T - 1), ty = threadIdx.x / K5_T;
    const int c1 = blockIdx.y * K5_T + ty, c2 = blockIdx.x * K5_T + tx;
    const bool mine = c1 < C && c2 < C && c2 >= c1;
    double acc = 1.0, extra = 0.0; long long esum = 0; long long sm = 0;
    for (int rb = r0; rb < r1; rb += K5_R) {
        const int nr = min(K5_R, r1 - rb);
        for (int i = threadIdx.x; i < K5_R * K5_T; i += K5_T * K5_T) {
            const int rr = i / K5_T, cc = i % K5_T;
            if (rr < nr) { const size_t row = (size_t)(rb + rr) * Cpad; const int a = blockIdx.y * K5_T + cc, b = blockIdx.x * K5_T + cc;
                sA[rr][cc] = a < Cpad ? Et[row + a] : 0.0; mA[rr][cc] = a < Cpad ? mmT[row + a] : 0; sB[rr][cc] = b < Cpad ? Et[row + b] : 0.0; mB[rr][cc] = b < Cpad ? mmT[row + b] : 0; }
        }
        __syncthreads();
        if (mine) {
            for (int rr = 0; rr < nr; rr++) {
                const double t = __dadd_rn(sA[rr][ty], sB[rr][tx]);
                if (t > 1.0e-300) {
                    acc = __dmul_rn(acc, t);
                    const int hi = __double2hiint(acc);
                    esum += ((hi >> 20) & 0x7ff) - 1023;
                    acc = __hiloint2double((hi & 0x800fffff) | (1023 << 20), __double2loint(acc));
                } else {      // both exponentials vanished: this term from the log-likelihoods, relative to what the closing formula adds per read
                    const size_t row = (size_t)(rb + rr) * Cpad;
                    extra = __dadd_rn(extra, __dadd_rn(log_avg_dev(LLt[row + c1], LLt[row + c2]), -__dadd_rn(rowmax[rb + rr], c_ty.log_half)));
                }
                sm += min(mA[rr][ty], mB[rr][tx]);
            }
        }
        __syncthreads();
    }
    if (mine) {
        const double pl = __dadd_rn(__dadd_rn(__dadd_rn(*msum, __dmul_rn((double)(r1 - r0), c_ty.log_half)), __dadd_rn(__dmul_rn((double)esum, c_ty.log_two), log(acc))), extra);
        const size_t idx = (size_t)c1 * C - (size_t)c1 * (c1 - 1) / 2 + (size_t)(c2 - c1);
        pair_ll[idx] = pl; pair_mavg[idx] = (colsum[c1] + colsum[c2]) / 2.0; pair_mmin[idx] = (double)sm;
    }
}
cudaError_t launch_allele_pair_prod(const double* LLt, const int32_t* mmT, int32_t C, int32_t Cpad, int32_t r0, int32_t r1, double* Et, double* rowmax, double* colsum, double* msum,
                                    double* pair_ll, double* pair_mavg, double* pair_mmin, cudaStream_t st) {
    if (C <= 0) return cudaSuccess;
    if (r1 > r0) k_read_exp_shift<<<(unsigned)(((size_t)(r1 - r0) * 32 + 127) / 128), 128, 0, st>>>(LLt, C, Cpad, r0, r1, Et, rowmax);
    k_typing_col_sums<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(mmT, rowmax, C, Cpad, r0, r1, colsum, msum);
    const unsigned nt = (unsigned)((C + K5_T - 1) / K5_T);
    k_allele_pair_prod<<<dim3(nt, nt), K5_T * K5_T, 0, st>>>(LLt, Et, mmT, colsum, msum, rowmax, C, Cpad, r0, r1, pair_ll, pair_mavg, pair_mmin);
    return cudaGetLastError();
}

} // namespace hlala
