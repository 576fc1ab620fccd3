// extern "C" boundary of libhlala_b200.so (include/hlala_b200.h). Host logic in C++, compute in CUDA; no CPU fallback.
#include "../../include/hlala_b200.h"
#include "../host/prg_graph.h"
#include "../host/dp_pack.h"
#include "../host/insert_size.h"
#include "extend_lean.h"
#include "chain_params.h"
#include "align_kernels.h"
#include "typing_kernels.h"
#include "seed_kernels.h"
#include "../host/kmer_index.h"
#include "../host/bam_reader.h"
#include "../host/hla_typing.h"
#include "../host/hla_eval.h"
#include "../host/truth_levels.h"
#include "../host/read_simulator.h"
#include "../host/contig_mapper.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <chrono>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

using namespace hlala;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

struct CudaError : std::runtime_error { explicit CudaError(const std::string& m) : std::runtime_error(m) {} };
#define CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

struct DevBuf {
    void* p = nullptr; size_t bytes = 0, cap = 0; bool fresh = false;   // fresh: the last alloc() returned new (uninitialised) memory
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    // grow-only: a buffer that is large enough is reused, so that a workspace kept across calls does not pay cudaMalloc/cudaFree again
    void alloc(size_t n) {
        if (n <= cap && p) { bytes = n; fresh = false; return; }
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        bytes = n; fresh = true; if (n) { CUDA_OK(cudaMalloc(&p, n)); cap = n; }
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    template <class T> void upload(const T* h, size_t count, cudaStream_t st = 0) { alloc(count * sizeof(T)); if (count) CUDA_OK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st)); }
    template <class T> void upload(const std::vector<T>& v, cudaStream_t st = 0) { upload(v.data(), v.size(), st); }
    template <class T> void download(T* h, size_t count, cudaStream_t st = 0) const { if (count) CUDA_OK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, st)); }
};

// Score tables with the reference's own arithmetic (Utilities::PhredToPCorrect Utilities.cpp:357-377,
// extensionAligner::scoreOneAlignment extensionAligner.cpp:58-66,127-146), evaluated by the host libm once.
ScoreTables make_score_tables(bool long_reads = false) {
    ScoreTables t;
    double rate_deletions = log(0.001), rate_insertions = log(0.001);
    if (long_reads) { rate_deletions = log(0.075); rate_insertions = log(0.075); }      // extensionAligner.cpp:60-64
    t.rate_match_mismatch = log(1 - exp(rate_deletions) - exp(rate_insertions));
    t.rate_deletion = rate_deletions;
    t.ins_term = rate_insertions + log(1.0 / 4.0);
    for (int q = 0; q < 256; q++) {
        if (q < 33) { t.log_match[q] = NAN; t.log_mismatch[q] = NAN; continue; }   // the reference asserts illuminaPhred >= 0
        int illuminaPhred = q - 33;
        double log10_pWrong = (double)illuminaPhred / (double)-10;
        double pWrong = exp(log(10) * log10_pWrong);
        double pCorrect = 1 - pWrong;
        if (pCorrect > 0.999) pCorrect = 0.999;
        if (pCorrect == 0) pCorrect = 0.00001;
        t.log_match[q] = log(pCorrect);
        double pIncorrect = 1 - pCorrect; pIncorrect *= (1.0 / 3.0);
        t.log_mismatch[q] = log(pIncorrect);
    }
    return t;
}

} // namespace

struct hlala_graph {
    FlatGraph h;
    bool on_gpu = false; int device = -1; int n_sm = 148;
    DevGraph d{};
    std::vector<std::unique_ptr<DevBuf>> bufs;
    template <class T> const T* up(const std::vector<T>& v) { bufs.emplace_back(new DevBuf()); bufs.back()->upload(v); return bufs.back()->as<T>(); }
    // workspace of hlala_align_pairs kept across calls (device buffers are grow-only), one call at a time per handle
    std::shared_ptr<void> ws; std::mutex ws_mu;
    // k-mer index (hlala_kmer_index_build)
    std::unique_ptr<KmerIndex> kix; DevKmerIndex dkix{}; bool kix_on_gpu = false; std::vector<std::unique_ptr<DevBuf>> kbufs;
    template <class T> const T* kup(const std::vector<T>& v) { kbufs.emplace_back(new DevBuf()); kbufs.back()->upload(v); return kbufs.back()->as<T>(); }
};

namespace {

// Host side of a batch: the reference's chain order (sortChainsInSeeds, processBAM.cpp:1945-1968: std::sort ascending by AS
// then std::reverse — same algorithm, same comparator outcomes, hence the same permutation) and primary lookup.
struct PreparedBatch {
    int64_t n_reads = 0; int32_t n_chains = 0;
    std::vector<int32_t> chain_order, read_primary, slot_read;
    void build(const hlala_seed_batch_t& b) {
        n_reads = b.n_reads; n_chains = b.chain_off[b.n_reads];
        chain_order.resize(n_chains); read_primary.assign(n_reads, -1); slot_read.resize(n_chains);
        auto work = [&](int64_t r0, int64_t r1) {
            std::vector<int32_t> idx;
            for (int64_t r = r0; r < r1; r++) {
                int32_t c0 = b.chain_off[r], c1 = b.chain_off[r + 1];
                idx.resize(c1 - c0); for (int32_t i = 0; i < c1 - c0; i++) idx[i] = c0 + i;
                std::sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return b.chain_as[x] < b.chain_as[y]; });
                std::reverse(idx.begin(), idx.end());
                for (int32_t i = 0; i < c1 - c0; i++) {
                    chain_order[c0 + i] = idx[i]; slot_read[c0 + i] = (int32_t)r;
                    if (read_primary[r] < 0 && !(b.chain_flag[idx[i]] & 0x100)) read_primary[r] = c0 + i;
                }
            }
        };
        unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        if (const char* e = getenv("HLALA_HOST_THREADS")) nt = (unsigned)std::max(1, std::min(64, atoi(e)));     // several ranks on one node: each keeps to its share of the cores
        if (n_reads < 20000) nt = 1;
        std::vector<std::thread> th; int64_t per = (n_reads + nt - 1) / nt;
        for (unsigned t = 0; t < nt; t++) { int64_t r0 = t * per, r1 = std::min<int64_t>(n_reads, r0 + per); if (r0 < r1) th.emplace_back(work, r0, r1); }
        for (auto& t : th) t.join();
    }
};

struct DeviceBatch {
    DevBuf read_off, bases, quals, chain_off, chain_contig, chain_pos, chain_flag, chain_as, cigar_off, cigar, chain_order, read_primary, slot_read;
    DevBatch view{};
    // the caller's arrays first (asynchronous from pinned memory: they travel while the host sorts the chains), then what the host prepared
    void upload_raw(const hlala_seed_batch_t& b, cudaStream_t st) {
        int64_t nb = b.read_off[b.n_reads]; int32_t nc = b.chain_off[b.n_reads]; int32_t ncg = b.cigar_off[nc];
        read_off.upload(b.read_off, (size_t)b.n_reads + 1, st); bases.upload(b.bases, (size_t)nb, st); quals.upload(b.quals, (size_t)nb, st);
        chain_off.upload(b.chain_off, (size_t)b.n_reads + 1, st);
        chain_contig.upload(b.chain_contig, (size_t)nc, st); chain_pos.upload(b.chain_pos, (size_t)nc, st); chain_flag.upload(b.chain_flag, (size_t)nc, st); chain_as.upload(b.chain_as, (size_t)nc, st);
        cigar_off.upload(b.cigar_off, (size_t)nc + 1, st); cigar.upload(b.cigar, (size_t)ncg, st);
    }
    void upload(const hlala_seed_batch_t& b, const PreparedBatch& pb, cudaStream_t st) {
        int32_t nc = pb.n_chains;
        chain_order.upload(pb.chain_order, st); read_primary.upload(pb.read_primary, st); slot_read.upload(pb.slot_read, st);
        view.n_reads = b.n_reads; view.n_chains = nc;
        view.read_off = read_off.as<int64_t>(); view.bases = bases.as<uint8_t>(); view.quals = quals.as<uint8_t>();
        view.chain_off = chain_off.as<int32_t>(); view.chain_contig = chain_contig.as<int32_t>(); view.chain_pos = chain_pos.as<int32_t>();
        view.chain_flag = chain_flag.as<uint16_t>(); view.chain_as = chain_as.as<int32_t>(); view.cigar_off = cigar_off.as<int32_t>(); view.cigar = cigar.as<uint32_t>();
        view.chain_order = chain_order.as<int32_t>(); view.read_primary = read_primary.as<int32_t>(); view.slot_read = slot_read.as<int32_t>();
    }
};

int check_batch(const hlala_seed_batch_t* b) {
    if (!b || b->n_reads < 0 || (b->n_reads & 1)) return fail(HLALA_E_ARG, "seed batch: n_reads must be even and non-negative");
    if (!b->read_off || !b->chain_off || !b->cigar_off) return fail(HLALA_E_ARG, "seed batch: null offset arrays");
    return 0;
}

struct ChainScratch {   // per-chain results of the whole batch
    DevBuf status, n_cols, seed_begin, seed_end, ll, first_level, last_level, error_count, id_first, id_last;
    void alloc(int32_t n_chains) {
        size_t nc = (size_t)std::max(n_chains, 1);
        status.alloc(nc * 4); n_cols.alloc(nc * 4); seed_begin.alloc(nc * 4); seed_end.alloc(nc * 4); ll.alloc(nc * 8); first_level.alloc(nc * 4); last_level.alloc(nc * 4);
        error_count.alloc(4); id_first.alloc(nc * 4); id_last.alloc(nc * 4);
    }
    void fill(ChainParams& P) {
        P.status = status.as<int32_t>(); P.n_cols = n_cols.as<int32_t>(); P.seed_begin = seed_begin.as<int32_t>(); P.seed_end = seed_end.as<int32_t>(); P.ll = ll.as<double>();
        P.first_level = first_level.as<int32_t>(); P.last_level = last_level.as<int32_t>();
        P.error_count = error_count.as<int32_t>(); P.id_first = id_first.as<int32_t>(); P.id_last = id_last.as<int32_t>();
    }
};

// One in-flight wave: its column scratch, work lists, extension buffers, DP working memory and stream. Waves alternate between the
// lanes, so the low-occupancy tail of one wave's extension cascade overlaps the next wave's chain kernel and first DP tier.
struct Lane {
    DevBuf c_edge, c_schar, c_fromseed, pending_slots, pending_count, todo_slots, todo_count, defer_slots, defer_count, dp_tasks, dp_task_count, dp_task_bin, dp_task_hist, dp_sorted, dp_task_key, dp_key_sorted, dp_sort_temp;
    DevBuf pair_gslab[2];
    DevBuf pair_defer, pair_defer_count;   // pairs queued for the large tier of the pair kernel
    DevBuf ln_rec, ln_ahead;   // thread-per-extension tier: 16-byte cell records and the ahead table of every resident thread
    DevBuf ext_edge, ext_s, ext_n, ext_nlvl, ext_rc, dp_scratch, wd_scratch, gd_scratch; int32_t ext_cap = 0;
    DevBuf q_ctr, q_a, q_b;   // task queues of the extension cascade (counters, two ping-pong lists of deferred tasks)
    int32_t* n_pending_host = nullptr; cudaStream_t stream = nullptr; cudaEvent_t front_done = nullptr, done = nullptr;
    Lane() {}
    Lane(const Lane&) = delete; Lane& operator=(const Lane&) = delete;
    ~Lane() { if (n_pending_host) cudaFreeHost(n_pending_host); if (stream) cudaStreamDestroy(stream); if (front_done) cudaEventDestroy(front_done); if (done) cudaEventDestroy(done); }
    void alloc(int32_t wave_chains, int32_t maxcol, size_t dp_bytes, size_t wd_bytes, size_t gd_bytes, size_t ln_rec_bytes, size_t ln_ahead_bytes) {
        size_t wc = (size_t)std::max(wave_chains, 1);
        c_edge.alloc(wc * maxcol * 4); c_schar.alloc(wc * maxcol); c_fromseed.alloc(wc * maxcol);
        pending_slots.alloc(wc * 4); pending_count.alloc(4); todo_slots.alloc(wc * 4); todo_count.alloc(4); defer_slots.alloc(wc * 4); defer_count.alloc(4);
        dp_scratch.alloc(dp_bytes); wd_scratch.alloc(wd_bytes); gd_scratch.alloc(gd_bytes); q_ctr.alloc(128);
        pair_defer.alloc(wc * 2 + 64); pair_defer_count.alloc(4);     // at most one entry per pair; a pair has >= 2 chains
        dp_tasks.alloc(wc * 8); dp_task_key.alloc(wc * 8); dp_key_sorted.alloc(wc * 8); dp_task_count.alloc(4); dp_task_bin.alloc(wc * 2); dp_task_hist.alloc(1024); dp_sorted.alloc(wc * 8); ln_rec.alloc(ln_rec_bytes); ln_ahead.alloc(ln_ahead_bytes);
        if (!n_pending_host) { CUDA_OK(cudaMallocHost((void**)&n_pending_host, 4)); *n_pending_host = 0; }
        if (!stream) CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (!front_done) { CUDA_OK(cudaEventCreateWithFlags(&front_done, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming)); }
    }
    void fill(ChainParams& P) {
        P.c_edge = c_edge.as<int32_t>(); P.c_schar = c_schar.as<uint8_t>(); P.c_fromseed = c_fromseed.as<uint8_t>();
        P.pending_slots = pending_slots.as<int32_t>(); P.pending_count = pending_count.as<int32_t>();
        P.todo_slots = todo_slots.as<int32_t>(); P.todo_count = todo_count.as<int32_t>();
        P.defer_slots = defer_slots.as<int32_t>(); P.defer_count = defer_count.as<int32_t>();
        P.dp_tasks = dp_tasks.as<int32_t>(); P.dp_task_count = dp_task_count.as<int32_t>(); P.dp_task_bin = dp_task_bin.as<uint8_t>(); P.dp_task_hist = dp_task_hist.as<int32_t>(); P.dp_task_key = dp_task_key.as<int32_t>();
    }
};

// Shared-memory slab of one warp of the chain kernels. The backtrack pool holds one entry per (column, node of the column's level):
// max(4 x max_columns, 2560) entries cover a 150-column read inside a 14-node-wide gene block; a chain that needs more gets
// HLALA_E_CAPACITY. The staged edge window falls back to L2 reads when the window is larger than its capacity.
void chain_caps(int32_t maxcol, bool bt16, int tier, ChainParams& P) {
    P.maxcol = maxcol; P.bt16 = bt16 ? 1 : 0; P.tier = tier;
    if (tier == 0) {   // small slab (~10 KB per warp): covers ~99.9 % of short-read chains at ~20 resident warps per SM
        P.slab_cols = std::min<int32_t>(maxcol, 256); P.wcap = 64;
        long long pool_entries = 1024; P.pool_cap = (int32_t)(bt16 ? pool_entries / 2 : pool_entries); P.win_cap = 512;
        return;
    }
    P.slab_cols = maxcol; P.wcap = K1_WCAP;
    long long pool_entries = std::max<long long>(4LL * maxcol, 2560);       // entries needed
    P.pool_cap = (int32_t)(bt16 ? (pool_entries + 1) / 2 : pool_entries);    // in 32-bit words when 16-bit entries are used
    P.win_cap = (int32_t)((std::max<long long>(2LL * maxcol, 512) + 3) & ~3LL);     // a multiple of 4 words (bulk copies are 16-byte granular)
    if (maxcol > 2040) return;       // the slab of such chains lives in HBM (Pipeline::chain_params)
    while (k1_slab_bytes(P.slab_cols, P.pool_cap, P.win_cap, P.wcap) * K1_WARPS > 220000 && P.win_cap > 256) P.win_cap /= 2;
    while (k1_slab_bytes(P.slab_cols, P.pool_cap, P.win_cap, P.wcap) * K1_WARPS > 220000 && P.pool_cap > 512) P.pool_cap /= 2;
}

// boost::math::pdf(normal) as Boost.Math computes it (processBAM.cpp:2342-2346, 3446-3472)
double normal_pdf(double mean, double sd, double x) {
    double exponent = x - mean; exponent *= -exponent; exponent /= 2 * sd * sd;
    double result = exp(exponent); result /= sd * sqrt(2 * 3.14159265358979323846264338327950288);
    return result;
}
// Utilities::PCorrectToPhred (Utilities.cpp:178-203)
int pcorrect_to_phred(double PCorrect) {
    double pWrong = 1 - PCorrect; if (pWrong == 0) pWrong = 1e-100;
    double phred1 = -10.0 * log10(pWrong);
    if ((phred1 + 33) > 255) phred1 = 255 - 33;
    return (int)round(phred1 + 33);
}
// thr[k] = smallest double q in [0,1] with PCorrectToPhred(q) >= 33 + k (the function is monotone in q)
std::vector<double> phred_thresholds() {
    std::vector<double> thr(223, 0.0);
    for (int k = 1; k <= 222; k++) {
        uint64_t lo = 0, hi; double one = 1.0; memcpy(&hi, &one, 8);     // bit patterns of non-negative doubles are ordered
        if (pcorrect_to_phred(1.0) < 33 + k) { thr[k] = 2.0; continue; }
        while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; double q; memcpy(&q, &mid, 8); if (pcorrect_to_phred(q) >= 33 + k) hi = mid; else lo = mid + 1; }
        memcpy(&thr[k], &lo, 8);
    }
    return thr;
}

struct Pipeline {
    hlala_graph* g = nullptr; int32_t maxcol = 0;
    PreparedBatch pb; DeviceBatch db; ChainScratch cs;
    std::vector<std::unique_ptr<Lane>> lanes; int n_lanes_max = 3; cudaEvent_t fork_ev = nullptr;
    int32_t n_dp_threads = 0; int32_t n_wd_warps = 0;
    ~Pipeline() { if (fork_ev) cudaEventDestroy(fork_ev); }
    bool scalar_dp_only = false;   // test hook: run every extension through the scalar kernel
    int32_t n_gd_groups = 0; bool group_dp = true; bool dp_trace = false;
    int32_t ln_batch = 16; bool lean_big = false; bool sort_by_level = true;      // extension tasks in level order (HLALA_DP_SORT=clip: by clipped length, longest first)
    bool long_mode = false; DevBuf gslab; size_t gslab_bytes = 0;     // long-read mode (alignOneLongRead): one read per unit, no extension DP, chain buffers in HBM
    int32_t n_ln_threads = 0; bool lean_dp = true;   // first DP tier: one thread per extension (extend_lean.h); false: the 8-lane group kernel (A/B runs)
    DevBuf bpl_ws; std::vector<int32_t> bpl_host;   // per-level coverage of a host-buffer call
    DevBuf is_table, phred_thr; double is_mean = -1, is_sd = -1, is_pen = 0; int32_t is_dmin = 0, is_n = 0;
    DevBuf pair_mapq, read_mapq, read_reverse, chosen_slot, pair_ll, pair_status, digest;
    DevBuf o_n_cols, o_level, o_edge, o_gchar, o_schar, o_fromseed, o_mapq; bool have_columns = false;
    int launches = 0; int64_t algo_bytes = 0; int32_t n_errors = 0;
    bool timing = false; std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timed;   // (kernel class, start/stop)
    int64_t chain_kernel_bytes = 0;
    DevBuf dp_bytes;   // algorithmic bytes of the extension tasks of the timed runs (k_dp_task_bytes)
    void tic(int cls, cudaStream_t st) { if (!timing) return; cudaEvent_t a, b; CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b)); CUDA_OK(cudaEventRecord(a, st)); timed.push_back({cls, {a, b}}); }
    void toc(cudaStream_t st) { if (!timing) return; CUDA_OK(cudaEventRecord(timed.back().second.second, st)); }
    void collect_timing(double ms[6], int launches_per_class[6]) {
        for (int i = 0; i < 6; i++) { ms[i] = 0; launches_per_class[i] = 0; }
        for (auto& t : timed) { float f = 0; CUDA_OK(cudaEventSynchronize(t.second.second)); CUDA_OK(cudaEventElapsedTime(&f, t.second.first, t.second.second)); ms[t.first] += f; launches_per_class[t.first]++; cudaEventDestroy(t.second.first); cudaEventDestroy(t.second.second); }
        timed.clear();
    }
    std::vector<int64_t> wave_pair;    // wave w covers pairs [wave_pair[w], wave_pair[w+1])
    size_t scratch_budget = 0;   // bytes of per-chain column scratch per wave; 0 = from the device's memory (see prepare)
    bool allow_env_budget = true;
    bool dedup = true;    // false: align every same-strand chain (parity tests of the chain kernels)

    void prepare(hlala_graph* graph, const hlala_seed_batch_t& b, int32_t mc, cudaStream_t st) {
        g = graph; maxcol = mc; have_columns = false;
        if (scratch_budget == 0) {
            // The extension tiers are persistent grids that pop tasks from a queue: the more tasks a launch holds, the smaller the share of its
            // tail (measured on B200, 1 M pairs: one wave 445 ms, ten waves on three lanes 507 ms). So: one wave if the column scratch of the
            // whole batch fits 40 % of the device memory, else three lanes sharing that much.
            size_t free_b = 0, total_b = 0; CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
            const size_t avail = std::min(total_b / 5 * 2, free_b / 5 * 3), need = (size_t)std::max(b.chain_off[b.n_reads], 1) * (size_t)mc * 6;     // never more than 60 % of what is free right now
            scratch_budget = need <= avail ? need : avail / 3;
        }
        if (const char* e = allow_env_budget ? getenv("HLALA_WAVE_BYTES") : nullptr) scratch_budget = (size_t)strtoull(e, nullptr, 10);   // test hook: force several waves
        db.upload_raw(b, st); pb.build(b); db.upload(b, pb, st); host_chain_off.assign(b.chain_off, b.chain_off + b.n_reads + 1); host_read_off.assign(b.read_off, b.read_off + b.n_reads + 1);
        // waves: consecutive pairs whose chains fit the column-scratch budget
        wave_pair.assign(1, 0); int32_t max_wave_chains = 0;
        { const int64_t np = b.n_reads / 2; const int64_t cap = std::max<int64_t>(1024, (int64_t)(scratch_budget / ((size_t)mc * 6)));
          int64_t p0 = 0;
          while (p0 < np) {
              int64_t p1 = p0; int32_t c0 = b.chain_off[2 * p0];
              while (p1 < np && (int64_t)(b.chain_off[2 * (p1 + 1)] - c0) <= cap) p1++;
              if (p1 == p0) p1 = p0 + 1;
              max_wave_chains = std::max(max_wave_chains, b.chain_off[2 * p1] - c0);
              wave_pair.push_back(p1); p0 = p1;
          } }
        cs.alloc(pb.n_chains);
        size_t nr = (size_t)std::max<int64_t>(b.n_reads, 2), np = nr / 2;
        pair_mapq.alloc(np * 8); read_mapq.alloc(nr * 8); read_reverse.alloc(nr); chosen_slot.alloc(nr * 4); pair_ll.alloc(np * 8); pair_status.alloc(np * 4); digest.alloc(32);
        phred_thr.upload(phred_thresholds(), st);
        if (getenv("HLALA_SCALAR_DP")) scalar_dp_only = true;            // test hook (tools/scale_parity.py): the scalar DP only
        n_dp_threads = g->n_sm * (scalar_dp_only ? 64 : 8);     // otherwise the scalar kernel is the last resort of the cascade (a handful of tasks per wave)
        n_wd_warps = wd_warps_for(g->n_sm);
        n_gd_groups = gd_groups_for(g->n_sm);
        n_ln_threads = ln_threads_for_any(g->n_sm);
        lean_big = getenv("HLALA_LEAN_BIG") != nullptr;                  // experiment: a second thread-per-extension pass with doubled capacities before the warp tiers
        if (getenv("HLALA_NO_LEAN_DP")) lean_dp = false;                 // test hook: the round-1 cascade (8-lane groups first)
        if (const char* e = allow_env_budget ? getenv("HLALA_LANES") : nullptr) n_lanes_max = std::max(1, atoi(e));   // test hook
        const int n_lanes = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_lanes_max, wave_pair.size() - 1));
        while ((int)lanes.size() > n_lanes) lanes.pop_back();
        for (int i = 0; i < n_lanes; i++) {
            if ((int)lanes.size() <= i) lanes.emplace_back(new Lane());
            Lane& L = *lanes[(size_t)i];
            L.alloc(max_wave_chains, mc, (size_t)n_dp_threads * dp_thread_scratch_bytes(), (size_t)n_wd_warps * wd_warp_scratch_bytes(), (size_t)n_gd_groups * gd_group_scratch_bytes(),
                    (size_t)n_ln_threads * ln_thread_rec_bytes(), (size_t)n_ln_threads * ln_thread_ahead_bytes());
            // DP working memory is zeroed once; its hashes are generation-stamped and reused without clearing afterwards
            if (L.dp_scratch.fresh) CUDA_OK(cudaMemsetAsync(L.dp_scratch.p, 0, L.dp_scratch.bytes, st));
            if (L.wd_scratch.fresh) CUDA_OK(cudaMemsetAsync(L.wd_scratch.p, 0, L.wd_scratch.bytes, st));
            if (L.gd_scratch.fresh) CUDA_OK(cudaMemsetAsync(L.gd_scratch.p, 0, L.gd_scratch.bytes, st));
        }
        if (!fork_ev) CUDA_OK(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
        CUDA_OK(cudaStreamSynchronize(st));   // the lanes' own streams start from initialised scratch
        if (getenv("HLALA_DP_TRACE")) dp_trace = true;
        if (const char* e = getenv("HLALA_DP_SORT")) sort_by_level = strcmp(e, "clip") != 0;
        if (const char* e = getenv("HLALA_LN_BATCH")) ln_batch = std::max(1, std::min(32, atoi(e)));
        if (getenv("HLALA_NO_GROUP_DP")) group_dp = false;               // test hook: first tier = warp kernel, tiny configuration
        if (allow_env_budget && getenv("HLALA_ALIGN_DUPLICATES")) dedup = false;   // align the chains k_prepare would skip
    }
    // algorithmic bytes: bases+quals, seed records + CIGARs, translation + graph window per chain column,
    // chosen alignment columns written once, per-pair scalars, coverage RMW. Only the bench asks for them.
    void compute_algorithmic_bytes(const hlala_seed_batch_t& b) {
        int64_t nb = b.read_off[b.n_reads]; int64_t ncg = b.cigar_off[pb.n_chains];
        int64_t cols = 0; for (int32_t c = 0; c < pb.n_chains; c++) for (int32_t k = b.cigar_off[c]; k < b.cigar_off[c + 1]; k++) { int op = b.cigar[k] & 15; if (op == 0 || op == 1 || op == 2 || op == 7 || op == 8) cols += b.cigar[k] >> 4; }
        algo_bytes = 2 * nb + 24ll * pb.n_chains + 4 * ncg + cols * (4 + 7) + 2 * nb * 12 / 2 + 40ll * (b.n_reads / 2) + 4 * nb;
        // chain kernel alone ("Kernel 1"): per chain record+CIGAR 24+4*ops, read bases+quals once per chain (2L), per column
        // translation 4 + contig base 1 + graph window (edge_pack 4*1.5 + level offsets 8 + gap flag 1) + output (edge 4 + read char 1 + seed flag 1), 44 B scalars
        { int64_t rb = 0; for (int64_t r = 0; r < b.n_reads; r++) rb += (b.read_off[r + 1] - b.read_off[r]) * (int64_t)(b.chain_off[r + 1] - b.chain_off[r]);
          chain_kernel_bytes = 24ll * pb.n_chains + 4 * ncg + 2 * rb + cols * (4 + 1 + 6 + 8 + 1 + 6) + 44ll * pb.n_chains; }
        // per input chain, for aligned_bytes(): columns, CIGAR operations, read length
        cb_cols.assign((size_t)pb.n_chains, 0); cb_ops.assign((size_t)pb.n_chains, 0); cb_len.assign((size_t)pb.n_chains, 0);
        for (int64_t r = 0; r < b.n_reads; r++) for (int32_t c = b.chain_off[r]; c < b.chain_off[r + 1]; c++) {
            int32_t cc = 0; for (int32_t k = b.cigar_off[c]; k < b.cigar_off[c + 1]; k++) { int op = b.cigar[k] & 15; if (op == 0 || op == 1 || op == 2 || op == 7 || op == 8) cc += (int32_t)(b.cigar[k] >> 4); }
            cb_cols[(size_t)c] = cc; cb_ops[(size_t)c] = b.cigar_off[c + 1] - b.cigar_off[c]; cb_len[(size_t)c] = (int32_t)(b.read_off[r + 1] - b.read_off[r]); }
        algo_fixed = 2 * nb + 2 * nb * 12 / 2 + 40ll * (b.n_reads / 2) + 4 * nb;
    }
    std::vector<int32_t> cb_cols, cb_ops, cb_len; int64_t algo_fixed = 0;
    // The same two figures counted over the chains the run ALIGNED (status 0) only: k_prepare marks chains the pair stage would discard as duplicates
    // and chains on the other strand, and no kernel touches their CIGARs, columns or graph windows. out: {chains aligned, chain kernel bytes, whole path bytes}
    void aligned_bytes(int64_t out[3]) {
        std::vector<int32_t> st((size_t)std::max(pb.n_chains, 1)); cs.status.download(st.data(), (size_t)pb.n_chains, 0); CUDA_OK(cudaStreamSynchronize(0));
        int64_t n = 0, ck = 0, chains_part = 0;
        for (int32_t s = 0; s < pb.n_chains; s++) { if (st[(size_t)s] != CH_OK) continue; const size_t c = (size_t)pb.chain_order[(size_t)s]; n++;
            ck += 24 + 4ll * cb_ops[c] + 2ll * cb_len[c] + (int64_t)cb_cols[c] * (4 + 1 + 6 + 8 + 1 + 6) + 44;
            chains_part += 24 + 4ll * cb_ops[c] + (int64_t)cb_cols[c] * (4 + 7); }
        out[0] = n; out[1] = ck; out[2] = algo_fixed + chains_part;
    }
    // defaults of everything the test hooks (environment variables read in prepare) can change; a workspace kept across calls starts from them
    void reset_config() { ln_batch = 16; sort_by_level = true; long_mode = false; scratch_budget = 0; allow_env_budget = true; dedup = true; n_lanes_max = 3; scalar_dp_only = false; group_dp = true; lean_dp = true; dp_trace = false; }
    void ensure_columns() {
        if (have_columns) return;
        size_t n = (size_t)std::max<int64_t>(pb.n_reads, 2) * maxcol;
        o_n_cols.alloc((size_t)std::max<int64_t>(pb.n_reads, 2) * 4); o_level.alloc(n * 4); o_edge.alloc(n * 4); o_gchar.alloc(n); o_schar.alloc(n); o_fromseed.alloc(n); o_mapq.alloc(n);
        have_columns = true;
    }
    ChainParams chain_params(int tier, Lane& L) {
        ChainParams P{}; P.g = g->d; P.b = db.view; chain_caps(maxcol, g->h.max_edges_per_level <= 255 && g->h.max_nodes_per_level <= 256, tier, P); P.do_extension = long_mode ? 0 : 1; P.long_mode = long_mode ? 1 : 0; cs.fill(P); L.fill(P);
        {   // score field of the packed Viterbi keys: 12 bits (key_shift 20) cover max_columns <= 2040; longer chains take bits from the rank field, which must still hold the widest level's edges
            int ks = 20; while (ks > 12 && (long long)maxcol + 2 >= (1LL << (32 - ks))) ks--;
            if ((long long)maxcol + 2 >= (1LL << (32 - ks)) || (long long)g->h.max_edges_per_level >= (1LL << ks)) throw std::runtime_error("max_columns " + std::to_string(maxcol) + " with levels of " + std::to_string(g->h.max_edges_per_level) + " edges exceeds the packed Viterbi key of the chain kernel");
            P.key_shift = ks;
        }
        if (tier == 1 && maxcol > 2040) {     // chains of thousands of columns: slab in HBM, one slice per resident warp
            P.gslab = 1; P.win_cap = 4096; P.gslab_bytes = k1_gslab_bytes(P.slab_cols, P.pool_cap);
            const int warps = chain_seed_max_warps(P, g->n_sm); gslab.alloc((size_t)std::max(warps, 1) * P.gslab_bytes); P.gslab_base = gslab.as<unsigned char>();
        }
        return P;
    }

    void begin_run(cudaStream_t st) {
        launches = 0;
        CUDA_OK(cudaMemsetAsync(cs.error_count.p, 0, 4, st));
        CUDA_OK(cudaMemsetAsync(cs.status.p, 0xFF, (size_t)std::max(pb.n_chains, 1) * 4, st));
        CUDA_OK(cudaMemsetAsync(digest.p, 0, 32, st));
    }
    // chain stage for the slots of one wave, first half: strand/duplicate rules, seed projection; leaves the number of chains that
    // need an extension in the lane's pinned counter once front_done has fired
    ChainParams wave_front(size_t w, Lane& L) {
        cudaStream_t st = L.stream;
        ChainParams P0 = chain_params(0, L), P = chain_params(1, L);
        for (ChainParams* q : {&P0, &P}) {
            q->slot_base = db_chain_off(2 * wave_pair[w]); q->slot_end = db_chain_off(2 * wave_pair[w + 1]);
            q->read_begin = (int32_t)(2 * wave_pair[w]); q->read_end = (int32_t)(2 * wave_pair[w + 1]); q->dedup = dedup ? 1 : 0;
        }
        CUDA_OK(cudaMemsetAsync(L.pending_count.p, 0, 4, st)); CUDA_OK(cudaMemsetAsync(L.todo_count.p, 0, 4, st)); CUDA_OK(cudaMemsetAsync(L.defer_count.p, 0, 4, st)); CUDA_OK(cudaMemsetAsync(L.dp_task_count.p, 0, 4, st)); CUDA_OK(cudaMemsetAsync(L.dp_task_hist.p, 0, 1024, st));
        if (sort_by_level) CUDA_OK(cudaMemsetAsync(L.dp_task_key.p, 0x7f, L.dp_task_key.bytes, st));
        if (P.slot_end > P.slot_base) {
            CUDA_OK(launch_prepare(P0, st));
            tic(0, st); CUDA_OK(launch_chain_seed(P0, g->n_sm, st)); CUDA_OK(launch_chain_seed(P, g->n_sm, st)); toc(st); launches += 3;
        }
        CUDA_OK(cudaMemcpyAsync(L.n_pending_host, L.pending_count.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaEventRecord(L.front_done, st));
        return P;
    }
    // second half: extension cascade + chain finish
    void wave_back(const ChainParams& P, Lane& L) {
        cudaStream_t st = L.stream;
        CUDA_OK(cudaEventSynchronize(L.front_done));
        const int32_t n_pending = *L.n_pending_host;
        if (n_pending > 0) {
            if (n_pending > L.ext_cap) {
                CUDA_OK(cudaStreamSynchronize(st));   // an earlier wave of this lane may still read the old buffers
                L.ext_cap = n_pending + n_pending / 4 + 64; size_t nt = (size_t)L.ext_cap * 2;
                L.ext_edge.alloc(nt * dp_ext_cap() * 4); L.ext_s.alloc(nt * dp_ext_cap()); L.ext_n.alloc(nt * 4); L.ext_nlvl.alloc(nt * 4); L.ext_rc.alloc(nt * 4); L.q_a.alloc(nt * 4); L.q_b.alloc(nt * 4);
            }
            ExtParams E{}; E.C = P; E.n_pending = n_pending; E.ext_edge = L.ext_edge.as<int32_t>(); E.ext_s = L.ext_s.as<uint8_t>(); E.ext_n = L.ext_n.as<int32_t>();
            E.ext_nlvl = L.ext_nlvl.as<int32_t>(); E.ext_rc = L.ext_rc.as<int32_t>(); E.dp_scratch = L.dp_scratch.as<unsigned char>(); E.n_dp_threads = n_dp_threads;
            E.wd_scratch = L.wd_scratch.as<unsigned char>(); E.n_wd_warps = n_wd_warps; E.gd_scratch = L.gd_scratch.as<unsigned char>(); E.n_gd_groups = n_gd_groups;
            E.ln_rec = L.ln_rec.as<LnRec>(); E.ln_ahead = L.ln_ahead.as<uint32_t>(); E.n_ln_threads = n_ln_threads; E.ln_batch = ln_batch;
            if (timing) { if (!dp_bytes.p) { dp_bytes.alloc(8); CUDA_OK(cudaMemsetAsync(dp_bytes.p, 0, 8, st)); } CUDA_OK(launch_dp_task_bytes(E, dp_bytes.as<unsigned long long>(), st)); }
            tic(1, st);
            if (scalar_dp_only) { E.only_deferred = 0; CUDA_OK(launch_extend(E, st)); launches += 1; }
            else {
                // cascade: 8-lane groups (every task) -> warp kernels with growing shared-memory capacities -> scalar kernel; each tier only
                // re-runs what the previous one deferred for capacity
                auto trace = [&](const char* what, cudaEvent_t a) {   // HLALA_DP_TRACE=1: time and deferral count of each tier (debug runs only)
                    cudaEvent_t b; CUDA_OK(cudaEventCreate(&b)); CUDA_OK(cudaEventRecord(b, st)); CUDA_OK(cudaEventSynchronize(b)); float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, a, b));
                    std::vector<int32_t> rc((size_t)2 * n_pending); L.ext_rc.download(rc.data(), rc.size(), st); CUDA_OK(cudaStreamSynchronize(st));
                    long long nd = 0; for (int32_t v : rc) nd += (v == -100);
                    fprintf(stderr, "[dp-trace] %-12s %8.2f ms, tasks %d, still deferred %lld\n", what, ms, 2 * n_pending, nd); cudaEventDestroy(b); CUDA_OK(cudaEventRecord(a, st)); };
                cudaEvent_t tr0 = nullptr; if (dp_trace) { CUDA_OK(cudaEventCreate(&tr0)); CUDA_OK(cudaEventRecord(tr0, st)); }
                int32_t* ctr = L.q_ctr.as<int32_t>(); int32_t* qa = L.q_a.as<int32_t>(); int32_t* qb = L.q_b.as<int32_t>();
                CUDA_OK(cudaMemsetAsync(ctr, 0, 128, st));
                if (lean_dp) {
                    // thread-per-extension tiers over the list of tasks that run; the other sides of the pending chains have no extension
                    const size_t nt = (size_t)2 * n_pending;
                    CUDA_OK(cudaMemsetAsync(E.ext_rc, 0, nt * 4, st)); CUDA_OK(cudaMemsetAsync(E.ext_n, 0, nt * 4, st)); CUDA_OK(cudaMemsetAsync(E.ext_nlvl, 0, nt * 4, st));
                    if (sort_by_level) {
                        const int n_sort = (int)std::min<size_t>(nt, L.dp_task_key.bytes / 4); size_t need = 0;
                        CUDA_OK(sort_dp_tasks_by_level(L.dp_task_key.as<int32_t>(), L.dp_key_sorted.as<int32_t>(), P.dp_tasks, L.dp_sorted.as<int32_t>(), n_sort, nullptr, 0, &need, st));
                        if (need > L.dp_sort_temp.bytes) { CUDA_OK(cudaStreamSynchronize(st)); L.dp_sort_temp.alloc(need + 256); }
                        CUDA_OK(sort_dp_tasks_by_level(L.dp_task_key.as<int32_t>(), L.dp_key_sorted.as<int32_t>(), P.dp_tasks, L.dp_sorted.as<int32_t>(), n_sort, L.dp_sort_temp.p, L.dp_sort_temp.bytes, &need, st));
                    } else CUDA_OK(launch_sort_dp_tasks(P, L.dp_sorted.as<int32_t>(), st));
                    E.in_list = L.dp_sorted.as<int32_t>(); E.in_count = L.dp_task_count.as<int32_t>(); E.pop = ctr + 8; E.out_list = qa; E.out_count = ctr + 9;
                    CUDA_OK(launch_extend_lean(E, g->n_sm, 0, st)); if (dp_trace) trace("lean std", tr0);
                    launches += 1;
                    if (lean_big) {
                        E.in_list = qa; E.in_count = ctr + 9; E.pop = ctr + 10; E.out_list = qb; E.out_count = ctr + 11;
                        CUDA_OK(launch_extend_lean(E, g->n_sm, 1, st)); launches += 1; if (dp_trace) trace("lean big", tr0);
                    }
                    toc(st);
                    tic(4, st);
                    E.in_list = lean_big ? qb : qa; E.in_count = lean_big ? ctr + 11 : ctr + 9; E.pop = ctr + 12; E.out_list = lean_big ? qa : qb; E.out_count = ctr + 13;
                    if (!lean_big) { CUDA_OK(launch_extend_warp(E, g->n_sm, 0, true, st)); if (dp_trace) trace("warp tiny", tr0);
                        E.in_list = qb; E.in_count = ctr + 13; E.pop = ctr + 14; E.out_list = qa; E.out_count = ctr + 15;
                        CUDA_OK(launch_extend_warp(E, g->n_sm, 1, true, st)); if (dp_trace) trace("warp small", tr0);
                        E.in_list = qa; E.in_count = ctr + 15; E.pop = ctr + 16; E.out_list = qb; E.out_count = ctr + 17;
                        CUDA_OK(launch_extend_warp(E, g->n_sm, 2, true, st)); if (dp_trace) trace("warp large", tr0); toc(st);
                        tic(5, st); E.only_deferred = 1; CUDA_OK(launch_extend(E, st)); launches += 4; if (dp_trace) { trace("scalar", tr0); cudaEventDestroy(tr0); }
                    } else {
                    CUDA_OK(launch_extend_warp(E, g->n_sm, 0, true, st)); if (dp_trace) trace("warp tiny", tr0);
                    E.in_list = qa; E.in_count = ctr + 13; E.pop = ctr + 14; E.out_list = qb; E.out_count = ctr + 15;
                    CUDA_OK(launch_extend_warp(E, g->n_sm, 1, true, st)); if (dp_trace) trace("warp small", tr0);
                    E.in_list = qb; E.in_count = ctr + 15; E.pop = ctr + 16; E.out_list = qa; E.out_count = ctr + 17;
                    CUDA_OK(launch_extend_warp(E, g->n_sm, 2, true, st)); if (dp_trace) trace("warp large", tr0); toc(st);
                    tic(5, st); E.only_deferred = 1; CUDA_OK(launch_extend(E, st)); launches += 4; if (dp_trace) { trace("scalar", tr0); cudaEventDestroy(tr0); }
                    }
                } else {
                if (group_dp) { E.in_list = nullptr; E.in_count = nullptr; E.pop = ctr + 0; E.out_list = qa; E.out_count = ctr + 1; CUDA_OK(launch_extend_group(E, g->n_sm, st)); launches += 1; if (dp_trace) trace("group8", tr0); }
                toc(st);
                tic(4, st);
                E.in_list = group_dp ? qa : nullptr; E.in_count = ctr + 1; E.pop = ctr + 2; E.out_list = qb; E.out_count = ctr + 3;
                CUDA_OK(launch_extend_warp(E, g->n_sm, 0, group_dp, st)); if (dp_trace) trace("warp tiny", tr0);
                E.in_list = qb; E.in_count = ctr + 3; E.pop = ctr + 4; E.out_list = qa; E.out_count = ctr + 5;
                CUDA_OK(launch_extend_warp(E, g->n_sm, 1, true, st)); if (dp_trace) trace("warp small", tr0);
                E.in_list = qa; E.in_count = ctr + 5; E.pop = ctr + 6; E.out_list = qb; E.out_count = ctr + 7;
                CUDA_OK(launch_extend_warp(E, g->n_sm, 2, true, st)); if (dp_trace) trace("warp large", tr0); toc(st);
                tic(5, st); E.only_deferred = 1; CUDA_OK(launch_extend(E, st)); launches += 4; if (dp_trace) { trace("scalar", tr0); cudaEventDestroy(tr0); }
                }
            }
            toc(st);
            tic(2, st); CUDA_OK(launch_chain_finish(E, g->n_sm, st)); toc(st); launches += 1;
        }
    }
    ChainParams run_chains_wave(size_t w, Lane& L) { ChainParams P = wave_front(w, L); wave_back(P, L); return P; }
    std::vector<int32_t> host_chain_off;
    std::vector<int64_t> host_read_off;
    bool keep_columns = false;         // session runs materialise the chosen alignments' columns (typing stage input)
    int32_t db_chain_off(int64_t r) const { return host_chain_off[(size_t)r]; }
    void set_insert_size(double mean, double sd, cudaStream_t st) {
        if (mean == is_mean && sd == is_sd) return;
        if (!(sd > 0)) throw std::runtime_error("insert size sd must be positive");
        double pen = normal_pdf(mean, sd, mean + 8 * sd);
        if (!(pen > 0 && pen <= 1)) throw std::runtime_error("insert size penalty out of range (the reference asserts 0 < pdf <= 1)");
        is_pen = log(pen);
        long long lo = (long long)floor(mean - 40 * sd) - 2, hi = (long long)ceil(mean + 40 * sd) + 2;
        if (hi - lo + 1 > 4000000) throw std::runtime_error("insert size table too large");
        std::vector<double> t((size_t)(hi - lo + 1));
        for (long long d = lo; d <= hi; d++) { double p = normal_pdf(mean, sd, (double)d); t[(size_t)(d - lo)] = (p <= 0) ? is_pen : log(p); }
        if (normal_pdf(mean, sd, (double)(lo - 1)) > 0 || normal_pdf(mean, sd, (double)(hi + 1)) > 0) throw std::runtime_error("insert size table does not cover the support of the density");
        is_table.upload(t, st); is_dmin = (int32_t)lo; is_n = (int32_t)t.size(); is_mean = mean; is_sd = sd;
    }
    void run_pairs_wave(size_t w, const ChainParams& C, Lane& L, int32_t* bases_per_level_dev, bool want_columns) {
        cudaStream_t st = L.stream;
        PairParams Q{}; Q.g = g->d; Q.b = db.view; Q.maxcol = maxcol;
        Q.slot_base = C.slot_base; Q.pair_begin = wave_pair[w]; Q.pair_end = wave_pair[w + 1];
        Q.status = cs.status.as<int32_t>(); Q.n_cols = cs.n_cols.as<int32_t>(); Q.ll = cs.ll.as<double>(); Q.first_level = cs.first_level.as<int32_t>(); Q.last_level = cs.last_level.as<int32_t>();
        Q.id_first = cs.id_first.as<int32_t>(); Q.id_last = cs.id_last.as<int32_t>(); Q.c_edge = L.c_edge.as<int32_t>(); Q.c_schar = L.c_schar.as<uint8_t>(); Q.c_fromseed = L.c_fromseed.as<uint8_t>();
        Q.is_table = is_table.as<double>(); Q.is_dmin = is_dmin; Q.is_n = is_n; Q.is_penalty = is_pen; Q.phred_thr = phred_thr.as<double>();
        Q.pair_mapq = pair_mapq.as<double>(); Q.read_mapq = read_mapq.as<double>(); Q.read_reverse = read_reverse.as<uint8_t>(); Q.chosen_slot = chosen_slot.as<int32_t>();
        Q.pair_ll = pair_ll.as<double>(); Q.pair_status = pair_status.as<int32_t>();
        if (want_columns) { Q.out_n_cols = o_n_cols.as<int32_t>(); Q.out_level = o_level.as<int32_t>(); Q.out_edge = o_edge.as<int32_t>(); Q.out_gchar = o_gchar.as<uint8_t>(); Q.out_schar = o_schar.as<uint8_t>(); Q.out_fromseed = o_fromseed.as<uint8_t>(); Q.out_mapq = o_mapq.as<uint8_t>(); }
        Q.bases_per_level = bases_per_level_dev; Q.error_count = cs.error_count.as<int32_t>(); Q.digest = digest.as<unsigned long long>();
        Q.defer_list = L.pair_defer.as<int32_t>(); Q.defer_count = L.pair_defer_count.as<int32_t>(); Q.unpaired = long_mode ? 1 : 0;
        for (int t = 0; t < 2; t++) {     // thousands of columns per read: the per-warp pair slabs do not fit shared memory
            const size_t per_warp = pair_gslab_bytes(maxcol, t); Q.gslab[t] = nullptr;
            if (per_warp) { const size_t warps = (size_t)g->n_sm * 2 * K3_WARPS; L.pair_gslab[t].alloc(warps * per_warp); Q.gslab[t] = L.pair_gslab[t].as<unsigned char>(); Q.gslab_warps = warps; }
        }
        if (Q.pair_end > Q.pair_begin) { CUDA_OK(cudaMemsetAsync(Q.defer_count, 0, 4, st)); tic(3, st); CUDA_OK(launch_pair(Q, g->n_sm, st)); toc(st); launches += 2; }
    }
    // lanes start after everything queued on the caller's stream and the caller's stream continues after the lanes
    void fork_lanes(cudaStream_t st) { CUDA_OK(cudaEventRecord(fork_ev, st)); for (auto& L : lanes) CUDA_OK(cudaStreamWaitEvent(L->stream, fork_ev, 0)); }
    void join_lanes(cudaStream_t st) { for (auto& L : lanes) { CUDA_OK(cudaEventRecord(L->done, L->stream)); CUDA_OK(cudaStreamWaitEvent(st, L->done, 0)); } }
    // the whole path: per wave chain stage (+ extension) then pair stage; consecutive waves run on different lanes
    void run(double mean, double sd, int32_t* bases_per_level_dev, bool want_columns, cudaStream_t st) {
        set_insert_size(mean, sd, st);
        if (want_columns) ensure_columns();
        begin_run(st);
        fork_lanes(st);
        for (size_t w = 0; w + 1 < wave_pair.size(); w++) { Lane& L = *lanes[w % lanes.size()]; ChainParams C = run_chains_wave(w, L); run_pairs_wave(w, C, L, bases_per_level_dev, want_columns); }
        join_lanes(st);
    }
    void fetch(hlala_pair_out_t* out, cudaStream_t st) {
        size_t nr = (size_t)pb.n_reads, np = nr / 2;
        if (out->pair_mapq) pair_mapq.download(out->pair_mapq, np, st);
        if (out->read_mapq) read_mapq.download(out->read_mapq, nr, st);
        if (out->read_reverse) read_reverse.download(out->read_reverse, nr, st);
        if (out->chosen_slot) chosen_slot.download(out->chosen_slot, nr, st);
        if (out->pair_ll) pair_ll.download(out->pair_ll, np, st);
        if (have_columns) {
            size_t n = nr * maxcol;
            if (out->n_cols) o_n_cols.download(out->n_cols, nr, st);
            if (out->level) o_level.download(out->level, n, st);
            if (out->edge) o_edge.download(out->edge, n, st);
            if (out->gchar) o_gchar.download(out->gchar, n, st);
            if (out->schar) o_schar.download(out->schar, n, st);
            if (out->from_seed) o_fromseed.download(out->from_seed, n, st);
            if (out->mapq) o_mapq.download(out->mapq, n, st);
        }
        CUDA_OK(cudaMemcpyAsync(&n_errors, cs.error_count.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    std::string error_breakdown() {
        std::vector<int32_t> st((size_t)pb.n_chains); cs.status.download(st.data(), st.size(), 0); CUDA_OK(cudaStreamSynchronize(0));
        std::map<int, long long> cnt; for (int32_t v : st) if (v < 0) cnt[v]++;
        std::string r; for (auto& kv : cnt) r += " chains with status " + std::to_string(kv.first) + ": " + std::to_string(kv.second) + ";";
        return r;
    }
};

template <class F> int guarded(F&& f) {
    try { return f(); }
    catch (const CudaError& e) { return fail(HLALA_E_CUDA, e.what()); }
    catch (const std::exception& e) { return fail(HLALA_E_IO, e.what()); }
    catch (...) { return fail(HLALA_E_IO, "unknown exception"); }
}

} // namespace

extern "C" {

const char* hlala_last_error(void) { return g_err.c_str(); }

int hlala_graph_load(const char* dir, hlala_graph_t** out) {
    if (!dir || !out) return fail(HLALA_E_ARG, "hlala_graph_load: null argument");
    *out = nullptr;
    return guarded([&]() { std::unique_ptr<hlala_graph> g(new hlala_graph()); load_prg_dir(dir, g->h); *out = g.release(); return 0; });
}
void hlala_graph_free(hlala_graph_t* g) { delete g; }

int hlala_graph_to_gpu(hlala_graph_t* g, int device) {
    if (!g) return fail(HLALA_E_ARG, "hlala_graph_to_gpu: null graph");
    if (g->on_gpu && g->device == device) return 0;
    return guarded([&]() {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(HLALA_E_CUDA, "no CUDA device available: the alignment path has no CPU fallback");
        CUDA_OK(cudaSetDevice(device));
        cudaDeviceProp prop; CUDA_OK(cudaGetDeviceProperties(&prop, device));
        g->n_sm = prop.multiProcessorCount;
        const FlatGraph& h = g->h;
        if (h.max_nodes_per_level > K1_WCAP) return fail(HLALA_E_CAPACITY, "graph has a level with more than 256 nodes; kernels are built for <= 256");
        std::vector<uint32_t> pack((size_t)h.n_edges);
        for (int32_t e = 0; e < h.n_edges; e++) {
            int32_t f = h.edge_from[e], t = h.edge_to[e];
            pack[e] = (uint32_t)(f - h.level_node_off[h.node_level[f]]) | ((uint32_t)(t - h.level_node_off[h.node_level[t]]) << 8) | ((uint32_t)h.edge_emis[e] << 16);
        }
        pack.resize(pack.size() + 4, 0u);      // slack: the chain kernel's bulk copies read the 16-byte aligned superset of an edge range
        std::vector<int32_t> leo(h.level_edge_off); leo.resize((size_t)h.n_levels + 1, h.n_edges);
        g->bufs.clear(); g->kbufs.clear(); g->kix_on_gpu = false;
        DevGraph& d = g->d;
        d.n_levels = h.n_levels; d.n_nodes = h.n_nodes; d.n_edges = h.n_edges; d.n_contigs = h.n_contigs;
        d.level_node_off = g->up(h.level_node_off); d.level_edge_off = g->up(leo); d.edge_pack = g->up(pack); { std::vector<uint32_t> dpp = make_dp_pack(h); d.lvl4 = g->up(make_lvl4(h, dpp)); d.dp_pack = g->up(dpp); } d.edge_ord = g->up(h.edge_ord); d.edge_from = g->up(h.edge_from); d.edge_to = g->up(h.edge_to);
        d.node_out_off = g->up(h.node_out_off); d.node_out = g->up(h.node_out); d.node_in_off = g->up(h.node_in_off); d.node_in = g->up(h.node_in);
        d.path_off = g->up(h.path_off); d.path_edges = g->up(h.path_edges); d.path_from = g->up(h.path_from); d.path_to = g->up(h.path_to);
        d.jump_fwd_off = g->up(h.jump_fwd_off); d.jump_fwd_path = g->up(h.jump_fwd_path); d.jump_bwd_off = g->up(h.jump_bwd_off); d.jump_bwd_path = g->up(h.jump_bwd_path);
        {
            std::vector<I4> adj((size_t)h.n_nodes + 1), oa((size_t)h.n_edges), ia((size_t)h.n_edges), jf((size_t)h.n_paths), jb((size_t)h.n_paths);
            for (int32_t n = 0; n <= h.n_nodes; n++) adj[n] = I4{h.node_out_off[n], h.node_in_off[n], h.jump_fwd_off[n], h.jump_bwd_off[n]};
            for (int32_t k = 0; k < h.n_edges; k++) { int32_t e = h.node_out[k]; oa[k] = I4{e, h.edge_to[e], (int32_t)pack[e], 0}; int32_t e2 = h.node_in[k]; ia[k] = I4{e2, h.edge_from[e2], (int32_t)pack[e2], 0}; }
            for (int32_t k = 0; k < h.n_paths; k++) {
                int32_t p = h.jump_fwd_path[k], t = h.path_to[p]; jf[k] = I4{p, t, t - h.level_node_off[h.node_level[t]], h.path_off[p + 1] - h.path_off[p]};
                int32_t q = h.jump_bwd_path[k], f = h.path_from[q]; jb[k] = I4{q, f, f - h.level_node_off[h.node_level[f]], h.path_off[q + 1] - h.path_off[q]};
            }
            std::vector<uint8_t> gf((size_t)h.n_nodes, 0);
            for (int32_t e = 0; e < h.n_edges; e++) if (h.edge_emis[e] == '_') { gf[h.edge_from[e]] |= 1; gf[h.edge_to[e]] |= 2; }
            d.node_gapflags = g->up(gf);
            d.adj4 = g->up(adj); d.out_adj4 = g->up(oa); d.in_adj4 = g->up(ia); d.jf4 = g->up(jf); d.jb4 = g->up(jb);
        }
        d.gap_stretch = g->up(h.gap_stretch); d.contig_off = g->up(h.contig_off); d.contig_seq = g->up(h.contig_seq); d.contig_level = g->up(h.contig_level);
        d.contig_prg_id = g->up(h.contig_prg_id); d.anchor_off = g->up(h.anchor_off); d.anchor_prg_id = g->up(h.anchor_prg_id); d.anchor_pos = g->up(h.anchor_pos);
        ScoreTables t = make_score_tables(), tl = make_score_tables(true);
        CUDA_OK(upload_score_tables(t, tl));
        CUDA_OK(cudaDeviceSynchronize());
        g->on_gpu = true; g->device = device;
        return 0;
    });
}

int64_t hlala_graph_n_levels(const hlala_graph_t* g) { return g ? g->h.n_levels : -1; }
int64_t hlala_graph_n_nodes(const hlala_graph_t* g) { return g ? g->h.n_nodes : -1; }
int64_t hlala_graph_n_edges(const hlala_graph_t* g) { return g ? g->h.n_edges : -1; }
int64_t hlala_graph_n_paths(const hlala_graph_t* g) { return g ? g->h.n_paths : -1; }
int64_t hlala_graph_n_contigs(const hlala_graph_t* g) { return g ? g->h.n_contigs : -1; }
const char* hlala_graph_level_name(const hlala_graph_t* g, int64_t level) {
    if (!g || level < 0 || level >= (int64_t)g->h.level_names.size()) return nullptr;
    return g->h.level_names[(size_t)level].c_str();
}
int64_t hlala_graph_array(const hlala_graph_t* g, const char* name, const void** data) {
    if (!g || !name || !data) return -1;
    const FlatGraph& h = g->h; std::string n(name);
#define ARR(x) if (n == #x) { *data = h.x.data(); return (int64_t)h.x.size(); }
    ARR(level_node_off) ARR(node_ord) ARR(node_level) ARR(level_edge_off) ARR(edge_from) ARR(edge_to) ARR(edge_ord) ARR(edge_emis)
    ARR(path_off) ARR(path_edges) ARR(path_from) ARR(path_to) ARR(jump_fwd_off) ARR(jump_fwd_path) ARR(jump_bwd_off) ARR(jump_bwd_path)
    ARR(gap_stretch) ARR(contig_off) ARR(contig_seq) ARR(contig_level) ARR(contig_prg_id) ARR(anchor_off) ARR(anchor_prg_id) ARR(anchor_pos)
#undef ARR
    return -1;
}

int hlala_align_chains(hlala_graph_t* g, const hlala_seed_batch_t* batch, hlala_chain_out_t* out) {
    if (!g || !out) return fail(HLALA_E_ARG, "hlala_align_chains: null argument");
    if (int rc = check_batch(batch)) return rc;
    if (!g->on_gpu) return fail(HLALA_E_CUDA, "graph is not on a GPU: call hlala_graph_to_gpu first (there is no CPU fallback)");
    if (out->max_columns < 32 || out->max_columns > 2040) return fail(HLALA_E_ARG, "max_columns must be in [32, 2040]");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(g->device));
        cudaStream_t st = 0;
        Pipeline pl; pl.scratch_budget = (size_t)1 << 62; pl.allow_env_budget = false; pl.dedup = false; pl.prepare(g, *batch, out->max_columns, st);   // one wave: the chain records are exported whole
        pl.begin_run(st);
        {   // records of chains that are skipped (strand rule) never get a seed span: defined zeros in the exported arrays
            const size_t nb = (size_t)std::max(pl.pb.n_chains, 1) * 4; CUDA_OK(cudaMemsetAsync(pl.cs.seed_begin.p, 0, nb, st)); CUDA_OK(cudaMemsetAsync(pl.cs.seed_end.p, 0, nb, st));
        }
        Lane& L0 = *pl.lanes[0];
        pl.fork_lanes(st);
        if (pl.wave_pair.size() > 1) pl.run_chains_wave(0, L0);
        pl.join_lanes(st);
        const int32_t nc = pl.pb.n_chains, mc = out->max_columns; ChainScratch& cs = pl.cs;
        DevBuf o_level, o_edge, o_g; size_t ncol = (size_t)std::max(nc, 1) * mc;
        o_level.alloc(ncol * 4); o_edge.alloc(ncol * 4); o_g.alloc(ncol);
        CUDA_OK(launch_export_chain_columns(g->d, nc, mc, cs.n_cols.as<int32_t>(), cs.first_level.as<int32_t>(), L0.c_edge.as<int32_t>(), o_level.as<int32_t>(), o_edge.as<int32_t>(), o_g.as<uint8_t>(), st));
        if (out->chain_order) memcpy(out->chain_order, pl.pb.chain_order.data(), (size_t)nc * 4);
        if (out->status) cs.status.download(out->status, nc, st);
        if (out->n_cols) cs.n_cols.download(out->n_cols, nc, st);
        if (out->seed_begin) cs.seed_begin.download(out->seed_begin, nc, st);
        if (out->seed_end) cs.seed_end.download(out->seed_end, nc, st);
        if (out->ll) cs.ll.download(out->ll, nc, st);
        if (out->level) o_level.download(out->level, (size_t)nc * mc, st);
        if (out->edge) o_edge.download(out->edge, (size_t)nc * mc, st);
        if (out->gchar) o_g.download(out->gchar, (size_t)nc * mc, st);
        if (out->schar) L0.c_schar.download(out->schar, (size_t)nc * mc, st);
        if (out->from_seed) L0.c_fromseed.download(out->from_seed, (size_t)nc * mc, st);
        CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    });
}

int hlala_align_pairs(hlala_graph_t* g, const hlala_seed_batch_t* batch, double is_mean, double is_sd, hlala_pair_out_t* out, int32_t* bases_per_level) {
    if (!g || !out) return fail(HLALA_E_ARG, "hlala_align_pairs: null argument");
    if (int rc = check_batch(batch)) return rc;
    if (!g->on_gpu) return fail(HLALA_E_CUDA, "graph is not on a GPU: call hlala_graph_to_gpu first (there is no CPU fallback)");
    if (out->max_columns < 32 || out->max_columns > 2040) return fail(HLALA_E_ARG, "max_columns must be in [32, 2040]");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(g->device));
        cudaStream_t st = 0;
        std::lock_guard<std::mutex> lock(g->ws_mu);
        if (!g->ws) g->ws = std::shared_ptr<void>(std::make_shared<Pipeline>());
        Pipeline& pl = *static_cast<Pipeline*>(g->ws.get());
        const bool trace = getenv("HLALA_TRACE_HOST") != nullptr; auto t0 = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) { if (!trace) return; CUDA_OK(cudaDeviceSynchronize()); auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[host-trace] %-10s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count()); t0 = t1; };
        pl.reset_config();
        pl.prepare(g, *batch, out->max_columns, st);
        lap("prepare");
        DevBuf& bpl = pl.bpl_ws; size_t nl = (size_t)std::max(g->h.n_levels - 1, 1);
        if (bases_per_level) { bpl.alloc(nl * 4); CUDA_OK(cudaMemsetAsync(bpl.p, 0, nl * 4, st)); }
        const bool want_cols = out->level || out->edge || out->gchar || out->schar || out->from_seed || out->mapq || out->n_cols;
        pl.run(is_mean, is_sd, bases_per_level ? bpl.as<int32_t>() : nullptr, want_cols, st);
        lap("run");
        pl.fetch(out, st);
        if (bases_per_level) {
            std::vector<int32_t>& h = pl.bpl_host; h.resize(nl); bpl.download(h.data(), nl, st); CUDA_OK(cudaStreamSynchronize(st));
            for (size_t i = 0; i + 1 < (size_t)g->h.n_levels; i++) bases_per_level[i] += h[i];
        }
        lap("fetch");
        if (pl.n_errors > 0) return fail(HLALA_E_INVARIANT, std::to_string(pl.n_errors) + " chains/pairs violated a reference invariant (-5) or a kernel capacity (-4):" + pl.error_breakdown());
        return 0;
    });
}

// processBAM::alignReadsUnpaired_postSeedExtraction_andStoreInto / alignOneLongRead (processBAM.cpp:2224-2340, 3618-3838): every read on its own.
int hlala_align_long_reads(hlala_graph_t* g, const hlala_seed_batch_t* batch, hlala_pair_out_t* out, int32_t* bases_per_level) {
    if (!g || !out || !batch) return fail(HLALA_E_ARG, "hlala_align_long_reads: null argument");
    if (batch->n_reads < 0 || !batch->read_off || !batch->chain_off || !batch->cigar_off) return fail(HLALA_E_ARG, "seed batch: null offset arrays");
    if (out->max_columns < 32 || out->max_columns > 16384) return fail(HLALA_E_ARG, "max_columns must be in [32, 16384] in long-read mode");
    if (!g->on_gpu) return fail(HLALA_E_CUDA, "graph is not on a GPU: call hlala_graph_to_gpu first (there is no CPU fallback)");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(g->device)); cudaStream_t st = 0;
        // internally a read is a pair whose second mate is empty (no bases, no chains): the wave / slot / output indexing of the paired path is kept
        const int64_t n = batch->n_reads; std::vector<int64_t> ro((size_t)(2 * n + 1)); std::vector<int32_t> co((size_t)(2 * n + 1));
        for (int64_t r = 0; r < n; r++) { ro[(size_t)(2 * r)] = batch->read_off[r]; ro[(size_t)(2 * r + 1)] = batch->read_off[r + 1]; co[(size_t)(2 * r)] = batch->chain_off[r]; co[(size_t)(2 * r + 1)] = batch->chain_off[r + 1]; }
        ro[(size_t)(2 * n)] = batch->read_off[n]; co[(size_t)(2 * n)] = batch->chain_off[n];
        hlala_seed_batch_t b2 = *batch; b2.n_reads = 2 * n; b2.read_off = ro.data(); b2.chain_off = co.data();
        // the workspace kept in the graph handle across calls (device buffers are grow-only), as hlala_align_pairs does
        std::lock_guard<std::mutex> lock(g->ws_mu);
        if (!g->ws) g->ws = std::shared_ptr<void>(std::make_shared<Pipeline>());
        Pipeline& pl = *static_cast<Pipeline*>(g->ws.get());
        pl.reset_config(); pl.long_mode = true; pl.prepare(g, b2, out->max_columns, st);
        DevBuf& bpl = pl.bpl_ws; size_t nl = (size_t)std::max(g->h.n_levels - 1, 1);
        if (bases_per_level) { bpl.alloc(nl * 4); CUDA_OK(cudaMemsetAsync(bpl.p, 0, nl * 4, st)); }
        const bool want_cols = out->level || out->edge || out->gchar || out->schar || out->from_seed || out->mapq || out->n_cols;
        pl.run(100.0, 10.0, bases_per_level ? bpl.as<int32_t>() : nullptr, want_cols, st);       // the insert-size model is not used for single reads
        // fetch in the internal layout, hand the even rows (the reads) to the caller
        const size_t mc = (size_t)out->max_columns; const size_t nr2 = (size_t)(2 * n);
        std::vector<double> pm((size_t)n), rm(nr2), pll((size_t)n); std::vector<uint8_t> rr(nr2); std::vector<int32_t> cslot(nr2), ncol(nr2);
        hlala_pair_out_t tmp{}; tmp.max_columns = out->max_columns; tmp.pair_mapq = pm.data(); tmp.read_mapq = rm.data(); tmp.read_reverse = rr.data(); tmp.chosen_slot = cslot.data(); tmp.pair_ll = pll.data(); tmp.n_cols = ncol.data();
        std::vector<int32_t> lv, ed; std::vector<uint8_t> gc, sc, fs, mq;
        if (want_cols) { if (out->level) { lv.resize(nr2 * mc); tmp.level = lv.data(); } if (out->edge) { ed.resize(nr2 * mc); tmp.edge = ed.data(); } if (out->gchar) { gc.resize(nr2 * mc); tmp.gchar = gc.data(); }
            if (out->schar) { sc.resize(nr2 * mc); tmp.schar = sc.data(); } if (out->from_seed) { fs.resize(nr2 * mc); tmp.from_seed = fs.data(); } if (out->mapq) { mq.resize(nr2 * mc); tmp.mapq = mq.data(); } }
        pl.fetch(&tmp, st);
        for (int64_t r = 0; r < n; r++) {
            if (out->pair_mapq) out->pair_mapq[r] = pm[(size_t)r]; if (out->read_mapq) out->read_mapq[r] = rm[(size_t)(2 * r)]; if (out->read_reverse) out->read_reverse[r] = rr[(size_t)(2 * r)];
            if (out->chosen_slot) out->chosen_slot[r] = cslot[(size_t)(2 * r)]; if (out->pair_ll) out->pair_ll[r] = pll[(size_t)r]; if (out->n_cols) out->n_cols[r] = ncol[(size_t)(2 * r)];
            const size_t src = (size_t)(2 * r) * mc, dst = (size_t)r * mc;
            if (out->level) memcpy(out->level + dst, lv.data() + src, mc * 4); if (out->edge) memcpy(out->edge + dst, ed.data() + src, mc * 4);
            if (out->gchar) memcpy(out->gchar + dst, gc.data() + src, mc); if (out->schar) memcpy(out->schar + dst, sc.data() + src, mc);
            if (out->from_seed) memcpy(out->from_seed + dst, fs.data() + src, mc); if (out->mapq) memcpy(out->mapq + dst, mq.data() + src, mc);
        }
        if (bases_per_level) { std::vector<int32_t> h(nl); bpl.download(h.data(), nl, st); CUDA_OK(cudaStreamSynchronize(st)); for (size_t i = 0; i + 1 < (size_t)g->h.n_levels; i++) bases_per_level[i] += h[i]; }
        if (pl.n_errors > 0) return fail(HLALA_E_INVARIANT, std::to_string(pl.n_errors) + " chains/reads violated a reference invariant (-5) or a kernel capacity (-4):" + pl.error_breakdown());
        return 0;
    });
}

struct hlala_session { Pipeline pl; std::vector<uint8_t> typing_blob; bool own_cov = false; DevBuf cov; };

int hlala_session_create(hlala_graph_t* g, const hlala_seed_batch_t* batch, int32_t max_columns, hlala_session_t** out) {
    if (!g || !out) return fail(HLALA_E_ARG, "hlala_session_create: null argument");
    *out = nullptr;
    if (int rc = check_batch(batch)) return rc;
    if (!g->on_gpu) return fail(HLALA_E_CUDA, "graph is not on a GPU: call hlala_graph_to_gpu first (there is no CPU fallback)");
    if (max_columns < 32 || max_columns > 2040) return fail(HLALA_E_ARG, "max_columns must be in [32, 2040]");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(g->device));
        std::unique_ptr<hlala_session> s(new hlala_session());
        s->pl.prepare(g, *batch, max_columns, 0);
        s->pl.compute_algorithmic_bytes(*batch);
        CUDA_OK(cudaStreamSynchronize(0));
        *out = s.release();
        return 0;
    });
}
void hlala_session_free(hlala_session_t* s) { delete s; }

int hlala_session_run(hlala_session_t* s, double is_mean, double is_sd, uint64_t bases_per_level_dev, void* cuda_stream) {
    if (!s) return fail(HLALA_E_ARG, "hlala_session_run: null session");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(s->pl.g->device));
        cudaStream_t st = (cudaStream_t)cuda_stream;
        int32_t* cov = (int32_t*)(uintptr_t)bases_per_level_dev;
        if (!cov && s->own_cov) { const size_t nl = (size_t)std::max(s->pl.g->h.n_levels - 1, 1); s->cov.alloc(nl * 4); CUDA_OK(cudaMemsetAsync(s->cov.p, 0, nl * 4, st)); cov = s->cov.as<int32_t>(); }
        s->pl.run(is_mean, is_sd, cov, s->pl.keep_columns, st);
        return 0;
    });
}
int hlala_session_launches(const hlala_session_t* s) { return s ? s->pl.launches : -1; }
int hlala_session_set_timing(hlala_session_t* s, int on) {
    if (!s) return fail(HLALA_E_ARG, "null session");
    return guarded([&]() { s->pl.timing = on != 0; if (on) { CUDA_OK(cudaSetDevice(s->pl.g->device)); s->pl.dp_bytes.alloc(8); CUDA_OK(cudaMemset(s->pl.dp_bytes.p, 0, 8)); } return 0; });
}
int64_t hlala_session_dp_kernel_bytes(hlala_session_t* s) {
    if (!s || !s->pl.dp_bytes.p) return -1;
    unsigned long long v = 0; if (cudaSetDevice(s->pl.g->device) != cudaSuccess || cudaMemcpy(&v, s->pl.dp_bytes.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}
int hlala_session_timing(hlala_session_t* s, double ms[6], int launches[6]) {
    if (!s) return fail(HLALA_E_ARG, "null session");
    return guarded([&]() { s->pl.collect_timing(ms, launches); return 0; });
}
int64_t hlala_session_chain_kernel_bytes(const hlala_session_t* s) { return s ? s->pl.chain_kernel_bytes : -1; }
int64_t hlala_session_algorithmic_bytes(const hlala_session_t* s) { return s ? s->pl.algo_bytes : -1; }
int hlala_session_aligned_bytes(hlala_session_t* s, int64_t out[3]) {
    if (!s || !out) return fail(HLALA_E_ARG, "hlala_session_aligned_bytes: null argument");
    return guarded([&]() { CUDA_OK(cudaSetDevice(s->pl.g->device)); if (s->pl.cb_cols.empty()) return fail(HLALA_E_ARG, "hlala_session_aligned_bytes: session created without byte accounting"); s->pl.aligned_bytes(out); return 0; });
}
int hlala_session_fetch(hlala_session_t* s, hlala_pair_out_t* out) {
    if (!s || !out) return fail(HLALA_E_ARG, "hlala_session_fetch: null argument");
    return guarded([&]() { CUDA_OK(cudaSetDevice(s->pl.g->device)); s->pl.fetch(out, 0); return 0; });
}
int hlala_session_digest(hlala_session_t* s, int64_t out[4], double* sum_pair_ll) {
    if (!s || !out) return fail(HLALA_E_ARG, "hlala_session_digest: null argument");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(s->pl.g->device));
        unsigned long long d[4]; CUDA_OK(cudaMemcpy(d, s->pl.digest.p, 32, cudaMemcpyDeviceToHost));
        int32_t ne = 0; CUDA_OK(cudaMemcpy(&ne, s->pl.cs.error_count.p, 4, cudaMemcpyDeviceToHost));
        out[0] = (int64_t)d[0]; out[1] = (int64_t)d[1]; out[2] = (int64_t)d[2]; out[3] = ne;
        if (sum_pair_ll) {
            size_t np = (size_t)s->pl.pb.n_reads / 2; std::vector<double> v(np); s->pl.pair_ll.download(v.data(), np, 0); CUDA_OK(cudaStreamSynchronize(0));
            double acc = 0; for (double x : v) acc += x; *sum_pair_ll = acc;
        }
        return 0;
    });
}

} // extern "C"

// ===================================================================================================================
// HLA typing stage
namespace {

TypingScoreTables make_typing_tables(bool long_reads = false) {   // HLATyper.cpp:935-958, 2189-2216; Utilities.cpp:357-377, 1368-1379
    TypingScoreTables t;
    const double insertionP = long_reads ? 0.075 : 0.001, deletionP = long_reads ? 0.075 : 0.001;
    const double ll_ins = log(insertionP);
    t.ll_ins_actual = ll_ins + log(1.0 / 4.0); t.ll_del = log(deletionP); t.ll_mm = log(1 - insertionP - deletionP);
    t.log_half = log(0.5); t.log_two = log(1 + exp(0.0));
    for (int q = 0; q < 256; q++) {
        if (q < 33) { t.log_pc[q] = NAN; t.log_pi[q] = NAN; continue; }
        double pc = 1 - exp(log(10) * ((double)(q - 33) / (double)-10));
        if (pc > 0.999) pc = 0.999;
        if (pc == 0) pc = 0.001;
        t.log_pc[q] = log(pc);
        double pi = (1 - pc) * (1.0 / 3.0);
        t.log_pi[q] = log(pi);
    }
    return t;
}

struct GpuTypingDevice : TypingDevice {
    int device = 0, rank = 0, world = 1; hlala_allreduce_f64_fn allreduce = nullptr; void* ctx = nullptr; cudaStream_t st = 0;
    double ms[2] = {0, 0}; int launches[2] = {0, 0}; double work[2] = {0, 0};
    void abort_locus(int32_t C) override {     // see TypingDevice::abort_locus: NaNs into this locus' all-reduce
        if (world <= 1 || !allreduce) return;
        CUDA_OK(cudaSetDevice(device));
        const size_t npair = (size_t)C * ((size_t)C + 1) / 2; DevBuf d; d.alloc(3 * npair * 8); CUDA_OK(cudaMemsetAsync(d.p, 0xFF, 3 * npair * 8, st));
        allreduce(ctx, (uint64_t)(uintptr_t)d.p, (int64_t)(3 * npair), (void*)st); CUDA_OK(cudaStreamSynchronize(st));
    }
    void run_locus(const LocusDeviceInput& in, bool want_read_ll, LocusDeviceOutput& out) override {
        CUDA_OK(cudaSetDevice(device));   // entered from the host thread pool of run_typing (one locus at a time, in locus order)
        static const bool prof = getenv("HLALA_TYPING_PROFILE") != nullptr; auto tp0 = std::chrono::steady_clock::now();
        auto lapp = [&](const char* what) { if (!prof) return; auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[typing-device] C=%d R=%d %-12s %8.3f ms\n", in.C, in.R, what, std::chrono::duration<double, std::milli>(t1 - tp0).count()); tp0 = t1; };
        const int32_t C = in.C, P = in.P, R = in.R; const int32_t Cpad = (C + 31) / 32 * 32;
        const size_t npair = (size_t)C * ((size_t)C + 1) / 2;
        std::vector<uint8_t> ct((size_t)P * Cpad, (uint8_t)'_');
        for (int32_t c = 0; c < C; c++) { const std::string& s = (*in.cluster_seq)[(size_t)c]; for (int32_t p = 0; p < P; p++) ct[(size_t)p * Cpad + c] = (uint8_t)s[(size_t)p]; }
        DevBuf d_ct, d_off, d_pos, d_c0, d_q0, d_glen, d_ll, d_mm, d_pair;
        d_ct.upload(ct, st); d_off.upload(in.rec_off, st);
        if (!in.rec_pos.empty()) { d_pos.upload(in.rec_pos, st); d_c0.upload(in.rec_c0, st); d_q0.upload(in.rec_q0, st); d_glen.upload(in.rec_glen, st); }
        d_ll.alloc((size_t)std::max(R, 1) * Cpad * 8); d_mm.alloc((size_t)std::max(R, 1) * Cpad * 4); d_pair.alloc(3 * npair * 8);
        const int32_t r0 = (int32_t)((int64_t)R * rank / world), r1 = (int32_t)((int64_t)R * (rank + 1) / world);
        if (want_read_ll) { CUDA_OK(cudaMemsetAsync(d_ll.p, 0, d_ll.bytes, st)); CUDA_OK(cudaMemsetAsync(d_mm.p, 0, d_mm.bytes, st)); }
        int32_t max_rec = 0; double steps = 0;
        for (int32_t r = 0; r < R; r++) { max_rec = std::max(max_rec, in.rec_off[r + 1] - in.rec_off[r]); if (r >= r0 && r < r1) steps += (double)(in.rec_off[r + 1] - in.rec_off[r]); }
        lapp("alloc+upload");
        cudaEvent_t e0, e1, e2; CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1)); CUDA_OK(cudaEventCreate(&e2));
        CUDA_OK(cudaEventRecord(e0, st));
        CUDA_OK(launch_read_cluster_ll(d_ct.as<uint8_t>(), C, Cpad, r0, r1, max_rec, d_off.as<int32_t>(), d_pos.as<int16_t>(), d_c0.as<uint8_t>(), d_q0.as<uint8_t>(), d_glen.as<uint16_t>(), d_ll.as<double>(), d_mm.as<int32_t>(), st));
        CUDA_OK(cudaEventRecord(e1, st));
        double* pl = d_pair.as<double>();
        if (getenv("HLALA_TYPING_TERMWISE")) CUDA_OK(launch_allele_pair_ll(d_ll.as<double>(), d_mm.as<int32_t>(), C, Cpad, r0, r1, pl, pl + npair, pl + 2 * npair, st));     // A/B hook: one exp + one log per (pair, read)
        else {
            DevBuf d_e, d_rm, d_cs, d_ms; d_e.alloc((size_t)std::max(R, 1) * Cpad * 8); d_rm.alloc((size_t)std::max(R, 1) * 8); d_cs.alloc((size_t)std::max(C, 1) * 8); d_ms.alloc(8);
            CUDA_OK(launch_allele_pair_prod(d_ll.as<double>(), d_mm.as<int32_t>(), C, Cpad, r0, r1, d_e.as<double>(), d_rm.as<double>(), d_cs.as<double>(), d_ms.as<double>(), pl, pl + npair, pl + 2 * npair, st));
            CUDA_OK(cudaStreamSynchronize(st));     // the scratch buffers above are freed at the end of this scope
        }
        CUDA_OK(cudaEventRecord(e2, st));
        if (prof) { CUDA_OK(cudaStreamSynchronize(st)); lapp("kernels"); }
        if (world > 1) {
            if (!allreduce) throw std::runtime_error("hlala_typer_infer: world > 1 needs an all-reduce callback");
            if (allreduce(ctx, (uint64_t)(uintptr_t)d_pair.p, (int64_t)(3 * npair), (void*)st) != 0) throw std::runtime_error("hlala_typer_infer: all-reduce callback failed");
        }
        lapp("all-reduce");
        if (world > 1) { double first = 0; CUDA_OK(cudaMemcpyAsync(&first, pl, 8, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); if (first != first) throw std::runtime_error("hlala_typer_infer: this locus failed on another rank (poisoned all-reduce)"); }
        out.pair_ll.resize(npair); out.pair_mavg.resize(npair); out.pair_mmin.resize(npair);
        CUDA_OK(cudaMemcpyAsync(out.pair_ll.data(), pl, npair * 8, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(out.pair_mavg.data(), pl + npair, npair * 8, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(out.pair_mmin.data(), pl + 2 * npair, npair * 8, cudaMemcpyDeviceToHost, st));
        std::vector<double> llt; std::vector<int32_t> mmt;
        if (want_read_ll && R > 0) { llt.resize((size_t)R * Cpad); mmt.resize((size_t)R * Cpad); d_ll.download(llt.data(), llt.size(), st); d_mm.download(mmt.data(), mmt.size(), st); }
        CUDA_OK(cudaStreamSynchronize(st));
        lapp("download");
        float f = 0; CUDA_OK(cudaEventElapsedTime(&f, e0, e1)); ms[0] += f; CUDA_OK(cudaEventElapsedTime(&f, e1, e2)); ms[1] += f; launches[0] += (r1 > r0); launches[1]++;
        work[0] += steps * C; work[1] += (double)npair * (r1 - r0);
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        out.LL.clear(); out.mism.clear();
        if (want_read_ll) { out.LL.assign((size_t)C * R, 0.0); out.mism.assign((size_t)C * R, 0);
            for (int32_t r = 0; r < R; r++) for (int32_t c = 0; c < C; c++) { out.LL[(size_t)c * R + r] = llt[(size_t)r * Cpad + c]; out.mism[(size_t)c * R + r] = mmt[(size_t)r * Cpad + c]; } }
    }
};

} // namespace

struct hlala_typer { TypingTables T; std::vector<LocusCall> calls; double ms[2] = {0, 0}; int launches[2] = {0, 0}; double work[2] = {0, 0}; bool tables_on_device = false, tables_long = false; int device = -1;
                     std::vector<uint8_t> long_blob; };

extern "C" {

int hlala_typer_create(const char* dir, hlala_typer_t** out) {
    if (!dir || !out) return fail(HLALA_E_ARG, "hlala_typer_create: null argument");
    *out = nullptr;
    return guarded([&]() { std::unique_ptr<hlala_typer> t(new hlala_typer()); t->T.load(dir); *out = t.release(); return 0; });
}
void hlala_typer_free(hlala_typer_t* t) { delete t; }
int hlala_typer_n_loci(const hlala_typer_t* t) { return t ? (int)t->T.loci.size() : -1; }
const char* hlala_typer_locus_name(const hlala_typer_t* t, int l) { return (t && l >= 0 && l < (int)t->T.loci.size()) ? t->T.loci[(size_t)l].name.c_str() : nullptr; }
int hlala_typer_locus_dims(const hlala_typer_t* t, int l, int32_t* C, int32_t* P) {
    if (!t || l < 0 || l >= (int)t->T.loci.size()) return fail(HLALA_E_ARG, "hlala_typer_locus_dims: bad locus");
    if (C) *C = t->T.loci[(size_t)l].C(); if (P) *P = t->T.loci[(size_t)l].P(); return 0;
}

int hlala_session_set_keep_columns(hlala_session_t* s, int on) { if (!s) return fail(HLALA_E_ARG, "null session"); s->pl.keep_columns = on != 0; return 0; }

int hlala_session_typing_extract(hlala_session_t* s, const hlala_typer_t* t, const char* const* pair_names, int64_t pair_index_base,
                                 const uint8_t** blob, int64_t* blob_bytes, int64_t* n_selected) {
    if (!s || !t || !blob || !blob_bytes) return fail(HLALA_E_ARG, "hlala_session_typing_extract: null argument");
    Pipeline& pl = s->pl;
    if (!pl.have_columns) return fail(HLALA_E_ARG, "hlala_session_typing_extract: run the session with hlala_session_set_keep_columns(s, 1) first");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(pl.g->device)); cudaStream_t st = 0;
        const int64_t nr = pl.pb.n_reads, np = nr / 2;
        if (t->T.gene_bounds.size() > (size_t)TY_MAX_GENES) return fail(HLALA_E_CAPACITY, "more than 64 genes in segments.txt");
        GeneBounds gb; gb.n = (int32_t)t->T.gene_bounds.size(); for (int i = 0; i < gb.n; i++) { gb.first[i] = t->T.gene_bounds[(size_t)i].first; gb.last[i] = t->T.gene_bounds[(size_t)i].second; }
        DevBuf d_flag; d_flag.alloc((size_t)std::max<int64_t>(np, 1));
        CUDA_OK(launch_gene_filter(gb, np, pl.maxcol, pl.o_n_cols.as<int32_t>(), pl.o_level.as<int32_t>(), d_flag.as<uint8_t>(), st));
        std::vector<uint8_t> flag((size_t)np), rev((size_t)nr); std::vector<int32_t> ncols((size_t)nr); std::vector<double> rmq((size_t)nr);
        d_flag.download(flag.data(), (size_t)np, st); pl.o_n_cols.download(ncols.data(), (size_t)nr, st); pl.read_reverse.download(rev.data(), (size_t)nr, st); pl.read_mapq.download(rmq.data(), (size_t)nr, st);
        CUDA_OK(cudaStreamSynchronize(st));
        TypingReads tr; tr.col_off.push_back(0); tr.base_off.push_back(0); std::vector<int64_t> src;
        for (int64_t p = 0; p < np; p++) if (flag[(size_t)p]) {
            tr.pair_id.push_back(pair_index_base + p); tr.name.push_back(pair_names ? std::string(pair_names[p]) : "r" + std::to_string(pair_index_base + p));
            for (int m = 0; m < 2; m++) { const int64_t r = 2 * p + m; src.push_back(r); tr.col_off.push_back(tr.col_off.back() + ncols[(size_t)r]); tr.base_off.push_back(tr.base_off.back() + (pl.host_read_off[(size_t)r + 1] - pl.host_read_off[(size_t)r]));
                tr.reverse.push_back(rev[(size_t)r]); tr.mapq.push_back(rmq[(size_t)r]); }
        }
        const size_t ncol = (size_t)tr.col_off.back(), nb = (size_t)tr.base_off.back();
        tr.level.resize(ncol); tr.g.resize(ncol); tr.s.resize(ncol); tr.mq.resize(ncol); tr.bases.resize(nb); tr.quals.resize(nb);
        if (!src.empty()) {
            DevBuf d_src, d_co, d_bo, o_l, o_g, o_s, o_m, o_b, o_q; d_src.upload(src, st); d_co.upload(tr.col_off, st); d_bo.upload(tr.base_off, st);
            o_l.alloc(std::max<size_t>(ncol, 1) * 4); o_g.alloc(std::max<size_t>(ncol, 1)); o_s.alloc(std::max<size_t>(ncol, 1)); o_m.alloc(std::max<size_t>(ncol, 1)); o_b.alloc(std::max<size_t>(nb, 1)); o_q.alloc(std::max<size_t>(nb, 1));
            CUDA_OK(launch_gather_reads((int64_t)src.size(), d_src.as<int64_t>(), d_co.as<int64_t>(), d_bo.as<int64_t>(), pl.maxcol, pl.o_n_cols.as<int32_t>(), pl.o_level.as<int32_t>(), pl.o_gchar.as<uint8_t>(), pl.o_schar.as<uint8_t>(),
                                        pl.o_mapq.as<uint8_t>(), pl.db.view.read_off, pl.db.view.bases, pl.db.view.quals, o_l.as<int32_t>(), o_g.as<uint8_t>(), o_s.as<uint8_t>(), o_m.as<uint8_t>(), o_b.as<uint8_t>(), o_q.as<uint8_t>(), st));
            o_l.download(tr.level.data(), ncol, st); o_g.download(tr.g.data(), ncol, st); o_s.download(tr.s.data(), ncol, st); o_m.download(tr.mq.data(), ncol, st); o_b.download(tr.bases.data(), nb, st); o_q.download(tr.quals.data(), nb, st);
            CUDA_OK(cudaStreamSynchronize(st));
        }
        s->typing_blob = tr.serialize();
        *blob = s->typing_blob.data(); *blob_bytes = (int64_t)s->typing_blob.size(); if (n_selected) *n_selected = (int64_t)tr.n_pairs();
        return 0;
    });
}

// Long-read mode: gene filter of alignReadsUnpaired_postSeedExtraction_andStoreInto (processBAM.cpp:2297-2331) over the host arrays hlala_align_long_reads returned.
int hlala_typing_blob_from_long_reads(hlala_typer_t* t, const hlala_seed_batch_t* batch, const char* const* read_names, const hlala_pair_out_t* aligned,
                                      const uint8_t** blob, int64_t* blob_bytes, int64_t* n_selected) {
    if (!t || !batch || !aligned || !blob || !blob_bytes) return fail(HLALA_E_ARG, "hlala_typing_blob_from_long_reads: null argument");
    if (!aligned->n_cols || !aligned->level || !aligned->gchar || !aligned->schar || !aligned->mapq || !aligned->read_reverse || !aligned->pair_mapq)
        return fail(HLALA_E_ARG, "hlala_typing_blob_from_long_reads: the typing stage needs n_cols, level, gchar, schar, mapq, read_reverse and pair_mapq of hlala_align_long_reads");
    return guarded([&]() {
        TypingReads tr = long_read_typing_input(t->T, batch->n_reads, read_names, batch->read_off, batch->bases, batch->quals, aligned->max_columns, aligned->n_cols, aligned->level,
                                                aligned->gchar, aligned->schar, aligned->mapq, aligned->read_reverse, aligned->pair_mapq);
        t->long_blob = tr.serialize();
        *blob = t->long_blob.data(); *blob_bytes = (int64_t)t->long_blob.size(); if (n_selected) *n_selected = (int64_t)tr.n_pairs();
        return 0;
    });
}

int hlala_typer_infer(hlala_typer_t* t, int device, const uint8_t* const* blobs, const int64_t* blob_bytes, int n_blobs, double is_mean, double is_sd,
                      const char* out_dir, const char* g_nom_dir, int rank, int world, hlala_allreduce_f64_fn allreduce, void* allreduce_ctx, int keep_read_ll) {
    if (!t || !blobs || !blob_bytes || n_blobs < 1 || !g_nom_dir) return fail(HLALA_E_ARG, "hlala_typer_infer: null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(HLALA_E_ARG, "hlala_typer_infer: bad rank/world");
    try {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(HLALA_E_CUDA, "no CUDA device available: the typing kernels have no CPU fallback");
        CUDA_OK(cudaSetDevice(device));
        TypingReads all; for (int i = 0; i < n_blobs; i++) all.deserialize_append(blobs[i], (size_t)blob_bytes[i]);
        if (!t->tables_on_device || t->device != device || t->tables_long != all.long_reads) { CUDA_OK(upload_typing_tables(make_typing_tables(all.long_reads))); t->tables_on_device = true; t->device = device; t->tables_long = all.long_reads; }
        GpuTypingDevice dev; dev.device = device; dev.rank = rank; dev.world = world; dev.allreduce = allreduce; dev.ctx = allreduce_ctx;
        TypingOptions opt; opt.keep_read_ll = (keep_read_ll & HLALA_TYPER_KEEP_READ_LL) != 0;
        if (allreduce && !(keep_read_ll & HLALA_TYPER_CALLBACK_ANY_THREAD)) opt.threads = 1;   // the caller's all-reduce callback (NCCL, Python) is then only ever entered from the calling thread
        t->calls.clear();
        run_typing(t->T, all, is_mean, is_sd, out_dir ? std::string(out_dir) : std::string(), g_nom_dir, dev, opt, t->calls);
        for (int k = 0; k < 2; k++) { t->ms[k] = dev.ms[k]; t->launches[k] = dev.launches[k]; t->work[k] = dev.work[k]; }
        return 0;
    }
    catch (const CudaError& e) { return fail(HLALA_E_CUDA, e.what()); }
    catch (const std::runtime_error& e) { const std::string m = e.what(); return fail(m.find("reference assertion would fail") != std::string::npos ? HLALA_E_INVARIANT : HLALA_E_IO, m); }
    catch (const std::exception& e) { return fail(HLALA_E_IO, e.what()); }
}

int hlala_typer_result_dims(const hlala_typer_t* t, int l, int32_t* C, int32_t* R) {
    if (!t || l < 0 || l >= (int)t->calls.size()) return fail(HLALA_E_ARG, "hlala_typer_result_dims: no result for this locus");
    if (C) *C = t->calls[(size_t)l].C; if (R) *R = t->calls[(size_t)l].R; return 0;
}
int hlala_typer_result_read_ll(const hlala_typer_t* t, int l, double* ll, int32_t* mm) {
    if (!t || l < 0 || l >= (int)t->calls.size()) return fail(HLALA_E_ARG, "hlala_typer_result_read_ll: no result for this locus");
    const LocusCall& c = t->calls[(size_t)l];
    if (c.dev.LL.size() != (size_t)c.C * c.R) return fail(HLALA_E_ARG, "hlala_typer_result_read_ll: run hlala_typer_infer with keep_read_ll");
    if (ll) memcpy(ll, c.dev.LL.data(), c.dev.LL.size() * 8); if (mm) memcpy(mm, c.dev.mism.data(), c.dev.mism.size() * 4); return 0;
}
int hlala_typer_result_pair_ll(const hlala_typer_t* t, int l, double* pl, double* ma, double* mn) {
    if (!t || l < 0 || l >= (int)t->calls.size()) return fail(HLALA_E_ARG, "hlala_typer_result_pair_ll: no result for this locus");
    const LocusCall& c = t->calls[(size_t)l];
    if (pl) memcpy(pl, c.dev.pair_ll.data(), c.dev.pair_ll.size() * 8); if (ma) memcpy(ma, c.dev.pair_mavg.data(), c.dev.pair_mavg.size() * 8); if (mn) memcpy(mn, c.dev.pair_mmin.data(), c.dev.pair_mmin.size() * 8); return 0;
}
int hlala_typer_result_call(const hlala_typer_t* t, int l, const char** a1, const char** a2, double* q1, double* q2) {
    if (!t || l < 0 || l >= (int)t->calls.size()) return fail(HLALA_E_ARG, "hlala_typer_result_call: no result for this locus");
    const LocusCall& c = t->calls[(size_t)l];
    if (a1) *a1 = c.call1.c_str(); if (a2) *a2 = c.call2.c_str(); if (q1) *q1 = c.q1; if (q2) *q2 = c.q2; return 0;
}
int hlala_evaluate_types(const char* sample_id, const char* bestguess_file, const char* true_types_file, char* loci_out, int64_t loci_cap,
                         int32_t* counts_out, int32_t max_loci, char* summary_out, int64_t summary_cap) {
    if (!sample_id || !bestguess_file || !true_types_file) return fail(HLALA_E_ARG, "hlala_evaluate_types: null argument");
    int n = 0;
    int rc = guarded([&]() {
        InferredTypes inf; TrueTypes tru; read_inferred_types(sample_id, inf, bestguess_file); read_true_types(tru, true_types_file);
        std::string text; const std::map<std::string, std::pair<int, int>> r = evaluate_types(tru, inf, &text);
        std::string loci; int i = 0;
        for (const auto& kv : r) { if (i) loci += ';'; loci += kv.first; if (counts_out && i < max_loci) { counts_out[2 * i] = kv.second.first; counts_out[2 * i + 1] = kv.second.second; } i++; }
        if (loci_out && loci_cap > 0) { snprintf(loci_out, (size_t)loci_cap, "%s", loci.c_str()); }
        if (summary_out && summary_cap > 0) { snprintf(summary_out, (size_t)summary_cap, "%s", text.c_str()); }
        n = i; return 0;
    });
    return rc == 0 ? n : rc;
}

struct hlala_truth { TruthLevels t; };
int hlala_truth_load(const char* r1_levels, const char* r2_levels, hlala_truth_t** out) {
    if (!r1_levels || !out) return fail(HLALA_E_ARG, "hlala_truth_load: null argument");
    return guarded([&]() { std::unique_ptr<hlala_truth> h(new hlala_truth()); h->t.load(r1_levels, r2_levels ? r2_levels : ""); *out = h.release(); return 0; });
}
int64_t hlala_truth_n_reads(const hlala_truth_t* t) { return t ? (int64_t)t->t.reads.size() : (int64_t)fail(HLALA_E_ARG, "hlala_truth_n_reads: null handle"); }
int hlala_truth_evaluate(hlala_truth_t* t, int64_t n_pairs, const char* const* pair_names, int64_t pair_index_base, const hlala_pair_out_t* a, int64_t* per_pair) {
    if (!t || n_pairs < 0 || !a || !a->n_cols || !a->level || !a->schar || !a->read_reverse || a->max_columns <= 0) return fail(HLALA_E_ARG, "hlala_truth_evaluate: bad argument");
    try {
        const size_t cap = (size_t)a->max_columns;
        for (int64_t p = 0; p < n_pairs; p++) {
            if (pair_names && !pair_names[p]) return fail(HLALA_E_ARG, "hlala_truth_evaluate: a null read name");
            const std::string id = pair_names ? std::string(pair_names[p]) : "r" + std::to_string(pair_index_base + p);
            const int32_t* lv[2] = {a->level + (size_t)(2 * p) * cap, a->level + (size_t)(2 * p + 1) * cap};
            const uint8_t* sc[2] = {a->schar + (size_t)(2 * p) * cap, a->schar + (size_t)(2 * p + 1) * cap};
            const std::pair<int64_t, int64_t> r = t->t.evaluate(id, 2, lv, sc, a->n_cols + 2 * p, a->read_reverse + 2 * p);
            if (per_pair) { per_pair[2 * p] = r.first; per_pair[2 * p + 1] = r.second; }
        }
    } catch (const std::exception& e) { return fail(HLALA_E_INVARIANT, e.what()); }
    return 0;
}
int hlala_truth_totals(const hlala_truth_t* t, int64_t totals[3]) {
    if (!t || !totals) return fail(HLALA_E_ARG, "hlala_truth_totals: null argument");
    totals[0] = t->t.total; totals[1] = t->t.correct; totals[2] = t->t.reads_below_90; return 0;
}
void hlala_truth_free(hlala_truth_t* t) { delete t; }

int64_t hlala_simulate_read_pairs(const char* matrix, int32_t read_length, int32_t interpolate, int32_t drop1, int32_t drop2, const uint8_t* path, int64_t n_levels, double coverage,
                                  double diff_mean, double diff_sd, int32_t perfectly, int32_t include_deletions, const char* id_prefix, const char* out_prefix, int32_t append, double* error_rates) {
    if (!matrix || read_length <= 0 || !path || n_levels <= 0 || !out_prefix) return fail(HLALA_E_ARG, "hlala_simulate_read_pairs: bad argument");
    int64_t n = 0;
    int rc = guarded([&]() {
        ReadSimulator sim(matrix, (unsigned)read_length, interpolate != 0, drop1, drop2);
        if (error_rates) { const std::pair<double, double> e = sim.average_error_rates(); error_rates[0] = e.first; error_rates[1] = e.second; }
        const std::vector<SimulatedPair> pairs = sim.simulate_pairs_from_path(std::string((const char*)path, (size_t)n_levels), coverage, diff_mean, diff_sd, perfectly != 0,
                                                                              id_prefix ? id_prefix : "", include_deletions != 0);
        write_simulated_pairs(pairs, out_prefix, append != 0);
        n = (int64_t)pairs.size(); return 0;
    });
    return rc == 0 ? n : rc;
}

int hlala_exp_probe(int device, int64_t n, const double* x, double* y) {
    if (n < 0 || n > (1 << 30) || (n && (!x || !y))) return fail(HLALA_E_ARG, "hlala_exp_probe: bad argument");
    return guarded([&]() {
        int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(HLALA_E_CUDA, "no CUDA device available");
        CUDA_OK(cudaSetDevice(device));
        DevBuf dx, dy; dx.upload(x, (size_t)n); dy.alloc((size_t)n * 8);
        CUDA_OK(launch_exp_probe(dx.as<double>(), dy.as<double>(), (int)n, 0));
        dy.download(y, (size_t)n); CUDA_OK(cudaStreamSynchronize(0));
        return 0;
    });
}

int64_t hlala_simulate_from_graph(const hlala_graph_t* g, const char* graph_label, const char* matrix, int32_t read_length, double is_mean, double is_sd, int32_t n_genomes,
                                  const char* out_dir, double coverage, int32_t with_error, uint32_t seed) {
    if (!g || !matrix || !out_dir || read_length <= 0 || n_genomes < 0) return fail(HLALA_E_ARG, "hlala_simulate_from_graph: bad argument");
    int64_t n = 0;
    int rc = guarded([&]() { n = simulate_from_graph(g->h, graph_label ? graph_label : "", matrix, read_length, is_mean, is_sd, n_genomes, out_dir, coverage, with_error != 0, seed); return 0; });
    return rc == 0 ? n : rc;
}

int64_t hlala_simulate_individual(const char* prg_dir, const char* matrix, const char* out_dir, double is_mean, double is_sd, int32_t novel, int32_t with_error, uint32_t seed,
                                  char* types_out, int64_t types_cap) {
    if (!prg_dir || !matrix || !out_dir) return fail(HLALA_E_ARG, "hlala_simulate_individual: null argument");
    int64_t n = 0;
    int rc = guarded([&]() {
        const SimulatedIndividual r = simulate_one_individual(prg_dir, matrix, out_dir, is_mean, is_sd, novel != 0, with_error != 0, seed);
        std::string t; for (size_t i = 0; i < r.genes.size(); i++) { if (i) t += ';'; t += r.genes[i] + ":" + r.types[i].first + "/" + r.types[i].second; }
        if (types_out && types_cap > 0) snprintf(types_out, (size_t)types_cap, "%s", t.c_str());
        n = r.pairs; return 0;
    });
    return rc == 0 ? n : rc;
}

int hlala_typing_pair_probe(int device, int32_t C, int32_t R, const double* ll /* [C*R], index c*R + r */, const int32_t* mism /* [C*R] */, int termwise,
                            double* pair_ll, double* pair_mavg, double* pair_mmin, double* kernel_ms) {
    if (C <= 0 || R < 0 || !ll || !mism || !pair_ll) return fail(HLALA_E_ARG, "hlala_typing_pair_probe: bad argument");
    return guarded([&]() {
        int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(HLALA_E_CUDA, "no CUDA device available: the typing kernels have no CPU fallback");
        CUDA_OK(cudaSetDevice(device));
        TypingScoreTables t = make_typing_tables(); CUDA_OK(upload_typing_tables(t));
        const int32_t Cpad = (C + 31) / 32 * 32; const size_t npair = (size_t)C * ((size_t)C + 1) / 2;
        std::vector<double> llt((size_t)std::max(R, 1) * Cpad, 0.0); std::vector<int32_t> mmt((size_t)std::max(R, 1) * Cpad, 0);
        for (int32_t c = 0; c < C; c++) for (int32_t r = 0; r < R; r++) { llt[(size_t)r * Cpad + c] = ll[(size_t)c * R + r]; mmt[(size_t)r * Cpad + c] = mism[(size_t)c * R + r]; }
        cudaStream_t st = 0; DevBuf d_ll, d_mm, d_pair, d_e, d_rm, d_cs, d_ms;
        d_ll.upload(llt, st); d_mm.upload(mmt, st); d_pair.alloc(3 * npair * 8); d_e.alloc(llt.size() * 8); d_rm.alloc((size_t)std::max(R, 1) * 8); d_cs.alloc((size_t)C * 8); d_ms.alloc(8);
        double* pl = d_pair.as<double>(); cudaEvent_t e0, e1; CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
        for (int rep = 0; rep < 2; rep++) {     // the second run is the timed one
            CUDA_OK(cudaEventRecord(e0, st));
            if (termwise) CUDA_OK(launch_allele_pair_ll(d_ll.as<double>(), d_mm.as<int32_t>(), C, Cpad, 0, R, pl, pl + npair, pl + 2 * npair, st));
            else CUDA_OK(launch_allele_pair_prod(d_ll.as<double>(), d_mm.as<int32_t>(), C, Cpad, 0, R, d_e.as<double>(), d_rm.as<double>(), d_cs.as<double>(), d_ms.as<double>(), pl, pl + npair, pl + 2 * npair, st));
            CUDA_OK(cudaEventRecord(e1, st)); CUDA_OK(cudaStreamSynchronize(st));
        }
        float f = 0; CUDA_OK(cudaEventElapsedTime(&f, e0, e1)); if (kernel_ms) *kernel_ms = f; cudaEventDestroy(e0); cudaEventDestroy(e1);
        CUDA_OK(cudaMemcpy(pair_ll, pl, npair * 8, cudaMemcpyDeviceToHost));
        if (pair_mavg) CUDA_OK(cudaMemcpy(pair_mavg, pl + npair, npair * 8, cudaMemcpyDeviceToHost));
        if (pair_mmin) CUDA_OK(cudaMemcpy(pair_mmin, pl + 2 * npair, npair * 8, cudaMemcpyDeviceToHost));
        return 0;
    });
}

int hlala_typer_timing(const hlala_typer_t* t, double ms[2], int launches[2], double work[2]) {
    if (!t) return fail(HLALA_E_ARG, "null typer");
    for (int k = 0; k < 2; k++) { if (ms) ms[k] = t->ms[k]; if (launches) launches[k] = t->launches[k]; if (work) work[k] = t->work[k]; }
    return 0;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// k-mer seeding (reference seam B5 of SURVEY.md §8b)
struct hlala_kmer_chains {
    int64_t n_reads = 0, n_chains = 0, n_edges = 0, n_failed = 0, n_second_tier = 0;
    DevBuf chain_off, read_status, begin, end, edge_off, edges;
    double ms[3] = {0, 0, 0};
};

namespace {
void upload_kmer_index(hlala_graph* g) {
    if (g->kix_on_gpu) return;
    const KmerIndex& x = *g->kix; const FlatGraph& h = g->h;
    std::vector<int32_t> pf((size_t)x.n_pos), pt((size_t)x.n_pos);
    for (int64_t p = 0; p < x.n_pos; p++) { pf[(size_t)p] = h.edge_from[x.pos_edges[(size_t)x.pos_edge_off[(size_t)p]]]; pt[(size_t)p] = h.edge_to[x.pos_edges[(size_t)x.pos_edge_off[(size_t)p + 1] - 1]]; }
    g->kbufs.clear();
    DevKmerIndex& d = g->dkix; d.k = x.k; d.ht_mask = x.ht_mask; d.n_kmers = x.n_kmers; d.n_pos = x.n_pos;
    d.ht = g->kup(x.ht); d.kmer_bytes = g->kup(x.kmer_bytes); d.kmer_pos_off = g->kup(x.kmer_pos_off); d.pos_edge_off = g->kup(x.pos_edge_off);
    d.pos_edges = g->kup(x.pos_edges); d.pos_from = g->kup(pf); d.pos_to = g->kup(pt);
    CUDA_OK(cudaDeviceSynchronize());
    g->kix_on_gpu = true;
}
} // namespace

extern "C" {

int hlala_kmer_index_build(hlala_graph_t* g, int k) {
    if (!g) return fail(HLALA_E_ARG, "hlala_kmer_index_build: null graph");
    return guarded([&]() {
        std::unique_ptr<KmerIndex> x(new KmerIndex());
        try { build_kmer_index(g->h, k, *x); } catch (const std::exception& e) { return fail(HLALA_E_ARG, e.what()); }
        g->kix = std::move(x); g->kix_on_gpu = false; g->kbufs.clear();
        if (g->on_gpu) { CUDA_OK(cudaSetDevice(g->device)); upload_kmer_index(g); }
        return 0;
    });
}
int hlala_kmer_index_dims(const hlala_graph_t* g, int32_t* k, int64_t* n_kmers, int64_t* n_pos, int64_t* n_edges) {
    if (!g || !g->kix) return fail(HLALA_E_ARG, "hlala_kmer_index_dims: no k-mer index (call hlala_kmer_index_build)");
    if (k) *k = g->kix->k; if (n_kmers) *n_kmers = g->kix->n_kmers; if (n_pos) *n_pos = g->kix->n_pos; if (n_edges) *n_edges = (int64_t)g->kix->pos_edges.size();
    return 0;
}
int hlala_kmer_index_export(const hlala_graph_t* g, uint8_t* kmer_bytes, int64_t* pos_off, int64_t* edge_off, int32_t* edges) {
    if (!g || !g->kix) return fail(HLALA_E_ARG, "hlala_kmer_index_export: no k-mer index (call hlala_kmer_index_build)");
    const KmerIndex& x = *g->kix;
    if (kmer_bytes) memcpy(kmer_bytes, x.kmer_bytes.data(), x.kmer_bytes.size());
    if (pos_off) memcpy(pos_off, x.kmer_pos_off.data(), x.kmer_pos_off.size() * 8);
    if (edge_off) memcpy(edge_off, x.pos_edge_off.data(), x.pos_edge_off.size() * 8);
    if (edges) for (size_t i = 0; i < x.pos_edges.size(); i++) edges[i] = g->h.edge_ord[(size_t)x.pos_edges[i]];
    return 0;
}

int hlala_seed_kmers(hlala_graph_t* g, const hlala_read_batch_t* reads, hlala_kmer_chains_t** out) {
    if (!g || !reads || !out) return fail(HLALA_E_ARG, "hlala_seed_kmers: null argument");
    *out = nullptr;
    if (reads->n_reads < 0 || (reads->n_reads > 0 && (!reads->read_off || !reads->bases))) return fail(HLALA_E_ARG, "hlala_seed_kmers: bad read batch");
    if (!g->kix) return fail(HLALA_E_ARG, "hlala_seed_kmers: no k-mer index (call hlala_kmer_index_build)");
    if (!g->on_gpu) return fail(HLALA_E_CUDA, "graph is not on a GPU: call hlala_graph_to_gpu first (there is no CPU fallback)");
    if (reads->n_reads >= ((int64_t)1 << 31) - 1) return fail(HLALA_E_ARG, "hlala_seed_kmers: at most 2^31-2 reads per call");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(g->device));
        upload_kmer_index(g);
        cudaStream_t st = 0;
        std::unique_ptr<hlala_kmer_chains> R(new hlala_kmer_chains());
        const int64_t nr = reads->n_reads; R->n_reads = nr;
        const int64_t nb = nr ? reads->read_off[nr] : 0; int64_t maxL = 0;
        for (int64_t r = 0; r < nr; r++) maxL = std::max<int64_t>(maxL, reads->read_off[r + 1] - reads->read_off[r]);
        DevBuf d_off, d_bases, n_chains, recs, pool, counts, counters, chain_scratch, scan_scratch, ordered, ord_ne, temp, defer_list, defer_count, read_tier, big_chain_scratch, big_scan_scratch;
        std::vector<int64_t> off0(1, 0);
        d_off.upload(nr ? reads->read_off : off0.data(), (size_t)nr + 1, st); d_bases.upload(reads->bases, (size_t)nb, st);
        n_chains.alloc((size_t)std::max<int64_t>(nr, 1) * 4); R->read_status.alloc((size_t)std::max<int64_t>(nr, 1) * 4);
        SeedParams P{}; P.g = g->d; P.ix = g->dkix; P.n_reads = nr; P.read_off = d_off.as<int64_t>(); P.bases = d_bases.as<uint8_t>();
        P.ecap = (int32_t)std::max<int64_t>(512, 3 * maxL);
        const int warps = (int)std::min<int64_t>(seed_warps_for(g->n_sm), std::max<int64_t>(SEED_WARPS, ((nr + SEED_WARPS - 1) / SEED_WARPS) * SEED_WARPS));
        chain_scratch.alloc((size_t)warps * SEED_RC * (size_t)P.ecap * 4); scan_scratch.alloc((size_t)warps * seed_scan_scratch_ints() * 4);
        P.chain_scratch = chain_scratch.as<int32_t>(); P.scan_scratch = scan_scratch.as<int32_t>();
        P.read_n_chains = n_chains.as<int32_t>(); P.read_status = R->read_status.as<int32_t>();
        counts.alloc(16); counters.alloc(16); defer_list.alloc((size_t)std::max<int64_t>(nr, 1) * 4); defer_count.alloc(4); read_tier.alloc((size_t)std::max<int64_t>(nr, 1));
        P.defer_list = getenv("HLALA_SEED_NO_BIG_TIER") ? nullptr : defer_list.as<int32_t>(); P.defer_count = defer_count.as<int32_t>(); P.read_tier = read_tier.as<uint8_t>();     // (the variable is an A/B and test hook)
        P.rec_count = counts.as<unsigned long long>(); P.edge_count = counts.as<unsigned long long>() + 1; P.counters = counters.as<int32_t>();
        int64_t rec_cap = nr * 3 + 1024, edge_cap = nb * 3 + 65536;
        cudaEvent_t e0, e1, e2; CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1)); CUDA_OK(cudaEventCreate(&e2));
        unsigned long long cnt[2] = {0, 0}; int32_t ctr[4] = {0, 0, 0, 0};
        for (int attempt = 0; ; attempt++) {
            recs.alloc((size_t)rec_cap * sizeof(SeedChainRec)); pool.alloc((size_t)edge_cap * 4);
            P.recs = recs.as<SeedChainRec>(); P.rec_cap = rec_cap; P.edge_pool = pool.as<int32_t>(); P.edge_cap = edge_cap;
            CUDA_OK(cudaMemsetAsync(counts.p, 0, 16, st)); CUDA_OK(cudaMemsetAsync(counters.p, 0, 16, st)); CUDA_OK(cudaMemsetAsync(defer_count.p, 0, 4, st));
            CUDA_OK(cudaEventRecord(e0, st));
            CUDA_OK(launch_seed_chains(P, g->n_sm, st));
            int32_t n_def = 0; defer_count.download(&n_def, 1, st); CUDA_OK(cudaStreamSynchronize(st));
            if (n_def > 0) {     // second tier: 1024 running chains per read, for the reads the first tier queued
                SeedParams Q = P; const int bw = seed_big_warps_for(g->n_sm);
                big_chain_scratch.alloc((size_t)bw * SEED_RC_BIG * (size_t)P.ecap * 4); big_scan_scratch.alloc((size_t)bw * seed_scan_scratch_ints() * 4);
                Q.chain_scratch = big_chain_scratch.as<int32_t>(); Q.scan_scratch = big_scan_scratch.as<int32_t>();
                CUDA_OK(launch_seed_chains_big(Q, g->n_sm, st));
            }
            R->n_second_tier = n_def;
            CUDA_OK(cudaEventRecord(e1, st));
            counts.download(cnt, 2, st); counters.download(ctr, 4, st);
            CUDA_OK(cudaStreamSynchronize(st));
            if (!ctr[2]) break;
            if (attempt >= 8) return fail(HLALA_E_CAPACITY, "hlala_seed_kmers: chain output does not fit after 8 capacity doublings");
            rec_cap = std::max<int64_t>(rec_cap * 2, (int64_t)cnt[0] + 1024); edge_cap = std::max<int64_t>(edge_cap * 2, (int64_t)cnt[1] + 65536);   // the counters keep counting past the capacity
        }
        R->n_failed = ctr[1];
        // (read, order) order + contiguous edges as canonical ordinals
        R->chain_off.alloc((size_t)(nr + 1) * 8);
        size_t need = 0; CUDA_OK(seed_exclusive_scan_i32_to_i64(n_chains.as<int32_t>(), nr, R->chain_off.as<long long>(), nullptr, 0, &need, st));
        temp.alloc(need + 16);
        CUDA_OK(seed_exclusive_scan_i32_to_i64(n_chains.as<int32_t>(), nr, R->chain_off.as<long long>(), temp.p, need, &need, st));
        long long total_chains = 0; CUDA_OK(cudaMemcpyAsync(&total_chains, R->chain_off.as<long long>() + nr, 8, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
        R->n_chains = total_chains;
        ordered.alloc((size_t)std::max<long long>(total_chains, 1) * sizeof(SeedChainRec)); ord_ne.alloc((size_t)std::max<long long>(total_chains, 1) * 4);
        CUDA_OK(launch_seed_order(recs.as<SeedChainRec>(), (long long)cnt[0], R->chain_off.as<long long>(), R->read_status.as<int32_t>(), read_tier.as<uint8_t>(), ordered.as<SeedChainRec>(), ord_ne.as<int32_t>(), st));
        R->edge_off.alloc((size_t)(total_chains + 1) * 8);
        size_t need2 = 0; CUDA_OK(seed_exclusive_scan_i32_to_i64(ord_ne.as<int32_t>(), total_chains, R->edge_off.as<long long>(), nullptr, 0, &need2, st));
        if (need2 > need) { temp.alloc(need2 + 16); }
        CUDA_OK(seed_exclusive_scan_i32_to_i64(ord_ne.as<int32_t>(), total_chains, R->edge_off.as<long long>(), temp.p, std::max(need, need2), &need2, st));
        long long total_edges = 0; CUDA_OK(cudaMemcpyAsync(&total_edges, R->edge_off.as<long long>() + total_chains, 8, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
        R->n_edges = total_edges;
        R->begin.alloc((size_t)std::max<long long>(total_chains, 1) * 4); R->end.alloc((size_t)std::max<long long>(total_chains, 1) * 4); R->edges.alloc((size_t)std::max<long long>(total_edges, 1) * 4);
        CUDA_OK(launch_seed_gather(g->d, ordered.as<SeedChainRec>(), total_chains, R->edge_off.as<long long>(), pool.as<int32_t>(), R->begin.as<int32_t>(), R->end.as<int32_t>(), R->edges.as<int32_t>(), st));
        CUDA_OK(cudaEventRecord(e2, st)); CUDA_OK(cudaEventSynchronize(e2));
        float f1 = 0, f2 = 0; CUDA_OK(cudaEventElapsedTime(&f1, e0, e1)); CUDA_OK(cudaEventElapsedTime(&f2, e1, e2));
        R->ms[0] = f1; R->ms[1] = f2; R->ms[2] = 0;
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        *out = R.release();
        return 0;
    });
}
int hlala_kmer_chains_second_tier_reads(const hlala_kmer_chains_t* c, int64_t* n) { if (!c || !n) return fail(HLALA_E_ARG, "hlala_kmer_chains_second_tier_reads: null argument"); *n = c->n_second_tier; return 0; }
int hlala_kmer_chains_dims(const hlala_kmer_chains_t* c, int64_t* n_reads, int64_t* n_chains, int64_t* n_edges, int64_t* n_failed_reads) {
    if (!c) return fail(HLALA_E_ARG, "hlala_kmer_chains_dims: null result");
    if (n_reads) *n_reads = c->n_reads; if (n_chains) *n_chains = c->n_chains; if (n_edges) *n_edges = c->n_edges; if (n_failed_reads) *n_failed_reads = c->n_failed;
    return 0;
}
int hlala_kmer_chains_fetch(const hlala_kmer_chains_t* c, int64_t* chain_off, int32_t* read_status, int32_t* seq_begin, int32_t* seq_end, int64_t* edge_off, int32_t* edges) {
    if (!c) return fail(HLALA_E_ARG, "hlala_kmer_chains_fetch: null result");
    return guarded([&]() {
        if (chain_off) c->chain_off.download(chain_off, (size_t)c->n_reads + 1); if (read_status) c->read_status.download(read_status, (size_t)c->n_reads);
        if (seq_begin) c->begin.download(seq_begin, (size_t)c->n_chains); if (seq_end) c->end.download(seq_end, (size_t)c->n_chains);
        if (edge_off) c->edge_off.download(edge_off, (size_t)c->n_chains + 1); if (edges) c->edges.download(edges, (size_t)c->n_edges);
        CUDA_OK(cudaStreamSynchronize(0));
        return 0;
    });
}
int hlala_kmer_chains_timing(const hlala_kmer_chains_t* c, double ms[2]) {
    if (!c || !ms) return fail(HLALA_E_ARG, "hlala_kmer_chains_timing: null argument");
    ms[0] = c->ms[0]; ms[1] = c->ms[1]; return 0;
}
void hlala_kmer_chains_free(hlala_kmer_chains_t* c) { delete c; }

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// BAM ingest and the session's own coverage histogram (what the CLI needs on top of the kernels)
struct hlala_bam_batch { BamBatch b; std::vector<const char*> names; };

extern "C" {

static int bam_read_impl(const hlala_graph_t* g, const char* bam_path, int threads, bool long_reads, hlala_bam_batch_t** out) {
    if (!g || !bam_path || !out) return fail(HLALA_E_ARG, "hlala_bam_read: null argument");
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<hlala_bam_batch> B(new hlala_bam_batch());
        std::vector<int64_t> len; for (int32_t c = 0; c < g->h.n_contigs; c++) len.push_back(g->h.contig_off[(size_t)c + 1] - g->h.contig_off[(size_t)c]);
        read_bam_seeds(bam_path, g->h.contig_bam_name, len, threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency()), B->b, long_reads);
        if (long_reads) {     // hand the batch out in the single-read form hlala_align_long_reads takes: the empty second reads carry no bases and no chains
            BamBatch& b = B->b; const size_t n = b.pair_name.size();
            std::vector<int64_t> ro(n + 1); std::vector<int32_t> co(n + 1);
            for (size_t r = 0; r <= n; r++) { ro[r] = b.read_off[2 * r]; co[r] = b.chain_off[2 * r]; }
            b.read_off.swap(ro); b.chain_off.swap(co);
        }
        for (const std::string& n : B->b.pair_name) B->names.push_back(n.c_str());
        *out = B.release();
        return 0;
    });
}
int hlala_bam_read(const hlala_graph_t* g, const char* bam_path, int threads, hlala_bam_batch_t** out) { return bam_read_impl(g, bam_path, threads, false, out); }
int hlala_fastq_map_pairs(const hlala_graph_t* g, const char* fastq1, const char* fastq2, int threads, hlala_bam_batch_t** out) {
    if (!g || !fastq1 || !fastq2 || !out) return fail(HLALA_E_ARG, "hlala_fastq_map_pairs: null argument");
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<hlala_bam_batch> B(new hlala_bam_batch());
        MapperParams mp;   // tuning hooks for experiments (defaults are what the tests and the profile describe)
        if (const char* e = getenv("HLALA_MAPPER_READ_STEP")) mp.read_step = std::max(1, atoi(e));
        if (const char* e = getenv("HLALA_MAPPER_REF_STEP")) mp.ref_step = std::max(1, atoi(e));
        if (const char* e = getenv("HLALA_MAPPER_K")) mp.k = atoi(e);
        map_fastq_pairs(g->h, fastq1, fastq2, threads, mp, B->b);
        B->names.reserve(B->b.pair_name.size()); for (const std::string& n : B->b.pair_name) B->names.push_back(n.c_str());
        *out = B.release(); return 0;
    });
}
int hlala_bam_read_long(const hlala_graph_t* g, const char* bam_path, int threads, hlala_bam_batch_t** out) { return bam_read_impl(g, bam_path, threads, true, out); }
int hlala_bam_batch_view(const hlala_bam_batch_t* B, hlala_seed_batch_t* view, const char* const** pair_names) {
    if (!B || !view) return fail(HLALA_E_ARG, "hlala_bam_batch_view: null argument");
    const BamBatch& b = B->b;
    view->n_reads = (int64_t)b.read_off.size() - 1; view->read_off = b.read_off.data(); view->bases = b.bases.data(); view->quals = b.quals.data();
    view->chain_off = b.chain_off.data(); view->chain_contig = b.chain_contig.data(); view->chain_pos = b.chain_pos.data(); view->chain_flag = b.chain_flag.data(); view->chain_as = b.chain_as.data();
    view->cigar_off = b.cigar_off.data(); view->cigar = b.cigar.data();
    if (pair_names) *pair_names = B->names.data();
    return 0;
}
int hlala_bam_batch_stats(const hlala_bam_batch_t* B, int64_t counts[4], double* is_mean, double* is_sd, int64_t* is_n) {
    if (!B) return fail(HLALA_E_ARG, "hlala_bam_batch_stats: null argument");
    if (counts) { counts[0] = B->b.records; counts[1] = B->b.records_used; counts[2] = B->b.names_seen; counts[3] = B->b.pairs_incomplete; }
    if (is_mean) *is_mean = B->b.tlen_mean; if (is_sd) *is_sd = B->b.tlen_sd; if (is_n) *is_n = B->b.tlen_n;
    return 0;
}
void hlala_bam_batch_free(hlala_bam_batch_t* B) { delete B; }

// ---- the reference's insert-size estimate (processBAM::estimateInsertSize, processBAM.cpp:1071-1181)
int hlala_bam_insert_size_sample(const hlala_bam_batch_t* B, hlala_seed_batch_t* view, const int32_t** loaded_contigs, int32_t* n_loaded) {
    if (!B || !view) return fail(HLALA_E_ARG, "hlala_bam_insert_size_sample: null argument");
    const BamBatch::Sample& S = B->b.is_sample;
    memset(view, 0, sizeof *view);
    view->n_reads = (int64_t)S.read_off.size() - 1; if (view->n_reads < 0) view->n_reads = 0;
    view->read_off = S.read_off.data(); view->bases = S.bases.data(); view->quals = S.quals.data(); view->chain_off = S.chain_off.data();
    view->chain_contig = S.chain_contig.data(); view->chain_pos = S.chain_pos.data(); view->chain_flag = S.chain_flag.data(); view->chain_as = S.chain_as.data();
    view->cigar_off = S.cigar_off.data(); view->cigar = S.cigar.data();
    if (loaded_contigs) *loaded_contigs = S.loaded_contigs.data(); if (n_loaded) *n_loaded = (int32_t)S.loaded_contigs.size();
    return 0;
}
int hlala_insert_size_from_levels(const hlala_graph_t* g, int64_t n_pairs, const int32_t* first_level, const int32_t* last_level, const uint8_t* reverse,
                                  const int32_t* loaded_contigs, int32_t n_loaded, double* mean, double* sd, int64_t* used, int64_t* skipped) {
    if (!g || !first_level || !last_level || !reverse || !mean || !sd) return fail(HLALA_E_ARG, "hlala_insert_size_from_levels: null argument");
    return guarded([&]() { InsertSizeEstimate e = estimate_insert_size(g->h, n_pairs, first_level, last_level, reverse, loaded_contigs, n_loaded);
        *mean = e.mean; *sd = e.sd; if (used) *used = e.used; if (skipped) *skipped = e.skipped; return 0; });
}
int hlala_bam_insert_size(hlala_graph_t* g, const hlala_bam_batch_t* B, int32_t max_columns, double* mean, double* sd, int64_t* used, int64_t* skipped) {
    if (!g || !B || !mean || !sd) return fail(HLALA_E_ARG, "hlala_bam_insert_size: null argument");
    if (!g->on_gpu) return fail(HLALA_E_CUDA, "graph is not on a GPU: call hlala_graph_to_gpu first (there is no CPU fallback)");
    if (max_columns < 32 || max_columns > 2040) return fail(HLALA_E_ARG, "max_columns must be in [32, 2040]");
    hlala_seed_batch_t v; const int32_t* loaded = nullptr; int32_t n_loaded = 0;
    if (int rc = hlala_bam_insert_size_sample(B, &v, &loaded, &n_loaded)) return rc;
    if (v.n_reads < 2) return fail(HLALA_E_INVARIANT, "insert size: the BAM holds no complete read pair in its first 4000 read names");
    return guarded([&]() {
        CUDA_OK(cudaSetDevice(g->device)); cudaStream_t st = 0;
        // the primary records of the sample through the chain kernels and the extension DP, every chain aligned (one chain per read)
        Pipeline pl; pl.scratch_budget = (size_t)1 << 62; pl.allow_env_budget = false; pl.dedup = false; pl.prepare(g, v, max_columns, st);
        pl.begin_run(st); pl.fork_lanes(st);
        if (pl.wave_pair.size() > 1) pl.run_chains_wave(0, *pl.lanes[0]);
        pl.join_lanes(st);
        const int32_t nc = pl.pb.n_chains; std::vector<int32_t> status((size_t)nc), fl((size_t)nc), ll((size_t)nc);
        pl.cs.status.download(status.data(), (size_t)nc, st); pl.cs.first_level.download(fl.data(), (size_t)nc, st); pl.cs.last_level.download(ll.data(), (size_t)nc, st);
        CUDA_OK(cudaStreamSynchronize(st));
        std::vector<uint8_t> rev((size_t)nc);
        for (int32_t c = 0; c < nc; c++) { if (status[(size_t)c] != CH_OK) return fail(status[(size_t)c] < 0 ? status[(size_t)c] : HLALA_E_INVARIANT, "insert size: a primary record of the sample could not be aligned (status " + std::to_string(status[(size_t)c]) + ")");
            rev[(size_t)c] = (v.chain_flag[c] & 0x10) ? 1 : 0; }     // slot == chain: one chain per read
        InsertSizeEstimate e = estimate_insert_size(g->h, v.n_reads / 2, fl.data(), ll.data(), rev.data(), loaded, n_loaded);
        *mean = e.mean; *sd = e.sd; if (used) *used = e.used; if (skipped) *skipped = e.skipped;
        return 0;
    });
}

void hlala_graph_release_workspace(hlala_graph_t* g) { if (!g) return; std::lock_guard<std::mutex> lock(g->ws_mu); g->ws.reset(); }

int hlala_session_set_coverage(hlala_session_t* s, int on) { if (!s) return fail(HLALA_E_ARG, "null session"); s->own_cov = on != 0; return 0; }
int hlala_session_fetch_coverage(hlala_session_t* s, int32_t* bases_per_level) {
    if (!s || !bases_per_level) return fail(HLALA_E_ARG, "hlala_session_fetch_coverage: null argument");
    if (!s->cov.p) return fail(HLALA_E_ARG, "hlala_session_fetch_coverage: run the session with hlala_session_set_coverage(s, 1) first");
    return guarded([&]() { CUDA_OK(cudaSetDevice(s->pl.g->device)); CUDA_OK(cudaMemcpy(bases_per_level, s->cov.p, (size_t)std::max(s->pl.g->h.n_levels - 1, 1) * 4, cudaMemcpyDeviceToHost)); return 0; });
}

} // extern "C"
