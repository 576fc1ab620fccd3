// extern "C" boundary of libhlala_b200.so (include/hlala_b200.h). Host logic in C++, compute in CUDA; no CPU fallback.
#include "../../include/hlala_b200.h"
#include "../host/prg_graph.h"
#include "chain_params.h"
#include "align_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using namespace hlala;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

struct CudaError : std::runtime_error { explicit CudaError(const std::string& m) : std::runtime_error(m) {} };
#define CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void alloc(size_t n) { if (p) { cudaFree(p); p = nullptr; } bytes = n; if (n) CUDA_OK(cudaMalloc(&p, n)); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    template <class T> void upload(const T* h, size_t count, cudaStream_t st = 0) { alloc(count * sizeof(T)); if (count) CUDA_OK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st)); }
    template <class T> void upload(const std::vector<T>& v, cudaStream_t st = 0) { upload(v.data(), v.size(), st); }
    template <class T> void download(T* h, size_t count, cudaStream_t st = 0) const { if (count) CUDA_OK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, st)); }
};

// Score tables with the reference's own arithmetic (Utilities::PhredToPCorrect Utilities.cpp:357-377,
// extensionAligner::scoreOneAlignment extensionAligner.cpp:58-66,127-146), evaluated by the host libm once.
ScoreTables make_score_tables() {
    ScoreTables t;
    double rate_deletions = log(0.001), rate_insertions = log(0.001);
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
        std::vector<int32_t> idx;
        for (int64_t r = 0; r < n_reads; r++) {
            int32_t c0 = b.chain_off[r], c1 = b.chain_off[r + 1];
            idx.resize(c1 - c0); for (int32_t i = 0; i < c1 - c0; i++) idx[i] = c0 + i;
            std::sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return b.chain_as[x] < b.chain_as[y]; });
            std::reverse(idx.begin(), idx.end());
            for (int32_t i = 0; i < c1 - c0; i++) {
                chain_order[c0 + i] = idx[i]; slot_read[c0 + i] = (int32_t)r;
                if (read_primary[r] < 0 && !(b.chain_flag[idx[i]] & 0x100)) read_primary[r] = c0 + i;
            }
        }
    }
};

struct DeviceBatch {
    DevBuf read_off, bases, quals, chain_off, chain_contig, chain_pos, chain_flag, chain_as, cigar_off, cigar, chain_order, read_primary, slot_read;
    DevBatch view{};
    void upload(const hlala_seed_batch_t& b, const PreparedBatch& pb, cudaStream_t st) {
        int64_t nb = b.read_off[b.n_reads]; int32_t nc = pb.n_chains; int32_t ncg = b.cigar_off[nc];
        read_off.upload(b.read_off, (size_t)b.n_reads + 1, st); bases.upload(b.bases, (size_t)nb, st); quals.upload(b.quals, (size_t)nb, st);
        chain_off.upload(b.chain_off, (size_t)b.n_reads + 1, st);
        chain_contig.upload(b.chain_contig, (size_t)nc, st); chain_pos.upload(b.chain_pos, (size_t)nc, st); chain_flag.upload(b.chain_flag, (size_t)nc, st); chain_as.upload(b.chain_as, (size_t)nc, st);
        cigar_off.upload(b.cigar_off, (size_t)nc + 1, st); cigar.upload(b.cigar, (size_t)ncg, st);
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

struct ChainScratch {
    DevBuf status, n_cols, seed_begin, seed_end, ll, first_level, last_level, c_edge, c_schar, c_fromseed, error_count;
    void alloc(int32_t n_chains, int32_t maxcol) {
        size_t nc = (size_t)std::max(n_chains, 1);
        status.alloc(nc * 4); n_cols.alloc(nc * 4); seed_begin.alloc(nc * 4); seed_end.alloc(nc * 4); ll.alloc(nc * 8); first_level.alloc(nc * 4); last_level.alloc(nc * 4);
        c_edge.alloc(nc * maxcol * 4); c_schar.alloc(nc * maxcol); c_fromseed.alloc(nc * maxcol); error_count.alloc(4);
    }
    void fill(ChainParams& P) {
        P.status = status.as<int32_t>(); P.n_cols = n_cols.as<int32_t>(); P.seed_begin = seed_begin.as<int32_t>(); P.seed_end = seed_end.as<int32_t>(); P.ll = ll.as<double>();
        P.first_level = first_level.as<int32_t>(); P.last_level = last_level.as<int32_t>(); P.c_edge = c_edge.as<int32_t>(); P.c_schar = c_schar.as<uint8_t>(); P.c_fromseed = c_fromseed.as<uint8_t>();
        P.error_count = error_count.as<int32_t>();
    }
};

void chain_caps(int32_t maxcol, ChainParams& P) { P.maxcol = maxcol; P.pool_cap = 4 * maxcol; P.win_cap = 4 * maxcol; }

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
        std::vector<int32_t> leo(h.level_edge_off); leo.resize((size_t)h.n_levels + 1, h.n_edges);
        g->bufs.clear();
        DevGraph& d = g->d;
        d.n_levels = h.n_levels; d.n_nodes = h.n_nodes; d.n_edges = h.n_edges; d.n_contigs = h.n_contigs;
        d.level_node_off = g->up(h.level_node_off); d.level_edge_off = g->up(leo); d.edge_pack = g->up(pack); d.edge_ord = g->up(h.edge_ord);
        d.node_out_off = g->up(h.node_out_off); d.node_out = g->up(h.node_out); d.node_in_off = g->up(h.node_in_off); d.node_in = g->up(h.node_in);
        d.path_off = g->up(h.path_off); d.path_edges = g->up(h.path_edges); d.path_from = g->up(h.path_from); d.path_to = g->up(h.path_to);
        d.jump_fwd_off = g->up(h.jump_fwd_off); d.jump_fwd_path = g->up(h.jump_fwd_path); d.jump_bwd_off = g->up(h.jump_bwd_off); d.jump_bwd_path = g->up(h.jump_bwd_path);
        d.gap_stretch = g->up(h.gap_stretch); d.contig_off = g->up(h.contig_off); d.contig_seq = g->up(h.contig_seq); d.contig_level = g->up(h.contig_level);
        d.contig_prg_id = g->up(h.contig_prg_id); d.anchor_off = g->up(h.anchor_off); d.anchor_prg_id = g->up(h.anchor_prg_id); d.anchor_pos = g->up(h.anchor_pos);
        ScoreTables t = make_score_tables();
        CUDA_OK(upload_score_tables(t));
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
        PreparedBatch pb; pb.build(*batch);
        DeviceBatch db; db.upload(*batch, pb, st);
        const int32_t nc = pb.n_chains, mc = out->max_columns;
        ChainScratch cs; cs.alloc(nc, mc);
        CUDA_OK(cudaMemsetAsync(cs.error_count.p, 0, 4, st));
        CUDA_OK(cudaMemsetAsync(cs.status.p, 0xFF, (size_t)std::max(nc, 1) * 4, st));
        ChainParams P{}; P.g = g->d; P.b = db.view; chain_caps(mc, P); P.do_extension = 1; cs.fill(P);
        CUDA_OK(launch_chain_seed(P, g->n_sm, st));
        DevBuf o_level, o_edge, o_g; size_t ncol = (size_t)std::max(nc, 1) * mc;
        o_level.alloc(ncol * 4); o_edge.alloc(ncol * 4); o_g.alloc(ncol);
        CUDA_OK(launch_export_chain_columns(g->d, nc, mc, P.n_cols, P.first_level, P.c_edge, o_level.as<int32_t>(), o_edge.as<int32_t>(), o_g.as<uint8_t>(), st));
        if (out->chain_order) memcpy(out->chain_order, pb.chain_order.data(), (size_t)nc * 4);
        if (out->status) cs.status.download(out->status, nc, st);
        if (out->n_cols) cs.n_cols.download(out->n_cols, nc, st);
        if (out->seed_begin) cs.seed_begin.download(out->seed_begin, nc, st);
        if (out->seed_end) cs.seed_end.download(out->seed_end, nc, st);
        if (out->ll) cs.ll.download(out->ll, nc, st);
        if (out->level) o_level.download(out->level, (size_t)nc * mc, st);
        if (out->edge) o_edge.download(out->edge, (size_t)nc * mc, st);
        if (out->gchar) o_g.download(out->gchar, (size_t)nc * mc, st);
        if (out->schar) cs.c_schar.download(out->schar, (size_t)nc * mc, st);
        if (out->from_seed) cs.c_fromseed.download(out->from_seed, (size_t)nc * mc, st);
        CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    });
}

} // extern "C"
