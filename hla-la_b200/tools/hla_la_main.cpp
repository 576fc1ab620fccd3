// hlala-b200: command-line front end for the accelerated path, with the argument names of the reference binary's `--action HLA`
// (HLA-LA.cpp:196-210: --action --sampleID --BAM --outputDirectory --PRG_graph_dir). It is the seam B0 of SURVEY.md §8b:
//   void processBAM::alignReads_and_inferHLA(BAM, 0, IS_mean, IS_sd, outputDirectory, false, &HLAtyper, 1, "")   mapper/processBAM.h:132
// called at HLA-LA.cpp:799 — i.e. it takes the REMAPPED BAM (bwa mem -a -M against PRG_graph_dir/mapping_PRGonly/referenceGenome.fa; the
// read extraction and the bwa/samtools steps before it belong to the Perl/C++ control plane that is out of scope) and writes
//   <outputDirectory>/reads_per_level.txt         processBAM.cpp:1907-1913
//   <outputDirectory>/hla/*                        hla::HLATyper::HLATypeInference, hla/HLATyper.cpp:933-2810
// Everything goes through the C ABI (include/hlala_b200.h); there is no CPU fallback.
//
// --gpus N (N > 1): the read pairs are sharded in contiguous blocks over N GPUs of this node (one host thread, one replica of the flat graph and
// one NCCL communicator per GPU). The exchanges are the ones of SURVEY.md §8e: one ncclAllReduce(int32, sum) of the per-level coverage, the
// gene-overlapping alignment blobs of all shards handed to every rank (host memory of this process), and ONE ncclAllReduce(double, sum) of
// the allele-pair vectors per locus behind the callback of hlala_typer_infer. Rank 0 writes the files.
#include "../../include/hlala_b200.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <atomic>
#include <thread>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <sys/stat.h>
#include <vector>
#include <algorithm>

static std::chrono::steady_clock::time_point g_t0;
static void phase(const char* what) {   // wall time of every phase, on stdout
    const auto t1 = std::chrono::steady_clock::now(); fprintf(stdout, "hlala-b200: [%8.2f s] %s\n", std::chrono::duration<double>(t1 - g_t0).count(), what); g_t0 = t1;
}
static int die(const char* what) { fprintf(stderr, "hlala-b200: %s: %s\n", what, hlala_last_error()); return 1; }

// ---- multi-GPU -------------------------------------------------------------------------------------------------------------
namespace {
struct Shard {     // contiguous block of pairs of the BAM batch, offsets rebased to zero
    std::vector<int64_t> read_off; std::vector<int32_t> chain_off, cigar_off; hlala_seed_batch_t view; int64_t pair0 = 0;
    void cut(const hlala_seed_batch_t& b, int64_t p0, int64_t p1) {
        pair0 = p0; const int64_t r0 = 2 * p0, r1 = 2 * p1; const int32_t c0 = b.chain_off[r0], c1 = b.chain_off[r1]; const int32_t g0 = b.cigar_off[c0]; const int64_t q0 = b.read_off[r0];
        read_off.resize((size_t)(r1 - r0 + 1)); for (int64_t r = r0; r <= r1; r++) read_off[(size_t)(r - r0)] = b.read_off[r] - q0;
        chain_off.resize((size_t)(r1 - r0 + 1)); for (int64_t r = r0; r <= r1; r++) chain_off[(size_t)(r - r0)] = b.chain_off[r] - c0;
        cigar_off.resize((size_t)(c1 - c0 + 1)); for (int32_t c = c0; c <= c1; c++) cigar_off[(size_t)(c - c0)] = b.cigar_off[c] - g0;
        view = b; view.n_reads = r1 - r0; view.read_off = read_off.data(); view.bases = b.bases + q0; view.quals = b.quals + q0; view.chain_off = chain_off.data();
        view.chain_contig = b.chain_contig + c0; view.chain_pos = b.chain_pos + c0; view.chain_flag = b.chain_flag + c0; view.chain_as = b.chain_as + c0;
        view.cigar_off = cigar_off.data(); view.cigar = b.cigar + g0;
    }
};
struct Barrier {   // all ranks of this process meet here (C++17 has no std::barrier)
    std::atomic<int> count{0}; std::atomic<int> gen{0}; int n = 1;
    void wait() { const int g = gen.load(); if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); } else while (gen.load() == g) std::this_thread::yield(); }
};
struct RankCtx { ncclComm_t comm; int device; };
int nccl_sum_f64(void* ctx, uint64_t dev_ptr, int64_t count, void* stream) {     // hlala_allreduce_f64_fn: the one collective of the typing stage
    RankCtx* c = (RankCtx*)ctx; cudaSetDevice(c->device); cudaStream_t st = (cudaStream_t)stream;
    if (ncclAllReduce((const void*)dev_ptr, (void*)dev_ptr, (size_t)count, ncclDouble, ncclSum, c->comm, st) != ncclSuccess) return 1;
    return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}
}

static int evaluate_calls(std::map<std::string, std::string>& a, const std::string& out_dir);
// An alignment has one column per read base plus one per graph level it spans without a base (deletions, gap stretches of hundreds of levels): 640 columns hold
// 2 x 150 bp reads with room to spare (1 M pairs on the bench PRG: none beyond); at 2 x 250 bp a handful of 1 M pairs passed 689, so longer reads get 4 x their
// length (an alignment beyond max_columns is a reported capacity error, never a truncation).
static int default_max_columns(const hlala_seed_batch_t& b) {
    int64_t longest = 0; for (int64_t r = 0; r < b.n_reads; r++) longest = std::max<int64_t>(longest, b.read_off[r + 1] - b.read_off[r]);
    return longest <= 160 ? 640 : (int)std::min<int64_t>(2040, 4 * longest + 64);
}
// the seeds of the run: the remapped BAM (HLA-LA.cpp:775-779), or paired FASTQ files placed on the PRG contigs here (in place of BWAmapper::map, HLA-LA.cpp:742)
static int read_seeds(std::map<std::string, std::string>& a, hlala_graph_t* g, hlala_bam_batch_t** bam) {
    const int threads = a.count("threads") ? atoi(a["threads"].c_str()) : 0;
    if (a.count("BAM")) {
        if (hlala_bam_read(g, a["BAM"].c_str(), threads, bam)) return die("reading the BAM");
        phase("BAM read, records selected and grouped");
    } else {
        // HLA-LA.cpp:770: --mapAgainstCompleteGenome 1 maps against the whole genome plus the PRG; here the PRG contigs are all there is (extract the MHC reads first, as HLA-LA.pl does)
        if (a.count("mapAgainstCompleteGenome") && a["mapAgainstCompleteGenome"] != "0") fprintf(stderr, "hlala-b200: note: --mapAgainstCompleteGenome is not available without bwa; the reads are placed on the PRG contigs only\n");
        if (hlala_fastq_map_pairs(g, a["FASTQ1"].c_str(), a["FASTQ2"].c_str(), threads, bam)) return die("mapping the FASTQ files");
        phase("FASTQ read, reads placed on the PRG contigs (all placements, best one primary)");
    }
    return 0;
}
// --action testPRGMapping: the reference's synthetic integration test (HLA-LA.cpp:1386-1621, e.g. :1455-1527): reads with known levels from a random genome of the graph
// (simulator::simulateFromGraph) -> mapped (there: bwa; here: hlala_fastq_map_pairs) -> alignReads with a trueReadLevels object -> "Graph: <bases> <fraction on their true level>".
// The reference simulates its own small PRG first (Graph/graphSimulator); here the PRG is the one given by --PRG_graph_dir.
static int run_test_prg_mapping(std::map<std::string, std::string>& a) {
    if (!a.count("PRG_graph_dir") || !a.count("outputDirectory") || !a.count("qualityMatrixFile")) {
        fprintf(stderr, "usage: hlala-b200 --action testPRGMapping --PRG_graph_dir <dir> --outputDirectory <dir> --qualityMatrixFile <file> [--haploidCoverage 5] [--seed 1]\n"
                        "       [--insertSizeMean 100 --insertSizeSD 10] [--device <n>] [--maxColumns <n>] [--maxThreads <n>]\n");
        return 2;
    }
    const std::string prg = a["PRG_graph_dir"], out = a["outputDirectory"], matrix = a["qualityMatrixFile"];
    const double coverage = a.count("haploidCoverage") ? atof(a["haploidCoverage"].c_str()) : 5.0;                 // S.simulateFromGraph(&g, 1, dir, 5, true), HLA-LA.cpp:1456
    const double is_mean = a.count("insertSizeMean") ? atof(a["insertSizeMean"].c_str()) : 100.0, is_sd = a.count("insertSizeSD") ? atof(a["insertSizeSD"].c_str()) : 10.0;   // simulator.h:20
    const unsigned seed = a.count("seed") ? (unsigned)strtoul(a["seed"].c_str(), nullptr, 10) : 1u;
    const int device = a.count("device") ? atoi(a["device"].c_str()) : 0; int maxcol = a.count("maxColumns") ? atoi(a["maxColumns"].c_str()) : 640;
    mkdir(out.c_str(), 0777);
    g_t0 = std::chrono::steady_clock::now();
    hlala_graph_t* g = nullptr;
    if (hlala_graph_load(prg.c_str(), &g)) return die("loading the PRG");
    if (hlala_graph_to_gpu(g, device)) return die("uploading the PRG");
    phase("PRG loaded and on the GPU");
    const int64_t n_sim = hlala_simulate_from_graph(g, (prg + "/PRG/graph.txt").c_str(), matrix.c_str(), 101, is_mean, is_sd, 1, out.c_str(), coverage, 1, seed);
    if (n_sim < 0) return die("simulating reads from the graph");
    phase("read pairs simulated from one random diploid genome of the graph");
    const std::string f1 = out + "/R_1.fq", f2 = out + "/R_2.fq", l1 = out + "/R_1.levels", l2 = out + "/R_2.levels";
    hlala_bam_batch_t* bam = nullptr;
    if (hlala_fastq_map_pairs(g, f1.c_str(), f2.c_str(), a.count("threads") ? atoi(a["threads"].c_str()) : 0, &bam)) return die("mapping the simulated reads");
    phase("simulated reads placed on the PRG contigs");
    hlala_seed_batch_t batch; const char* const* names = nullptr; int64_t counts[4];
    hlala_bam_batch_view(bam, &batch, &names); hlala_bam_batch_stats(bam, counts, nullptr, nullptr, nullptr);
    if (!a.count("maxColumns") && batch.n_reads > 0) maxcol = default_max_columns(batch);
    const int64_t nr = batch.n_reads, np = nr / 2; const size_t cells = (size_t)nr * (size_t)maxcol;
    if (np == 0) { fprintf(stderr, "hlala-b200: none of the %lld simulated pairs could be placed\n", (long long)n_sim); return 1; }
    std::vector<double> pair_mapq((size_t)np), read_mapq((size_t)nr), pair_ll((size_t)np); std::vector<uint8_t> rev((size_t)nr), gch(cells), sch(cells), fs(cells), mq(cells);
    std::vector<int32_t> slot((size_t)nr), ncols((size_t)nr), level(cells), edge(cells);
    hlala_pair_out_t po; po.max_columns = maxcol; po.pair_mapq = pair_mapq.data(); po.read_mapq = read_mapq.data(); po.read_reverse = rev.data(); po.chosen_slot = slot.data(); po.pair_ll = pair_ll.data();
    po.n_cols = ncols.data(); po.level = level.data(); po.edge = edge.data(); po.gchar = gch.data(); po.schar = sch.data(); po.from_seed = fs.data(); po.mapq = mq.data();
    if (hlala_align_pairs(g, &batch, is_mean, is_sd, &po, nullptr)) return die("aligning the pairs");
    phase("pairs aligned on the GPU");
    hlala_truth_t* truth = nullptr;
    if (hlala_truth_load(l1.c_str(), l2.c_str(), &truth)) return die("reading the true levels");
    if (hlala_truth_evaluate(truth, np, names, 0, &po, nullptr)) return die("comparing with the true levels");
    int64_t tot[3]; hlala_truth_totals(truth, tot);
    phase("alignments compared with the true levels");
    fprintf(stdout, "hlala-b200: %lld pairs simulated, %lld placed with both mates, %lld dropped\n", (long long)n_sim, (long long)np, (long long)counts[3]);
    fprintf(stdout, "\tGraph: %lld %g\n", (long long)tot[0], tot[0] ? (double)tot[1] / (double)tot[0] : 0.0);     // the line of HLA-LA.cpp:1258
    fprintf(stdout, "\tReads with less than 90 %% of their bases on the true level: %lld\n", (long long)tot[2]);
    hlala_truth_free(truth); hlala_bam_batch_free(bam); hlala_graph_free(g);
    return 0;
}
static int run_long_reads(std::map<std::string, std::string>& a);
static int run_multi_gpu(std::map<std::string, std::string>& a, int n_gpus) {
    const std::string out_dir = a["outputDirectory"], prg = a["PRG_graph_dir"]; int maxcol = a.count("maxColumns") ? atoi(a["maxColumns"].c_str()) : 640;
    int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < n_gpus) { fprintf(stderr, "hlala-b200: --gpus %d but %d CUDA devices are visible\n", n_gpus, ndev); return 1; }
    { const unsigned hc = std::max(1u, std::thread::hardware_concurrency()); setenv("HLALA_HOST_THREADS", std::to_string(std::max(1u, hc / (unsigned)n_gpus)).c_str(), 0); }
    std::vector<int> devs((size_t)n_gpus); for (int i = 0; i < n_gpus; i++) devs[(size_t)i] = i;
    std::vector<ncclComm_t> comms((size_t)n_gpus);
    if (ncclCommInitAll(comms.data(), n_gpus, devs.data()) != ncclSuccess) { fprintf(stderr, "hlala-b200: ncclCommInitAll failed\n"); return 1; }
    g_t0 = std::chrono::steady_clock::now();
    // rank 0's graph handle also serves the BAM ingest (contig names); every rank loads its own replica (the flat cache makes that seconds)
    std::vector<hlala_graph_t*> graphs((size_t)n_gpus, nullptr); std::atomic<int> failed{0}; std::vector<std::string> err((size_t)n_gpus);
    { std::vector<std::thread> th; for (int r = 0; r < n_gpus; r++) th.emplace_back([&, r]() { if (hlala_graph_load(prg.c_str(), &graphs[(size_t)r]) || hlala_graph_to_gpu(graphs[(size_t)r], r)) { err[(size_t)r] = hlala_last_error(); failed++; } }); for (auto& t : th) t.join(); }
    if (failed) { for (auto& e : err) if (!e.empty()) fprintf(stderr, "hlala-b200: loading the PRG: %s\n", e.c_str()); return 1; }
    phase("PRG loaded and replicated on the GPUs");
    hlala_bam_batch_t* bam = nullptr;
    if (read_seeds(a, graphs[0], &bam)) return 1;
    hlala_seed_batch_t batch; const char* const* names = nullptr; int64_t counts[4]; double is_mean = 0, is_sd = 0; int64_t is_n = 0;
    hlala_bam_batch_view(bam, &batch, &names); hlala_bam_batch_stats(bam, counts, &is_mean, &is_sd, &is_n);
    if (!a.count("maxColumns") && batch.n_reads > 0) maxcol = default_max_columns(batch);
    if (a.count("insertSizeMean") && a.count("insertSizeSD")) { is_mean = atof(a["insertSizeMean"].c_str()); is_sd = atof(a["insertSizeSD"].c_str()); }
    else { int64_t used = 0, skipped = 0; if (hlala_bam_insert_size(graphs[0], bam, maxcol, &is_mean, &is_sd, &used, &skipped)) return die("estimating the insert size"); }
    const int64_t n_pairs = batch.n_reads / 2;
    fprintf(stdout, "hlala-b200: %lld records, %lld used, %lld pairs to align on %d GPUs; insert size %g +- %g\n", (long long)counts[0], (long long)counts[1], (long long)n_pairs, n_gpus, is_mean, is_sd);
    if (n_pairs < n_gpus) { fprintf(stderr, "hlala-b200: fewer read pairs than GPUs\n"); return 1; }
    if (!(is_sd > 0)) { fprintf(stderr, "hlala-b200: cannot estimate the insert size; pass --insertSizeMean / --insertSizeSD\n"); return 1; }
    mkdir(out_dir.c_str(), 0777);
    const int64_t nl = hlala_graph_n_levels(graphs[0]);
    std::vector<const uint8_t*> blobs((size_t)n_gpus, nullptr); std::vector<int64_t> blob_bytes((size_t)n_gpus, 0), n_sel((size_t)n_gpus, 0);
    std::vector<hlala_typer_t*> typers((size_t)n_gpus, nullptr); Barrier bar; bar.n = n_gpus;
    const std::string hla_dir = out_dir + "/hla";
    struct stat sbuf; const std::string gdir = a.count("hla_nom_g_dir") ? a["hla_nom_g_dir"] : (stat((prg + "/hla_nom_g.txt").c_str(), &sbuf) == 0 ? prg : std::string("."));
    auto rank_main = [&](int r) {
        auto fail = [&](const std::string& what) { err[(size_t)r] = what + ": " + hlala_last_error(); failed++; };
        cudaSetDevice(r);
        Shard sh; { const int64_t base = n_pairs / n_gpus, rem = n_pairs % n_gpus; const int64_t p0 = r * base + std::min<int64_t>(r, rem); sh.cut(batch, p0, p0 + base + (r < rem ? 1 : 0)); }
        hlala_session_t* s = nullptr; int32_t* cov = nullptr; cudaStream_t st = nullptr; cudaStreamCreate(&st);
        bool ok = cudaMalloc((void**)&cov, (size_t)std::max<int64_t>(nl - 1, 1) * 4) == cudaSuccess && cudaMemsetAsync(cov, 0, (size_t)std::max<int64_t>(nl - 1, 1) * 4, st) == cudaSuccess;
        if (!ok) fail("allocating the coverage histogram");
        if (ok && (hlala_session_create(graphs[(size_t)r], &sh.view, maxcol, &s) || hlala_session_set_keep_columns(s, 1) || hlala_session_run(s, is_mean, is_sd, (uint64_t)(uintptr_t)cov, st))) { ok = false; fail("aligning"); }
        if (ok) { int64_t dig[4]; double sll = 0; if (hlala_session_digest(s, dig, &sll) || dig[3] != 0) { ok = false; fail("chains/pairs violated a reference invariant or a kernel capacity"); } }
        bar.wait();                                              // nobody enters a collective unless every rank got here in good shape
        if (failed) return;
        ncclAllReduce(cov, cov, (size_t)(nl - 1), ncclInt32, ncclSum, comms[(size_t)r], st); cudaStreamSynchronize(st);      // exchange 1: per-level coverage
        if (r == 0) {
            std::vector<int32_t> h((size_t)(nl - 1)); cudaMemcpy(h.data(), cov, h.size() * 4, cudaMemcpyDeviceToHost);
            const std::string path = out_dir + "/reads_per_level.txt"; FILE* f = fopen(path.c_str(), "w");
            if (f) { for (int64_t l = 0; l + 1 < nl; l++) fprintf(f, "%lld\t%s\t%d\n", (long long)l, hlala_graph_level_name(graphs[0], l), h[(size_t)l]); fclose(f); } else fail("cannot write " + path);
        }
        if (hlala_typer_create(prg.c_str(), &typers[(size_t)r]) || hlala_session_typing_extract(s, typers[(size_t)r], names + sh.pair0, sh.pair0, &blobs[(size_t)r], &blob_bytes[(size_t)r], &n_sel[(size_t)r])) fail("selecting the gene-overlapping pairs");
        bar.wait();                                              // exchange 2: every rank now sees every shard's blob (one address space)
        if (failed) return;
        RankCtx ctx{comms[(size_t)r], r};
        if (hlala_typer_infer(typers[(size_t)r], r, blobs.data(), blob_bytes.data(), n_gpus, is_mean, is_sd, r == 0 ? hla_dir.c_str() : nullptr, gdir.c_str(), r, n_gpus, nccl_sum_f64, &ctx, HLALA_TYPER_CALLBACK_ANY_THREAD)) fail("HLA type inference");
        bar.wait();
        hlala_session_free(s); cudaFree(cov); cudaStreamDestroy(st);
    };
    { std::vector<std::thread> th; for (int r = 0; r < n_gpus; r++) th.emplace_back(rank_main, r); for (auto& t : th) t.join(); }
    if (failed) { for (auto& e : err) if (!e.empty()) fprintf(stderr, "hlala-b200: %s\n", e.c_str()); return 1; }
    phase("read pairs aligned on all GPUs, coverage all-reduced, HLA types inferred (one NCCL all-reduce per locus), files written");
    long long sel = 0; for (int64_t v : n_sel) sel += v; fprintf(stdout, "hlala-b200: %lld read pairs overlap the typed genes\n", sel);
    for (int l = 0; l < hlala_typer_n_loci(typers[0]); l++) {
        const char* a1 = nullptr; const char* a2 = nullptr; double q1 = 0, q2 = 0;
        if (hlala_typer_result_call(typers[0], l, &a1, &a2, &q1, &q2) == 0) fprintf(stdout, "%s\t%s\t%s\t%g\t%g\n", hlala_typer_locus_name(typers[0], l), a1, a2, q1, q2);
    }
    for (int r = 0; r < n_gpus; r++) { hlala_typer_free(typers[(size_t)r]); ncclCommDestroy(comms[(size_t)r]); hlala_graph_free(graphs[(size_t)r]); }
    hlala_bam_batch_free(bam);
    return 0;
}

// --longReads ont2d|pacbio (HLA-LA.cpp:756-799 with a long-read mode): the remapped BAM of single long reads -> reads_per_level.txt + hla/*. Seam:
//   processBAM::alignReads_and_inferHLA(BAM, 0, 0, 0, outputDirectory, false, &HLAtyper, 1, longReads)  ->  alignReadsUnpaired_postSeedExtraction_andStoreInto
// (processBAM.cpp:1855, 2224-2337) + HLATypeInference(no paired reads, unpaired reads, ..., longReads) (:1920). Reads are aligned in chunks (the column arrays of
// thousands of columns per read are large); the gene-overlapping reads of every chunk are packed, and one typing call sees them all.
static int run_long_reads(std::map<std::string, std::string>& a) {
    const std::string out_dir = a["outputDirectory"], prg = a["PRG_graph_dir"]; const int device = a.count("device") ? atoi(a["device"].c_str()) : 0;
    g_t0 = std::chrono::steady_clock::now();
    hlala_graph_t* g = nullptr;
    if (hlala_graph_load(prg.c_str(), &g)) return die("loading the PRG");
    if (hlala_graph_to_gpu(g, device)) return die("uploading the PRG");
    phase("PRG loaded and on the GPU");
    hlala_bam_batch_t* bam = nullptr;
    if (hlala_bam_read_long(g, a["BAM"].c_str(), a.count("threads") ? atoi(a["threads"].c_str()) : 0, &bam)) return die("reading the BAM");
    hlala_seed_batch_t batch; const char* const* names = nullptr; int64_t counts[4]; double m0 = 0, s0 = 0; int64_t n0 = 0;
    hlala_bam_batch_view(bam, &batch, &names); hlala_bam_batch_stats(bam, counts, &m0, &s0, &n0);
    phase("BAM read, primary records selected and grouped");
    const int64_t n = batch.n_reads; int64_t longest = 0; for (int64_t r = 0; r < n; r++) longest = std::max<int64_t>(longest, batch.read_off[r + 1] - batch.read_off[r]);
    fprintf(stdout, "hlala-b200: %lld records, %lld used, %lld read names, %lld without a primary record, %lld long reads to align (longest %lld bases)\n",
            (long long)counts[0], (long long)counts[1], (long long)counts[2], (long long)counts[3], (long long)n, (long long)longest);
    if (n == 0) { fprintf(stderr, "hlala-b200: no long read with a primary record on the PRG contigs\n"); return 1; }
    const int maxcol = a.count("maxColumns") ? atoi(a["maxColumns"].c_str()) : (int)std::min<int64_t>(16384, longest + longest / 4 + 64);
    mkdir(out_dir.c_str(), 0777);
    const int64_t nl = hlala_graph_n_levels(g); std::vector<int32_t> cov((size_t)std::max<int64_t>(nl - 1, 1), 0);
    hlala_typer_t* t = nullptr; if (hlala_typer_create(prg.c_str(), &t)) return die("loading the typing tables");
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, ((int64_t)1 << 30) / ((int64_t)maxcol * 12)));      // about 1 GB of column arrays per chunk
    std::vector<std::vector<uint8_t>> blobs; int64_t n_sel = 0;
    std::vector<double> mapq((size_t)chunk), ll((size_t)chunk); std::vector<uint8_t> rev((size_t)chunk), gc((size_t)chunk * maxcol), sc((size_t)chunk * maxcol), mq((size_t)chunk * maxcol);
    std::vector<int32_t> ncol((size_t)chunk), lv((size_t)chunk * maxcol);
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t r1 = std::min(n, r0 + chunk), m = r1 - r0;
        std::vector<int64_t> ro((size_t)m + 1); std::vector<int32_t> co((size_t)m + 1), go;
        const int32_t c0 = batch.chain_off[r0], c1 = batch.chain_off[r1]; const int32_t g0 = batch.cigar_off[c0];
        for (int64_t r = r0; r <= r1; r++) { ro[(size_t)(r - r0)] = batch.read_off[r] - batch.read_off[r0]; co[(size_t)(r - r0)] = batch.chain_off[r] - c0; }
        go.resize((size_t)(c1 - c0) + 1); for (int32_t c = c0; c <= c1; c++) go[(size_t)(c - c0)] = batch.cigar_off[c] - g0;
        hlala_seed_batch_t v = batch; v.n_reads = m; v.read_off = ro.data(); v.bases = batch.bases + batch.read_off[r0]; v.quals = batch.quals + batch.read_off[r0]; v.chain_off = co.data();
        v.chain_contig = batch.chain_contig + c0; v.chain_pos = batch.chain_pos + c0; v.chain_flag = batch.chain_flag + c0; v.chain_as = batch.chain_as + c0; v.cigar_off = go.data(); v.cigar = batch.cigar + g0;
        hlala_pair_out_t o{}; o.max_columns = maxcol; o.pair_mapq = mapq.data(); o.pair_ll = ll.data(); o.read_reverse = rev.data(); o.n_cols = ncol.data(); o.level = lv.data(); o.gchar = gc.data(); o.schar = sc.data(); o.mapq = mq.data();
        if (hlala_align_long_reads(g, &v, &o, cov.data())) return die("aligning the long reads");
        const uint8_t* blob = nullptr; int64_t nb = 0, ns = 0;
        if (hlala_typing_blob_from_long_reads(t, &v, names + r0, &o, &blob, &nb, &ns)) return die("selecting the gene-overlapping reads");
        blobs.emplace_back(blob, blob + nb); n_sel += ns;
    }
    phase("long reads aligned (seed projection, padding, likelihoods, mapping qualities)");
    {   const std::string path = out_dir + "/reads_per_level.txt"; FILE* f = fopen(path.c_str(), "w");
        if (!f) { fprintf(stderr, "hlala-b200: cannot write %s\n", path.c_str()); return 1; }
        for (int64_t l = 0; l + 1 < nl; l++) fprintf(f, "%lld\t%s\t%d\n", (long long)l, hlala_graph_level_name(g, l), cov[(size_t)l]);
        fclose(f); }
    fprintf(stdout, "hlala-b200: %lld long reads overlap the typed genes\n", (long long)n_sel);
    if (n_sel == 0) { fprintf(stderr, "hlala-b200: no long read overlaps a typed gene\n"); return 1; }
    std::vector<const uint8_t*> bp; std::vector<int64_t> bb; for (auto& b : blobs) { bp.push_back(b.data()); bb.push_back((int64_t)b.size()); }
    const std::string hla_dir = out_dir + "/hla";
    struct stat sb; const std::string gdir = a.count("hla_nom_g_dir") ? a["hla_nom_g_dir"] : (stat((prg + "/hla_nom_g.txt").c_str(), &sb) == 0 ? prg : std::string("."));
    if (hlala_typer_infer(t, device, bp.data(), bb.data(), (int)bp.size(), 0, 1, hla_dir.c_str(), gdir.c_str(), 0, 1, nullptr, nullptr, 0)) return die("HLA type inference");
    phase("HLA types inferred, hla/* written");
    for (int l = 0; l < hlala_typer_n_loci(t); l++) {
        const char* a1 = nullptr; const char* a2 = nullptr; double q1 = 0, q2 = 0;
        if (hlala_typer_result_call(t, l, &a1, &a2, &q1, &q2) == 0) fprintf(stdout, "%s\t%s\t%s\t%g\t%g\n", hlala_typer_locus_name(t, l), a1, a2, q1, q2);
    }
    hlala_typer_free(t); hlala_bam_batch_free(bam); hlala_graph_free(g);
    return evaluate_calls(a, out_dir);
}

// --trueHLA <file> (HLA-LA.cpp:801-810): hla/R1_bestguess.txt against known types, the reference's summary on stdout
static int evaluate_calls(std::map<std::string, std::string>& a, const std::string& out_dir) {
    if (!a.count("trueHLA") || a["trueHLA"].empty()) return 0;
    std::vector<char> loci(4096), text(1 << 16); std::vector<int32_t> counts(2 * 64);
    const std::string sample = a.count("sampleID") ? a["sampleID"] : std::string("sample");
    const int n = hlala_evaluate_types(sample.c_str(), (out_dir + "/hla/R1_bestguess.txt").c_str(), a["trueHLA"].c_str(), loci.data(), (int64_t)loci.size(), counts.data(), 64, text.data(), (int64_t)text.size());
    if (n < 0) return die("evaluating the inferred types");
    fputs(text.data(), stdout);
    return 0;
}

int main(int argc, char** argv) {
    std::map<std::string, std::string> a;
    // the argument surface of the reference binary (HLA-LA.cpp:60-92 collects --key value pairs; HLA-LA.pl:563 passes the keys below) plus this program's own
    static const char* const known[] = {"action", "sampleID", "BAM", "outputDirectory", "PRG_graph_dir", "trueHLA", "maxThreads", "threads", "insertSizeMean", "insertSizeSD", "device", "gpus", "maxColumns",
                                        "hla_nom_g_dir", "bwa_bin", "samtools_bin", "FASTQ1", "FASTQ2", "FASTQU", "longReads", "mapAgainstCompleteGenome", "remap_with_a", "workingDir", "graph",
                                        "qualityMatrixFile", "haploidCoverage", "seed"};
    if ((argc - 1) % 2 != 0) { fprintf(stderr, "hlala-b200: arguments come as --key value pairs; '%s' has no value\n", argv[argc - 1]); return 2; }
    for (int i = 1; i + 1 < argc; i += 2) {
        if (strncmp(argv[i], "--", 2) != 0) { fprintf(stderr, "hlala-b200: bad argument %s\n", argv[i]); return 2; }
        bool ok = false; for (const char* k : known) ok = ok || strcmp(k, argv[i] + 2) == 0;
        if (!ok) { fprintf(stderr, "hlala-b200: unknown argument %s\n", argv[i]); return 2; }
        a[argv[i] + 2] = argv[i + 1];
    }
    if (a.count("maxThreads") && !a.count("threads")) a["threads"] = a["maxThreads"];     // HLA-LA.pl passes --maxThreads
    if (!a.count("action")) { fprintf(stderr, "\n\nMissing --action parameter. Please don't try calling me directly; use HLA-LA.pl instead (see documentation on GitHub).\n\n"); return 2; }     // HLA-LA.cpp:104-108
    if (a["action"] == "testBinary") { fprintf(stdout, "\nHLA*LA binary functional!\n\n"); return 0; }                                                                                          // HLA-LA.cpp:129-132
    const bool from_fastq = a["action"] == "HLA" && !a.count("BAM") && a.count("FASTQ1") && a.count("FASTQ2");
    if (a["action"] == "HLA" && !a.count("BAM") && !from_fastq && (a.count("FASTQ1") || a.count("FASTQ2") || a.count("FASTQU"))) {
        // HLA-LA.cpp:745-790: the reference maps the extracted FASTQ with `bwa mem -a -M` (BWAmapper::map / mapLong / mapUnpaired) and continues with the remapped BAM.
        // Paired files (--FASTQ1 + --FASTQ2, what HLA-LA.pl:563 passes for short reads) are mapped here (hlala_fastq_map_pairs); unpaired / long reads are not.
        fprintf(stderr, "hlala-b200: --FASTQU / a single --FASTQ1: map the reads first (bwa mem -a -M <PRG_graph_dir>/mapping_PRGonly/referenceGenome.fa ... | samtools sort) and pass the result as --BAM\n");
        return 2;
    }
    if (from_fastq && a.count("longReads") && a["longReads"] != "0" && a["longReads"] != "") { fprintf(stderr, "hlala-b200: --longReads takes a --BAM (bwa mem -x ont2d|pacbio)\n"); return 2; }
    if (a["action"] == "testPRGMapping") return run_test_prg_mapping(a);
    const bool long_reads = a.count("longReads") && a["longReads"] != "0" && a["longReads"] != "";
    if (long_reads && a["longReads"] != "ont2d" && a["longReads"] != "pacbio") { fprintf(stderr, "hlala-b200: --longReads must be 0, ont2d or pacbio\n"); return 2; }     // HLA-LA.cpp:759
    if (a["action"] == "prepareGraph" && a.count("PRG_graph_dir")) {
        // HLA-LA.cpp:1341-1385 reads graph.txt, computes the gap-edge paths and serialises the pointer graph (hours and ~40 GB for the real
        // PRG). Here the flat arrays (gap paths included) are built in seconds and cached next to graph.txt (PRG/graph.hlala_b200.cache); the
        // boost archives serializedGRAPH* are neither read nor written. No GPU is needed for this action.
        g_t0 = std::chrono::steady_clock::now();
        hlala_graph_t* g = nullptr;
        if (hlala_graph_load(a["PRG_graph_dir"].c_str(), &g)) return die("loading the PRG");
        phase("graph.txt read, gap-edge paths computed, flat arrays cached");
        fprintf(stdout, "hlala-b200: %lld levels, %lld nodes, %lld edges, %lld gap-edge paths, %lld contigs\n", (long long)hlala_graph_n_levels(g), (long long)hlala_graph_n_nodes(g),
                (long long)hlala_graph_n_edges(g), (long long)hlala_graph_n_paths(g), (long long)hlala_graph_n_contigs(g));
        hlala_graph_free(g);
        return 0;
    }
    if (a["action"] != "HLA" || !(a.count("BAM") || from_fastq) || !a.count("outputDirectory") || !a.count("PRG_graph_dir")) {
        fprintf(stderr, "usage: hlala-b200 --action HLA --sampleID <id> --BAM <remapped.bam> --outputDirectory <dir> --PRG_graph_dir <dir>\n"
                        "       hlala-b200 --action HLA --sampleID <id> --FASTQ1 <R1.fq[.gz]> --FASTQ2 <R2.fq[.gz]> --outputDirectory <dir> --PRG_graph_dir <dir>   (reads placed on the PRG contigs here, no bwa)\n"
                        "       hlala-b200 --action prepareGraph --PRG_graph_dir <dir>\n"
                        "       hlala-b200 --action testBinary\n"
                        "       hlala-b200 --action testPRGMapping --PRG_graph_dir <dir> --outputDirectory <dir> --qualityMatrixFile <file>   (simulated reads -> placed -> aligned -> bases on their true level)\n"
                        "       [--longReads ont2d|pacbio]  (the BAM then holds single long reads: bwa mem -x ont2d|pacbio)\n"
                        "       [--insertSizeMean <m> --insertSizeSD <s>] [--device <n> | --gpus <N>] [--maxColumns <n>] [--maxThreads <n>] [--trueHLA <file>]\n");
        return 2;
    }
    if (long_reads) return run_long_reads(a);
    if (a.count("gpus") && atoi(a["gpus"].c_str()) > 1) return run_multi_gpu(a, atoi(a["gpus"].c_str()));
    const std::string out_dir = a["outputDirectory"], prg = a["PRG_graph_dir"];
    const int device = a.count("device") ? atoi(a["device"].c_str()) : 0; int maxcol = a.count("maxColumns") ? atoi(a["maxColumns"].c_str()) : 640;
    g_t0 = std::chrono::steady_clock::now();
    hlala_graph_t* g = nullptr;
    if (hlala_graph_load(prg.c_str(), &g)) return die("loading the PRG");
    phase("PRG loaded (graph.txt, gap paths, contigs, translations)");
    if (hlala_graph_to_gpu(g, device)) return die("uploading the PRG");
    phase("PRG on the GPU");
    hlala_bam_batch_t* bam = nullptr;
    if (read_seeds(a, g, &bam)) return 1;
    hlala_seed_batch_t batch; const char* const* names = nullptr; int64_t counts[4]; double is_mean = 0, is_sd = 0; int64_t is_n = 0;
    hlala_bam_batch_view(bam, &batch, &names); hlala_bam_batch_stats(bam, counts, &is_mean, &is_sd, &is_n);
    if (!a.count("maxColumns") && batch.n_reads > 0) maxcol = default_max_columns(batch);
    if (a.count("insertSizeMean") && a.count("insertSizeSD")) { is_mean = atof(a["insertSizeMean"].c_str()); is_sd = atof(a["insertSizeSD"].c_str()); }
    else {   // what HLA-LA.cpp:788 does: processBAM::estimateInsertSize on the first ~4000 read names
        int64_t used = 0, skipped = 0;
        if (hlala_bam_insert_size(g, bam, maxcol, &is_mean, &is_sd, &used, &skipped)) return die("estimating the insert size");
        is_n = used - skipped; phase("insert size estimated (primary records of the first 4000 read names aligned)");
    }
    fprintf(stdout, "hlala-b200: %lld records, %lld used, %lld read names, %lld incomplete pairs, %lld pairs to align; insert size %g +- %g (from %lld pairs)\n",
            (long long)counts[0], (long long)counts[1], (long long)counts[2], (long long)counts[3], (long long)(batch.n_reads / 2), is_mean, is_sd, (long long)is_n);
    if (batch.n_reads == 0) { fprintf(stderr, "hlala-b200: no complete read pair on the PRG contigs\n"); return 1; }
    if (!(is_sd > 0)) { fprintf(stderr, "hlala-b200: cannot estimate the insert size; pass --insertSizeMean / --insertSizeSD\n"); return 1; }
    mkdir(out_dir.c_str(), 0777);
    hlala_session_t* s = nullptr;
    if (hlala_session_create(g, &batch, maxcol, &s)) return die("creating the alignment session");
    hlala_session_set_keep_columns(s, 1); hlala_session_set_coverage(s, 1);
    if (hlala_session_run(s, is_mean, is_sd, 0, nullptr)) return die("aligning");
    phase("read pairs aligned (batch upload + kernels)");
    int64_t dig[4]; double sll = 0;
    if (hlala_session_digest(s, dig, &sll)) return die("reading the alignment digest");
    if (dig[3] != 0) { fprintf(stderr, "hlala-b200: %lld chains/pairs violated a reference invariant or a kernel capacity\n", (long long)dig[3]); return 1; }
    {   // reads_per_level.txt
        const int64_t nl = hlala_graph_n_levels(g); std::vector<int32_t> cov((size_t)(nl > 1 ? nl - 1 : 1));
        if (hlala_session_fetch_coverage(s, cov.data())) return die("reading the coverage");
        const std::string path = out_dir + "/reads_per_level.txt"; FILE* f = fopen(path.c_str(), "w");
        if (!f) { fprintf(stderr, "hlala-b200: cannot write %s\n", path.c_str()); return 1; }
        for (int64_t l = 0; l + 1 < nl; l++) fprintf(f, "%lld\t%s\t%d\n", (long long)l, hlala_graph_level_name(g, l), cov[(size_t)l]);
        fclose(f);
    }
    phase("reads_per_level.txt written");
    hlala_typer_t* t = nullptr;
    if (hlala_typer_create(prg.c_str(), &t)) return die("loading the typing tables");
    const uint8_t* blob = nullptr; int64_t blob_bytes = 0, n_sel = 0;
    if (hlala_session_typing_extract(s, t, names, 0, &blob, &blob_bytes, &n_sel)) return die("selecting the gene-overlapping pairs");
    fprintf(stdout, "hlala-b200: %lld read pairs overlap the typed genes\n", (long long)n_sel);
    const std::string hla_dir = out_dir + "/hla";
    struct stat sb; const std::string gdir = a.count("hla_nom_g_dir") ? a["hla_nom_g_dir"] : (stat((prg + "/hla_nom_g.txt").c_str(), &sb) == 0 ? prg : std::string("."));   // the reference opens hla_nom_g.txt in the current directory (HLATyper.cpp:4157)
    if (hlala_typer_infer(t, device, &blob, &blob_bytes, 1, is_mean, is_sd, hla_dir.c_str(), gdir.c_str(), 0, 1, nullptr, nullptr, 0)) return die("HLA type inference");
    phase("HLA types inferred, hla/* written");
    for (int l = 0; l < hlala_typer_n_loci(t); l++) {
        const char* a1 = nullptr; const char* a2 = nullptr; double q1 = 0, q2 = 0;
        if (hlala_typer_result_call(t, l, &a1, &a2, &q1, &q2) == 0) fprintf(stdout, "%s\t%s\t%s\t%g\t%g\n", hlala_typer_locus_name(t, l), a1, a2, q1, q2);
    }
    hlala_typer_free(t); hlala_session_free(s); hlala_bam_batch_free(bam); hlala_graph_free(g);
    return evaluate_calls(a, out_dir);
}
