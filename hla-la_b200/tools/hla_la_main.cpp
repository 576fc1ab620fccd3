// hlala-b200: command-line front end for the accelerated path, with the argument names of the reference binary's `--action HLA`
// (HLA-LA.cpp:196-210: --action --sampleID --BAM --outputDirectory --PRG_graph_dir). It is the seam B0 of SURVEY.md §8b:
//   void processBAM::alignReads_and_inferHLA(BAM, 0, IS_mean, IS_sd, outputDirectory, false, &HLAtyper, 1, "")   mapper/processBAM.h:132
// called at HLA-LA.cpp:799 — i.e. it takes the REMAPPED BAM (bwa mem -a -M against PRG_graph_dir/mapping_PRGonly/referenceGenome.fa; the
// read extraction and the bwa/samtools steps before it belong to the Perl/C++ control plane that is out of scope) and writes
//   <outputDirectory>/reads_per_level.txt         processBAM.cpp:1907-1913
//   <outputDirectory>/hla/*                        hla::HLATyper::HLATypeInference, hla/HLATyper.cpp:933-2810
// Everything goes through the C ABI (include/hlala_b200.h); there is no CPU fallback.
#include "../../include/hlala_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <sys/stat.h>
#include <vector>

static std::chrono::steady_clock::time_point g_t0;
static void phase(const char* what) {   // wall time of every phase, on stdout
    const auto t1 = std::chrono::steady_clock::now(); fprintf(stdout, "hlala-b200: [%8.2f s] %s\n", std::chrono::duration<double>(t1 - g_t0).count(), what); g_t0 = t1;
}
static int die(const char* what) { fprintf(stderr, "hlala-b200: %s: %s\n", what, hlala_last_error()); return 1; }

int main(int argc, char** argv) {
    std::map<std::string, std::string> a;
    for (int i = 1; i + 1 < argc; i += 2) { if (strncmp(argv[i], "--", 2) != 0) { fprintf(stderr, "hlala-b200: bad argument %s\n", argv[i]); return 2; } a[argv[i] + 2] = argv[i + 1]; }
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
    if (a["action"] != "HLA" || !a.count("BAM") || !a.count("outputDirectory") || !a.count("PRG_graph_dir")) {
        fprintf(stderr, "usage: hlala-b200 --action HLA --sampleID <id> --BAM <remapped.bam> --outputDirectory <dir> --PRG_graph_dir <dir>\n"
                        "       hlala-b200 --action prepareGraph --PRG_graph_dir <dir>\n"
                        "       [--insertSizeMean <m> --insertSizeSD <s>] [--device <n>] [--maxColumns <n>] [--threads <n>]\n");
        return 2;
    }
    const std::string out_dir = a["outputDirectory"], prg = a["PRG_graph_dir"];
    const int device = a.count("device") ? atoi(a["device"].c_str()) : 0; const int maxcol = a.count("maxColumns") ? atoi(a["maxColumns"].c_str()) : 640;
    g_t0 = std::chrono::steady_clock::now();
    hlala_graph_t* g = nullptr;
    if (hlala_graph_load(prg.c_str(), &g)) return die("loading the PRG");
    phase("PRG loaded (graph.txt, gap paths, contigs, translations)");
    if (hlala_graph_to_gpu(g, device)) return die("uploading the PRG");
    phase("PRG on the GPU");
    hlala_bam_batch_t* bam = nullptr;
    if (hlala_bam_read(g, a["BAM"].c_str(), a.count("threads") ? atoi(a["threads"].c_str()) : 0, &bam)) return die("reading the BAM");
    phase("BAM read, records selected and grouped");
    hlala_seed_batch_t batch; const char* const* names = nullptr; int64_t counts[4]; double is_mean = 0, is_sd = 0; int64_t is_n = 0;
    hlala_bam_batch_view(bam, &batch, &names); hlala_bam_batch_stats(bam, counts, &is_mean, &is_sd, &is_n);
    if (a.count("insertSizeMean")) is_mean = atof(a["insertSizeMean"].c_str());
    if (a.count("insertSizeSD")) is_sd = atof(a["insertSizeSD"].c_str());
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
    return 0;
}
