// Test-only: drives the product's typing HOST logic (hla-la_b200/host/hla_typing.cpp: tables, projection, filters, calls, writers) with a
// plain-loop stand-in for the two GPU kernels, so that the CPU test-suite can compare every file the host side writes with the oracle /
// the compiled reference without a GPU. The stand-in lives here, in tests/, and is never linked into libhlala_b200.so.
#include "../../hla-la_b200/host/hla_typing.h"
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>

using namespace hlala;

namespace {
typedef int (*host_allreduce_fn)(double* data, long long count);   // sum over ranks, in place (the gloo test passes torch.distributed.all_reduce)
struct LoopDevice : TypingDevice {   // HLATyper.cpp:2049-2364 restated as loops over the device input layout
    int rank = 0, world = 1; host_allreduce_fn allreduce = nullptr;   // reads split across ranks like the GPU device does (c_api.cu: r0 = R*rank/world)
    void abort_locus(int32_t C) override {     // poisoned contribution, as the GPU device does (c_api.cu)
        if (world <= 1 || !allreduce) return;
        std::vector<double> v((size_t)3 * ((size_t)C * ((size_t)C + 1) / 2), NAN); allreduce(v.data(), (long long)v.size());
    }
    void run_locus(const LocusDeviceInput& in, bool, LocusDeviceOutput& out) override {
        const int C = in.C, R = in.R; const int r0 = (int)((long long)R * rank / world), r1 = (int)((long long)R * (rank + 1) / world); const double ip = in.long_reads ? 0.075 : 0.001, dp = in.long_reads ? 0.075 : 0.001; const double ll_ins_actual = log(ip) + log(1.0 / 4.0), ll_del = log(dp), ll_mm = log(1 - ip - dp);
        out.LL.assign((size_t)C * R, 0); out.mism.assign((size_t)C * R, 0);
        for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) { double ll = 0; int mm = 0; const std::string& cs = (*in.cluster_seq)[c];
            for (int k = in.rec_off[r]; k < in.rec_off[r + 1]; k++) { char e = cs[in.rec_pos[k]]; char c0 = (char)in.rec_c0[k]; unsigned glen = in.rec_glen[k]; double lp = 0;
                if (e == '_') { if (c0 != '_') lp += ll_ins_actual * (1 + (glen - 1)); }
                else { if (c0 == '_') lp += ll_del; else { lp += ll_mm; double pc = 1 - exp(log(10) * ((double)((int)in.rec_q0[k] - 33) / (double)-10)); if (pc > 0.999) pc = 0.999; if (pc == 0) pc = 0.001; if (e == c0) lp += log(pc); else lp += log((1 - pc) * (1.0 / 3.0)); }
                       lp += ll_ins_actual * (glen - 1); }
                if (c0 != '_' && !(glen == 1 && c0 == e)) mm++;
                ll += lp; }
            out.LL[(size_t)c * R + r] = ll; out.mism[(size_t)c * R + r] = mm; }
        auto log_avg = [](double a, double b) { return a > b ? log(0.5) + (log(1 + exp(b - a)) + a) : log(0.5) + (log(1 + exp(a - b)) + b); };
        for (int c1 = 0; c1 < C; c1++) for (int c2 = c1; c2 < C; c2++) { double pl = 0, sa = 0, sm = 0;
            for (int r = r0; r < r1; r++) { int m1 = out.mism[(size_t)c1 * R + r], m2 = out.mism[(size_t)c2 * R + r]; pl += log_avg(out.LL[(size_t)c1 * R + r], out.LL[(size_t)c2 * R + r]); sa += (double)(m1 + m2) / 2.0; sm += m1 < m2 ? m1 : m2; }
            out.pair_ll.push_back(pl); out.pair_mavg.push_back(sa); out.pair_mmin.push_back(sm); }
        if (world > 1) {   // ONE all-reduce per locus over [pair_ll | pair_mavg | pair_mmin], as hlala_typer_infer does on the device
            if (!allreduce) throw std::runtime_error("world > 1 needs an all-reduce callback");
            std::vector<double> v; v.insert(v.end(), out.pair_ll.begin(), out.pair_ll.end()); v.insert(v.end(), out.pair_mavg.begin(), out.pair_mavg.end()); v.insert(v.end(), out.pair_mmin.begin(), out.pair_mmin.end());
            if (allreduce(v.data(), (long long)v.size()) != 0) throw std::runtime_error("all-reduce callback failed");
            if (!v.empty() && v[0] != v[0]) throw std::runtime_error("this locus failed on another rank (poisoned all-reduce)");
            const size_t n = out.pair_ll.size(); for (size_t i = 0; i < n; i++) { out.pair_ll[i] = v[i]; out.pair_mavg[i] = v[n + i]; out.pair_mmin[i] = v[2 * n + i]; }
        }
    }
};
std::string g_err;
}

extern "C" {
const char* typing_host_last_error() { return g_err.c_str(); }
// alignment arrays as every implementation's `pairs` exports them ([n_reads, cap]); pairs named "r<p>"
int typing_host_run_ranked(const char* prg_dir, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int cap, const int32_t* n_cols, const int32_t* level,
                           const uint8_t* g, const uint8_t* s, const uint8_t* mq, const uint8_t* reverse, const double* read_mapq, double is_mean, double is_sd, const char* out_dir, int roundtrip_blob,
                           int rank, int world, host_allreduce_fn allreduce, double* call_q /* [n_loci * 2] */, double* pair_ll_sum /* [n_loci] */);
// long-read mode: alignment arrays [n_reads, cap] of single reads (any implementation's long_reads output), reads named "r<index>"
int typing_host_run_long(const char* prg_dir, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int cap, const int32_t* n_cols, const int32_t* level,
                         const uint8_t* g, const uint8_t* s, const uint8_t* mq, const uint8_t* reverse, const double* read_mapq, const char* out_dir, int roundtrip_blob) {
    try {
        TypingTables T; T.load(prg_dir);
        TypingReads tr = long_read_typing_input(T, n_reads, nullptr, read_off, bases, quals, cap, n_cols, level, g, s, mq, reverse, read_mapq), use;
        if (roundtrip_blob) { std::vector<uint8_t> b = tr.serialize(); use.deserialize_append(b.data(), b.size()); } else use = tr;
        LoopDevice dev; TypingOptions opt; std::vector<LocusCall> calls;
        run_typing(T, use, 0, 1, out_dir, prg_dir, dev, opt, calls);
        return (int)calls.size();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int typing_host_run(const char* prg_dir, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int cap, const int32_t* n_cols, const int32_t* level,
                    const uint8_t* g, const uint8_t* s, const uint8_t* mq, const uint8_t* reverse, const double* read_mapq, double is_mean, double is_sd, const char* out_dir, int roundtrip_blob) {
    return typing_host_run_ranked(prg_dir, n_reads, read_off, bases, quals, cap, n_cols, level, g, s, mq, reverse, read_mapq, is_mean, is_sd, out_dir, roundtrip_blob, 0, 1, nullptr, nullptr, nullptr);
}
int typing_host_run_ranked(const char* prg_dir, long long n_reads, const int64_t* read_off, const uint8_t* bases, const uint8_t* quals, int cap, const int32_t* n_cols, const int32_t* level,
                    const uint8_t* g, const uint8_t* s, const uint8_t* mq, const uint8_t* reverse, const double* read_mapq, double is_mean, double is_sd, const char* out_dir, int roundtrip_blob,
                           int rank, int world, host_allreduce_fn allreduce, double* call_q, double* pair_ll_sum) {
    try {
        TypingTables T; T.load(prg_dir);
        TypingReads tr; tr.col_off.push_back(0); tr.base_off.push_back(0);
        for (long long p = 0; p + 1 < n_reads; p += 2) {
            int f[2] = {-1, -1}, l[2] = {-1, -1};
            for (int m = 0; m < 2; m++) for (int c = 0; c < n_cols[p + m]; c++) { int lv = level[(size_t)(p + m) * cap + c]; if (lv != -1) { if (f[m] == -1) f[m] = lv; l[m] = lv; } }
            if (!pair_overlaps_genes(T, f[0], l[0], f[1], l[1])) continue;
            tr.pair_id.push_back(p / 2); tr.name.push_back("r" + std::to_string(p / 2));
            for (int m = 0; m < 2; m++) { long long r = p + m; size_t o = (size_t)r * cap; int n = n_cols[r];
                tr.level.insert(tr.level.end(), level + o, level + o + n); tr.g.insert(tr.g.end(), g + o, g + o + n); tr.s.insert(tr.s.end(), s + o, s + o + n); tr.mq.insert(tr.mq.end(), mq + o, mq + o + n);
                tr.col_off.push_back((int64_t)tr.level.size());
                tr.bases.insert(tr.bases.end(), bases + read_off[r], bases + read_off[r + 1]); tr.quals.insert(tr.quals.end(), quals + read_off[r], quals + read_off[r + 1]); tr.base_off.push_back((int64_t)tr.bases.size());
                tr.reverse.push_back(reverse[r]); tr.mapq.push_back(read_mapq[r]); }
        }
        TypingReads use;
        if (roundtrip_blob) {   // through the wire format, split in two blobs like a 2-rank gather
            TypingReads a, b; a.col_off.push_back(0); a.base_off.push_back(0); b.col_off.push_back(0); b.base_off.push_back(0);
            size_t half = tr.n_pairs() / 2;
            for (size_t i = 0; i < tr.n_pairs(); i++) { TypingReads& d = i < half ? a : b; d.pair_id.push_back(tr.pair_id[i]); d.name.push_back(tr.name[i]);
                for (int m = 0; m < 2; m++) { size_t r = 2 * i + m; d.level.insert(d.level.end(), tr.level.begin() + tr.col_off[r], tr.level.begin() + tr.col_off[r + 1]); d.g.insert(d.g.end(), tr.g.begin() + tr.col_off[r], tr.g.begin() + tr.col_off[r + 1]);
                    d.s.insert(d.s.end(), tr.s.begin() + tr.col_off[r], tr.s.begin() + tr.col_off[r + 1]); d.mq.insert(d.mq.end(), tr.mq.begin() + tr.col_off[r], tr.mq.begin() + tr.col_off[r + 1]); d.col_off.push_back((int64_t)d.level.size());
                    d.bases.insert(d.bases.end(), tr.bases.begin() + tr.base_off[r], tr.bases.begin() + tr.base_off[r + 1]); d.quals.insert(d.quals.end(), tr.quals.begin() + tr.base_off[r], tr.quals.begin() + tr.base_off[r + 1]); d.base_off.push_back((int64_t)d.bases.size());
                    d.reverse.push_back(tr.reverse[r]); d.mapq.push_back(tr.mapq[r]); } }
            std::vector<uint8_t> ba = a.serialize(), bb = b.serialize();
            use.deserialize_append(ba.data(), ba.size()); use.deserialize_append(bb.data(), bb.size());
        } else use = tr;
        LoopDevice dev; dev.rank = rank; dev.world = world; dev.allreduce = allreduce; TypingOptions opt; if (allreduce) opt.threads = 1; std::vector<LocusCall> calls;
        run_typing(T, use, is_mean, is_sd, out_dir, prg_dir, dev, opt, calls);
        for (size_t l = 0; l < calls.size(); l++) { if (call_q) { call_q[2 * l] = calls[l].q1; call_q[2 * l + 1] = calls[l].q2; } if (pair_ll_sum) { double t = 0; for (double v : calls[l].dev.pair_ll) t += v; pair_ll_sum[l] = t; } }
        return (int)calls.size();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
}
