// See contig_mapper.h.
#include "contig_mapper.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <zlib.h>

namespace hlala {

namespace {

inline int code_of(uint8_t c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; } return 4; }
inline char comp_of(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; } return 'N'; }
std::string revcomp(const std::string& s) { std::string r(s.size(), 'N'); for (size_t i = 0; i < s.size(); i++) r[i] = comp_of(s[s.size() - 1 - i]); return r; }

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_S = 4 };
const int NEG = -(1 << 28);

struct Vote { int32_t contig; int32_t diag; };

} // namespace

ContigMapper::ContigMapper(const FlatGraph& g, const MapperParams& p) : g_(g), p_(p) {
    if (p_.k < 11 || p_.k > 31 || p_.ref_step < 1 || p_.read_step < 1 || p_.band < 4) throw std::runtime_error("ContigMapper: k in 11..31, steps >= 1, band >= 4");
    const uint64_t mask = (p_.k == 32) ? ~0ull : ((1ull << (2 * p_.k)) - 1);
    struct Entry { uint64_t key; uint32_t pos, ctg; bool operator<(const Entry& o) const { return key != o.key ? key < o.key : pos < o.pos; } };
    std::vector<Entry> e;
    if (g.contig_seq.size() >= (1ull << 32)) throw std::runtime_error("ContigMapper: contigs beyond 4 G bases");
    for (int32_t c = 0; c < g.n_contigs; c++) {
        const int64_t b = g.contig_off[(size_t)c], n = g.contig_off[(size_t)c + 1] - b;
        uint64_t key = 0; int run = 0;
        for (int64_t i = 0; i < n; i++) {
            const int cd = code_of(g.contig_seq[(size_t)(b + i)]);
            if (cd > 3) { run = 0; key = 0; continue; }
            key = ((key << 2) | (uint64_t)cd) & mask; run++;
            const int64_t start = i - p_.k + 1;
            if (run >= p_.k && start % p_.ref_step == 0) e.push_back(Entry{key, (uint32_t)(b + start), (uint32_t)c});
        }
    }
    std::sort(e.begin(), e.end());
    keys_.resize(e.size()); pos_.resize(e.size()); ctg_.resize(e.size());
    for (size_t i = 0; i < e.size(); i++) { keys_[i] = e[i].key; pos_[i] = e[i].pos; ctg_[i] = e[i].ctg; }
    const int bits = std::min(22, 2 * p_.k); bucket_shift_ = 2 * p_.k - bits;
    bucket_.assign(((size_t)1 << bits) + 1, (uint32_t)e.size());
    { size_t i = 0; for (size_t v = 0; v < ((size_t)1 << bits); v++) { while (i < e.size() && (keys_[i] >> bucket_shift_) < v) i++; bucket_[v] = (uint32_t)i; } }
}

// Banded affine-gap alignment of read[0..len) around `diag` (contig position of read base 0) on `contig`: any start and end in the read, a clipped end costs p_.clip.
bool ContigMapper::align(const uint8_t* rd, int len, int32_t contig, int64_t diag, Placement& out, int force_band) const {
    const int64_t cb = g_.contig_off[(size_t)contig], clen = g_.contig_off[(size_t)contig + 1] - cb;
    const uint8_t* ref = g_.contig_seq.data() + cb;
    const int go = p_.gap_open + p_.gap_extend, ge = p_.gap_extend;
    // Most placements have no gap at all: the best stretch of the diagonal itself (one pass). It is taken as it stands when it spans the whole read with at most six
    // scattered mismatches (SNPs against another haplotype, sequencing errors): a gap (7 or more) cannot win that back. Everything else goes through the banded alignment.
    if (diag >= 0 && diag + len <= clen) {
        int run = NEG, run_start = 0, run_mm = 0, ub = NEG, ub_s = 0, ub_e = 0, ub_mm = 0;
        for (int i = 0; i < len; i++) {
            const int rc = code_of(ref[diag + i]); const bool eq = rd[i] < 4 && rd[i] == rc; const int sc = eq ? p_.match : -p_.mismatch;
            const int fresh = (i == 0 ? 0 : -p_.clip);
            if (eq && fresh >= run) { run = fresh; run_start = i; run_mm = 0; }   // a stretch begins on a matching base when that is at least as good as going on
            if (run > NEG / 2) {
                run += sc; run_mm += !eq;
                const int fin = run + (i == len - 1 ? 0 : -p_.clip);
                if (eq && (fin > ub || (fin == ub && i > ub_e))) { ub = fin; ub_s = run_start; ub_e = i; ub_mm = run_mm; }
            }
        }
        if (ub > NEG / 2 && ub_s == 0 && ub_e == len - 1 && ub_mm <= 6) {   // only when it spans the whole read: a clipped stretch may be one side of an insertion or deletion
            const int raw = ub + (ub_s > 0 ? p_.clip : 0) + (ub_e < len - 1 ? p_.clip : 0);
            if (raw < p_.min_score) return false;
            out.contig = contig; out.pos = (int32_t)(diag + ub_s); out.score = raw; out.cigar.clear();
            if (ub_s > 0) out.cigar.push_back(((uint32_t)ub_s << 4) | OP_S);
            out.cigar.push_back(((uint32_t)(ub_e - ub_s + 1) << 4) | OP_M);
            if (ub_e < len - 1) out.cigar.push_back(((uint32_t)(len - 1 - ub_e) << 4) | OP_S);
            return true;
        }
    }
    // cell (i, b): read base i against contig position j = diag + i + b - B. A narrow band first; the full one only if the narrow alignment leans on the band's edge.
    static thread_local std::vector<int32_t> M, E, F; static thread_local std::vector<uint8_t> tb, rcode;
    // contig codes of the widest window once (9 = beyond the contig: such cells stay unreachable)
    const int BW = std::max(p_.band, force_band); rcode.resize((size_t)len + 2 * (size_t)BW);
    for (int64_t q = 0; q < (int64_t)rcode.size(); q++) { const int64_t j = diag - BW + q; rcode[(size_t)q] = (j >= 0 && j < clen) ? (uint8_t)code_of(ref[j]) : (uint8_t)9; }
  for (int attempt = force_band > 0 ? 1 : 0; attempt < 2; attempt++) {
    const int B = force_band > 0 ? force_band : attempt == 0 ? std::max(4, p_.band / 3) : p_.band, Wd = 2 * B + 1, Ws = Wd + 2; bool edge = false;   // rows carry one unreachable cell at either end
    M.assign((size_t)2 * Ws, NEG); E.assign((size_t)2 * Ws, NEG); F.assign((size_t)2 * Ws, NEG);
    tb.assign((size_t)len * Wd, 0);   // bits 0-1: M came from 0 start, 1 M, 2 E, 3 F; bit 2: E extended; bit 3: F extended
    int best = NEG, best_i = -1, best_b = -1;
    const int mt = p_.match, mm = -p_.mismatch, clipc = p_.clip;
    for (int i = 0; i < len; i++) {
        int32_t* Mc = &M[(size_t)(i & 1) * Ws] + 1; int32_t* Ec = &E[(size_t)(i & 1) * Ws] + 1; int32_t* Fc = &F[(size_t)(i & 1) * Ws] + 1;
        const int32_t* Mp = &M[(size_t)((i & 1) ^ 1) * Ws] + 1; const int32_t* Ep = &E[(size_t)((i & 1) ^ 1) * Ws] + 1; const int32_t* Fp = &F[(size_t)((i & 1) ^ 1) * Ws] + 1;
        const int start = i == 0 ? 0 : -clipc, endc = i == len - 1 ? 0 : -clipc;
        const uint8_t* rc = rcode.data() + (BW - B) + i;   // rc[b] = code of contig position diag + i + b - B
        const uint8_t ri = rd[i]; uint8_t* trow = tb.data() + (size_t)i * Wd;
        for (int b = 0; b < Wd; b++) {
            const uint8_t c = rc[b];
            if (c == 9) { Mc[b] = NEG; Ec[b] = NEG; Fc[b] = NEG; continue; }
            int from = start, which = 0;
            if (Mp[b] > from) { from = Mp[b]; which = 1; }      // the row before read base 0 is all unreachable: no special case for i == 0
            if (Ep[b] > from) { from = Ep[b]; which = 2; }
            if (Fp[b] > from) { from = Fp[b]; which = 3; }
            const int m = from + ((ri < 4 && ri == c) ? mt : mm);
            uint8_t t = (uint8_t)which;
            const int eo = Mc[b - 1] - go, ex = Ec[b - 1] - ge; int e = eo; if (ex > eo) { e = ex; t |= 4; }       // deletion: contig position consumed after read base i
            const int fo = Mp[b + 1] - go, fx = Fp[b + 1] - ge; int f = fo; if (fx > fo) { f = fx; t |= 8; }       // insertion: read base i without a contig position
            const int fin = m + endc;
            if (fin > best || (fin == best && i > best_i)) { best = fin; best_i = i; best_b = b; }
            Mc[b] = m; Ec[b] = e < NEG ? NEG : e; Fc[b] = f < NEG ? NEG : f; trow[b] = t;
        }
    }
    if (best_i < 0) { if (attempt == 0) continue; return false; }
    // trace back
    std::vector<uint8_t> ops; int i = best_i, b = best_b, state = 0;   // state 0 M, 1 E, 2 F
    int start_i = -1; int64_t start_j = -1;
    for (;;) {
        const uint8_t t = tb[(size_t)i * Wd + b];
        if (state == 0) {
            ops.push_back(OP_M); const int w = t & 3;
            if (w == 0) { start_i = i; start_j = diag + i + b - B; break; }
            i--; state = w - 1;   // diagonal predecessor keeps b
        } else if (state == 1) { ops.push_back(OP_D); const bool ext = t & 4; b--; state = ext ? 1 : 0; }
        else { ops.push_back(OP_I); const bool ext = t & 8; i--; b++; state = ext ? 2 : 0; }
        if (b <= 0 || b >= Wd - 1) edge = true;
        if (i < 0 || b < 0 || b >= Wd) { edge = true; start_i = -1; break; }
    }
    if (attempt == 0 && (edge || best_b <= 0 || best_b >= Wd - 1)) continue;
    if (start_i < 0) return false;
    std::reverse(ops.begin(), ops.end());
    const int raw = best + (start_i > 0 ? p_.clip : 0) + (best_i < len - 1 ? p_.clip : 0);
    if (raw < p_.min_score) return false;
    out.contig = contig; out.pos = (int32_t)start_j; out.score = raw; out.cigar.clear();
    if (start_i > 0) out.cigar.push_back(((uint32_t)start_i << 4) | OP_S);
    for (size_t k = 0; k < ops.size();) { size_t e2 = k; while (e2 < ops.size() && ops[e2] == ops[k]) e2++; out.cigar.push_back(((uint32_t)(e2 - k) << 4) | ops[k]); k = e2; }
    if (best_i < len - 1) out.cigar.push_back(((uint32_t)(len - 1 - best_i) << 4) | OP_S);
    return true;
  }
    return false;
}

bool ContigMapper::rescue(const std::string& seq, const Placement& mate, Placement& out) const {
    const int len = (int)seq.size(); if (len < p_.min_score) return false;
    const bool rev = !mate.reverse;                         // the mates of a pair lie on opposite strands
    const std::string s = rev ? revcomp(seq) : seq;
    std::vector<uint8_t> rd((size_t)len); for (int i = 0; i < len; i++) rd[(size_t)i] = (uint8_t)code_of((uint8_t)s[(size_t)i]);
    out.reverse = rev;
    return align(rd.data(), len, mate.contig, mate.pos, out, p_.rescue_window);
}

std::vector<Placement> ContigMapper::map_read(const std::string& seq) const {
    std::vector<Placement> res;
    const int len = (int)seq.size(); if (len < p_.k) return res;
    const uint64_t mask = (1ull << (2 * p_.k)) - 1;
    std::vector<uint8_t> rd((size_t)len); std::vector<Vote> votes;
    for (int strand = 0; strand < 2; strand++) {
        const std::string s = strand ? revcomp(seq) : seq;
        for (int i = 0; i < len; i++) rd[(size_t)i] = (uint8_t)code_of((uint8_t)s[(size_t)i]);
        for (int pass = 0; pass < 2; pass++) {   // second pass only if nothing voted: then the frequent k-mers speak too
            votes.clear(); uint64_t key = 0; int run = 0;
            for (int i = 0; i < len; i++) {
                if (rd[(size_t)i] > 3) { run = 0; key = 0; continue; }
                key = ((key << 2) | rd[(size_t)i]) & mask; run++;
                const int st = i - p_.k + 1;
                if (run < p_.k || st % p_.read_step) continue;
                const size_t bk = (size_t)(key >> bucket_shift_);
                auto lo = std::lower_bound(keys_.begin() + bucket_[bk], keys_.begin() + bucket_[bk + 1], key), hi = lo;
                while (hi != keys_.end() && *hi == key) ++hi;
                if (pass == 0 && hi - lo > p_.max_occ) continue;
                for (auto it = lo; it != hi; ++it) {
                    const size_t ix = (size_t)(it - keys_.begin()); const int32_t c = (int32_t)ctg_[ix];
                    votes.push_back(Vote{c, (int32_t)((int64_t)pos_[ix] - g_.contig_off[(size_t)c] - st)});
                }
            }
            if (!votes.empty()) break;
        }
        std::sort(votes.begin(), votes.end(), [](const Vote& a, const Vote& b) { return a.contig != b.contig ? a.contig < b.contig : a.diag < b.diag; });
        struct Cand { int32_t contig, diag, n; };
        std::vector<Cand> cands;
        for (size_t a = 0; a < votes.size();) {   // votes of one contig whose diagonals lie within half a band of the first one form a candidate
            size_t e = a; while (e < votes.size() && votes[e].contig == votes[a].contig && votes[e].diag - votes[a].diag <= p_.band / 2) e++;
            if ((int)(e - a) >= p_.min_votes) cands.push_back(Cand{votes[a].contig, votes[a + (e - a) / 2].diag, (int32_t)(e - a)});
            a = e;
        }
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.n > b.n; });
        if ((int)cands.size() > p_.max_candidates) cands.resize((size_t)p_.max_candidates);
        for (const Cand& c : cands) { Placement pl; pl.reverse = strand != 0; if (align(rd.data(), len, c.contig, c.diag, pl)) res.push_back(std::move(pl)); }
    }
    std::sort(res.begin(), res.end(), [](const Placement& a, const Placement& b) {
        if (a.score != b.score) return a.score > b.score;
        if (a.reverse != b.reverse) return !a.reverse;
        if (a.contig != b.contig) return a.contig < b.contig;
        if (a.pos != b.pos) return a.pos < b.pos;
        return a.cigar < b.cigar;
    });
    res.erase(std::unique(res.begin(), res.end(), [](const Placement& a, const Placement& b) { return a.contig == b.contig && a.reverse == b.reverse && a.pos == b.pos && a.cigar == b.cigar; }), res.end());
    if ((int)res.size() > p_.max_candidates) res.resize((size_t)p_.max_candidates);
    return res;
}

namespace {

struct FqRead { std::string name, seq, qual; };

std::vector<FqRead> read_fastq(const std::string& path) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1 << 20);
    std::vector<FqRead> out; std::vector<char> buf(1 << 20);
    auto line = [&](std::string& s) -> bool {
        s.clear();
        for (;;) { if (!gzgets(f, buf.data(), (int)buf.size())) return !s.empty(); s += buf.data(); if (!s.empty() && s.back() == '\n') break; }
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        return true;
    };
    std::string h, s, p, q;
    while (line(h)) {
        if (h.empty()) continue;
        if (h[0] != '@' || !line(s) || !line(p) || !line(q) || p.empty() || p[0] != '+' || s.size() != q.size()) { gzclose(f); throw std::runtime_error(path + ": malformed FASTQ record near " + h.substr(0, 60)); }
        FqRead r; size_t e = h.find_first_of(" \t"); r.name = h.substr(1, e == std::string::npos ? std::string::npos : e - 1);
        if (r.name.size() > 2 && r.name[r.name.size() - 2] == '/' && (r.name.back() == '1' || r.name.back() == '2')) r.name.resize(r.name.size() - 2);
        for (char& c : s) if (c >= 'a' && c <= 'z') c = (char)(c - 32);   // soft-masked bases are bases
        r.seq = s; r.qual = q; out.push_back(std::move(r));
    }
    gzclose(f);
    return out;
}

} // namespace

void map_fastq_pairs(const FlatGraph& g, const std::string& fastq1, const std::string& fastq2, int threads, const MapperParams& p, BamBatch& out) {
    std::vector<FqRead> r[2] = {read_fastq(fastq1), read_fastq(fastq2)};
    if (r[0].size() != r[1].size()) throw std::runtime_error("the FASTQ files hold different numbers of reads (" + std::to_string(r[0].size()) + " / " + std::to_string(r[1].size()) + ")");
    const size_t n = r[0].size();
    for (size_t i = 0; i < n; i++) if (r[0][i].name != r[1][i].name) throw std::runtime_error("read " + std::to_string(i + 1) + " of the FASTQ files: names differ (" + r[0][i].name + " / " + r[1][i].name + ")");
    const bool timing = getenv("HLALA_MAPPER_TIMING") != nullptr; auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (timing) { auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[mapper] %-28s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count()); t0 = t1; } };
    lap("FASTQ files read");
    ContigMapper mapper(g, p);
    lap("contig index built");
    std::vector<std::vector<Placement>> hits[2]; hits[0].resize(n); hits[1].resize(n);
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    std::atomic<size_t> next(0);
    auto work = [&]() { for (;;) { const size_t a = next.fetch_add(64); if (a >= n) break; for (size_t i = a; i < std::min(n, a + 64); i++) for (int m = 0; m < 2; m++) hits[m][i] = mapper.map_read(r[m][i].seq); } };
    { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(work); for (auto& t : th) t.join(); }
    lap("reads placed");
    std::atomic<int64_t> rescued(0);
    {   // mate rescue for the pairs with exactly one placed mate
        std::atomic<size_t> nx(0);
        auto rw = [&]() { for (;;) { const size_t a = nx.fetch_add(256); if (a >= n) break; for (size_t i = a; i < std::min(n, a + 256); i++) {
            const bool h0 = !hits[0][i].empty(), h1 = !hits[1][i].empty(); if (h0 == h1) continue;
            const int m = h0 ? 1 : 0; Placement pl;
            if (mapper.rescue(r[m][i].seq, hits[m ^ 1][i][0], pl)) { hits[m][i].push_back(std::move(pl)); rescued++; } } } };
        std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(rw); for (auto& t : th) t.join();
        lap("mates rescued");
    }
    std::vector<size_t> order(n); std::iota(order.begin(), order.end(), (size_t)0);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return r[0][a].name != r[0][b].name ? r[0][a].name < r[0][b].name : a < b; });
    out = BamBatch(); out.read_off.push_back(0); out.chain_off.push_back(0); out.cigar_off.push_back(0);
    out.tlen_n = rescued.load();   // for FASTQ input hlala_bam_batch_stats' is_n reports the mates placed by rescue
    BamBatch::Sample& S = out.is_sample; S.read_off.push_back(0); S.chain_off.push_back(0); S.cigar_off.push_back(0);
    out.names_seen = (int64_t)n;
    for (size_t oi = 0; oi < n; oi++) {
        const size_t i = order[oi];
        out.records += (int64_t)(hits[0][i].size() + hits[1][i].size());
        if (hits[0][i].empty() || hits[1][i].empty()) { out.pairs_incomplete++; continue; }
        if (oi > 0 && r[0][i].name == r[0][order[oi - 1]].name) throw std::runtime_error("read name " + r[0][i].name + " occurs twice");
        out.pair_name.push_back(r[0][i].name);
        const bool sample = (int64_t)S.pair_name.size() < 2000; if (sample) S.pair_name.push_back(r[0][i].name);
        for (int m = 0; m < 2; m++) {
            const std::vector<Placement>& H = hits[m][i]; const Placement& prim = H[0]; const bool mate_rev = hits[m ^ 1][i][0].reverse;
            const std::string seq = prim.reverse ? revcomp(r[m][i].seq) : r[m][i].seq; std::string qual = r[m][i].qual; if (prim.reverse) std::reverse(qual.begin(), qual.end());
            out.bases.insert(out.bases.end(), seq.begin(), seq.end()); out.quals.insert(out.quals.end(), qual.begin(), qual.end()); out.read_off.push_back((int64_t)out.bases.size());
            std::vector<size_t> ord(H.size()); std::iota(ord.begin(), ord.end(), (size_t)0);
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return H[a].contig != H[b].contig ? H[a].contig < H[b].contig : H[a].pos < H[b].pos; });   // a coordinate-sorted BAM
            for (size_t k : ord) {
                const uint16_t fl = (uint16_t)(0x1 | (m == 0 ? 0x40 : 0x80) | (H[k].reverse ? 0x10 : 0) | (mate_rev ? 0x20 : 0) | (k == 0 ? 0 : 0x100));
                out.chain_contig.push_back(H[k].contig); out.chain_pos.push_back(H[k].pos); out.chain_flag.push_back(fl); out.chain_as.push_back(H[k].score);
                out.cigar.insert(out.cigar.end(), H[k].cigar.begin(), H[k].cigar.end()); out.cigar_off.push_back((int32_t)out.cigar.size());
                out.records_used++;
            }
            out.chain_off.push_back((int32_t)out.chain_contig.size());
            if (sample) {
                S.bases.insert(S.bases.end(), seq.begin(), seq.end()); S.quals.insert(S.quals.end(), qual.begin(), qual.end()); S.read_off.push_back((int64_t)S.bases.size());
                S.chain_contig.push_back(prim.contig); S.chain_pos.push_back(prim.pos); S.chain_as.push_back(prim.score);
                S.chain_flag.push_back((uint16_t)(0x1 | (m == 0 ? 0x40 : 0x80) | (prim.reverse ? 0x10 : 0) | (mate_rev ? 0x20 : 0)));
                S.cigar.insert(S.cigar.end(), prim.cigar.begin(), prim.cigar.end()); S.cigar_off.push_back((int32_t)S.cigar.size()); S.chain_off.push_back((int32_t)S.chain_contig.size());
            }
        }
    }
    lap("batch assembled");
    S.names_seen = (int64_t)S.pair_name.size();
    S.loaded_contigs.resize((size_t)g.n_contigs); std::iota(S.loaded_contigs.begin(), S.loaded_contigs.end(), 0);   // every translation is at hand here
}

} // namespace hlala
