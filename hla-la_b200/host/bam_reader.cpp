// See bam_reader.h. BGZF: RFC 1952 members with a 'BC' extra subfield holding the block size (SAM/BAM specification §4.1).
#include "bam_reader.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <thread>
#include <unordered_map>
#include <zlib.h>

namespace hlala {

namespace {

struct Block { size_t in_off, in_len, out_off, out_len; };

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb"); if (!f) throw std::runtime_error("Cannot open BAM file " + path);
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v((size_t)std::max<long>(n, 0));
    if (n > 0 && fread(v.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); throw std::runtime_error("Cannot read BAM file " + path); }
    fclose(f); return v;
}
inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// all BGZF blocks of the file inflated into one buffer
std::vector<uint8_t> inflate_bgzf(const std::vector<uint8_t>& in, int threads, const std::string& path) {
    std::vector<Block> blocks; size_t pos = 0, out = 0;
    while (pos < in.size()) {
        if (pos + 18 > in.size() || in[pos] != 31 || in[pos + 1] != 139 || in[pos + 2] != 8 || !(in[pos + 3] & 4)) throw std::runtime_error(path + ": not a BGZF file (bad block header)");
        const size_t xlen = le16(&in[pos + 10]); size_t x = pos + 12; const size_t xend = x + xlen; long bsize = -1;
        if (xend > in.size()) throw std::runtime_error(path + ": truncated BGZF block");
        while (x + 4 <= xend) { const size_t slen = le16(&in[x + 2]); if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2) bsize = le16(&in[x + 4]); x += 4 + slen; }
        if (bsize < 0) throw std::runtime_error(path + ": BGZF block without BC subfield");
        const size_t total = (size_t)bsize + 1;
        if (pos + total > in.size() || total < xlen + 20) throw std::runtime_error(path + ": truncated BGZF block");
        const size_t isize = le32(&in[pos + total - 4]);
        blocks.push_back({xend, pos + total - 8 - xend, out, isize}); out += isize; pos += total;
    }
    std::vector<uint8_t> data(out);
    std::atomic<size_t> next(0); std::atomic<bool> bad(false);
    auto work = [&]() {
        z_stream zs; memset(&zs, 0, sizeof zs); if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
        for (;;) {
            const size_t b = next.fetch_add(1); if (b >= blocks.size()) break;
            const Block& B = blocks[b]; if (B.out_len == 0) continue;
            inflateReset(&zs);
            zs.next_in = const_cast<Bytef*>(in.data() + B.in_off); zs.avail_in = (uInt)B.in_len; zs.next_out = data.data() + B.out_off; zs.avail_out = (uInt)B.out_len;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) bad = true;
        }
        inflateEnd(&zs);
    };
    const int nt = std::max(1, std::min(threads, (int)std::max<size_t>(1, blocks.size() / 4)));
    std::vector<std::thread> th; for (int t = 1; t < nt; t++) th.emplace_back(work);
    work(); for (auto& t : th) t.join();
    if (bad) throw std::runtime_error(path + ": BGZF block does not inflate");
    return data;
}

struct Rec { int32_t ref, pos, as; uint16_t flag; uint32_t cigar_at, n_cigar; size_t seq_at; int32_t l_seq; size_t order; };

} // namespace

void read_bam_seeds(const std::string& path, const std::vector<std::string>& contig_names, const std::vector<int64_t>& contig_len, int threads, BamBatch& out) {
    out = BamBatch();
    const std::vector<uint8_t> raw = read_file(path);
    const std::vector<uint8_t> d = inflate_bgzf(raw, threads, path);
    size_t p = 0; auto need = [&](size_t n) { if (p + n > d.size()) throw std::runtime_error(path + ": truncated BAM"); };
    need(12); if (memcmp(&d[0], "BAM\1", 4) != 0) throw std::runtime_error(path + ": not a BAM file");
    const uint32_t l_text = le32(&d[4]); p = 8; need(l_text + 4); p += l_text;
    const uint32_t n_ref = le32(&d[p]); p += 4;
    std::unordered_map<std::string, int32_t> prg_contig; for (size_t i = 0; i < contig_names.size(); i++) prg_contig[contig_names[i]] = (int32_t)i;
    std::vector<int32_t> ref2contig(n_ref, -1);
    for (uint32_t r = 0; r < n_ref; r++) {
        need(4); const uint32_t l_name = le32(&d[p]); p += 4; need(l_name + 4);
        std::string name((const char*)&d[p], l_name ? l_name - 1 : 0); p += l_name; p += 4;
        auto it = prg_contig.find(name); if (it != prg_contig.end()) ref2contig[r] = it->second;
    }
    // ---- pass over the records: keep what extractSeeds2 keeps, grouped by name
    struct Group { std::vector<Rec> mate[2]; };
    std::map<std::string, Group> groups;   // byte order of the names == std::map<std::string, protoSeeds>
    std::vector<uint32_t> cig; size_t order = 0;
    while (p < d.size()) {
        need(4); const uint32_t bs = le32(&d[p]); p += 4; need(bs); const uint8_t* r = &d[p]; const size_t rec_end = p + bs; p = rec_end;
        if (bs < 32) throw std::runtime_error(path + ": BAM record shorter than its fixed part");
        out.records++;
        const int32_t ref = (int32_t)le32(r), pos = (int32_t)le32(r + 4); const uint32_t l_read_name = r[8]; const uint32_t n_cigar = le16(r + 12); const uint16_t flag = le16(r + 14); const int32_t l_seq = (int32_t)le32(r + 16);
        const size_t at_name = 32, at_cigar = at_name + l_read_name, at_seq = at_cigar + 4ull * n_cigar, at_qual = at_seq + (size_t)(l_seq + 1) / 2, at_aux = at_qual + (size_t)l_seq;
        if (at_aux > bs) throw std::runtime_error(path + ": BAM record fields exceed the record");
        if (flag & 0x4) continue;                                   // IsMapped (processBAM.cpp:727)
        if (ref < 0 || ref >= (int32_t)n_ref || ref2contig[ref] < 0) continue;   // not an interesting contig (:739)
        if (n_cigar == 0) continue;                                 // :755
        int64_t reflen = 0; for (uint32_t k = 0; k < n_cigar; k++) { const uint32_t c = le32(r + at_cigar + 4 * k); const int op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += c >> 4; }
        const int contig = ref2contig[ref]; const int64_t stop = (int64_t)pos + reflen - 1;
        if (!(pos >= 0 && pos <= contig_len[contig] - 1 && stop >= 0 && stop <= contig_len[contig] - 1)) continue;   // inside the interval = the whole contig (:764-766)
        if (!(flag & 0x1)) throw std::runtime_error(path + ": record of an unpaired read (the reference asserts IsPaired, processBAM.cpp:780)");
        int32_t as = 0; bool have_as = false;
        for (size_t a = at_aux; a + 3 <= bs;) {   // aux fields: tag[2] type value
            const char t0 = (char)r[a], t1 = (char)r[a + 1], ty = (char)r[a + 2]; a += 3; size_t len = 0; int64_t v = 0; bool is_int = true;
            switch (ty) {
            case 'A': len = 1; is_int = false; break; case 'c': len = 1; v = (int8_t)r[a]; break; case 'C': len = 1; v = r[a]; break;
            case 's': len = 2; v = (int16_t)le16(r + a); break; case 'S': len = 2; v = le16(r + a); break; case 'i': len = 4; v = (int32_t)le32(r + a); break; case 'I': len = 4; v = le32(r + a); break;
            case 'f': len = 4; is_int = false; break;
            case 'Z': case 'H': { size_t e = a; while (e < bs && r[e]) e++; len = e - a + 1; is_int = false; break; }
            case 'B': { if (a + 5 > bs) throw std::runtime_error(path + ": truncated aux array"); const char st = (char)r[a]; const uint32_t cnt = le32(r + a + 1); const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; len = 5 + es * cnt; is_int = false; break; }
            default: throw std::runtime_error(path + ": unknown aux type in BAM record");
            }
            if (a + len > bs) throw std::runtime_error(path + ": truncated aux field");
            if (t0 == 'A' && t1 == 'S' && is_int) { as = (int32_t)v; have_as = true; }
            a += len;
        }
        if (!have_as) throw std::runtime_error(path + ": record without an integer AS tag (the reference asserts it, processBAM.cpp:4314-4336)");
        Rec rec; rec.ref = contig; rec.pos = pos; rec.as = as; rec.flag = flag; rec.cigar_at = (uint32_t)cig.size(); rec.n_cigar = n_cigar; rec.seq_at = (size_t)(r - d.data()) + at_seq; rec.l_seq = l_seq; rec.order = order++;
        for (uint32_t k = 0; k < n_cigar; k++) cig.push_back(le32(r + at_cigar + 4 * k));
        std::string name((const char*)r + at_name, l_read_name ? l_read_name - 1 : 0);
        groups[name].mate[(flag & 0x40) ? 0 : 1].push_back(rec);
        out.records_used++;
    }
    out.names_seen = (int64_t)groups.size();
    // ---- complete pairs -> flat batch
    static const char SEQ16[] = "=ACMGRSVTWYHKDBN";
    out.read_off.assign(1, 0); out.chain_off.assign(1, 0); out.cigar_off.assign(1, 0);
    double s1 = 0, s2 = 0;
    for (auto& kv : groups) {
        Group& G = kv.second;
        int prim[2] = {-1, -1};
        for (int m = 0; m < 2; m++) {
            // the record whose SEQ/QUAL stand for the read: the first primary record after sortChainsInSeeds (processBAM.cpp:1945: std::sort ascending by AS, then std::reverse)
            std::vector<int> idx(G.mate[m].size()); for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
            std::sort(idx.begin(), idx.end(), [&](int x, int y) { return G.mate[m][(size_t)x].as < G.mate[m][(size_t)y].as; }); std::reverse(idx.begin(), idx.end());
            for (int i : idx) if (!(G.mate[m][(size_t)i].flag & 0x100)) { prim[m] = i; break; }
        }
        if (prim[0] < 0 || prim[1] < 0) { out.pairs_incomplete++; continue; }   // protoSeeds::isComplete
        out.pair_name.push_back(kv.first);
        for (int m = 0; m < 2; m++) {
            const Rec& P = G.mate[m][(size_t)prim[m]];
            for (int32_t i = 0; i < P.l_seq; i++) { const uint8_t b = d[P.seq_at + (size_t)i / 2]; out.bases.push_back((uint8_t)SEQ16[(i & 1) ? (b & 15) : (b >> 4)]); }
            const uint8_t* q = &d[P.seq_at + (size_t)(P.l_seq + 1) / 2];
            for (int32_t i = 0; i < P.l_seq; i++) out.quals.push_back((uint8_t)(q[i] + 33));
            out.read_off.push_back((int64_t)out.bases.size());
            for (const Rec& c : G.mate[m]) {
                out.chain_contig.push_back(c.ref); out.chain_pos.push_back(c.pos); out.chain_flag.push_back(c.flag); out.chain_as.push_back(c.as);
                out.cigar.insert(out.cigar.end(), cig.begin() + c.cigar_at, cig.begin() + c.cigar_at + c.n_cigar); out.cigar_off.push_back((int32_t)out.cigar.size());
            }
            out.chain_off.push_back((int32_t)out.chain_contig.size());
        }
        // insert-size sample: both primaries on one contig, opposite strands, forward mate upstream; gap = start of the downstream mate - end of the upstream mate - 1
        const Rec& A = G.mate[0][(size_t)prim[0]]; const Rec& B = G.mate[1][(size_t)prim[1]];
        if (A.ref == B.ref && ((A.flag ^ B.flag) & 0x10)) {
            auto end_of = [&](const Rec& c) { int64_t rl = 0; for (uint32_t k = 0; k < c.n_cigar; k++) { const uint32_t cg = cig[c.cigar_at + k]; const int op = cg & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += cg >> 4; } return (int64_t)c.pos + rl - 1; };
            const Rec& F = (A.flag & 0x10) ? B : A; const Rec& Rv = (A.flag & 0x10) ? A : B;
            if (F.pos <= Rv.pos) { const double gap = (double)(Rv.pos - end_of(F) - 1); if (fabs(gap) < 2000) { s1 += gap; s2 += gap * gap; out.tlen_n++; } }
        }
    }
    if (out.tlen_n > 1) { out.tlen_mean = s1 / (double)out.tlen_n; const double var = s2 / (double)out.tlen_n - out.tlen_mean * out.tlen_mean; out.tlen_sd = var > 0 ? sqrt(var) : 0; }
}

} // namespace hlala
