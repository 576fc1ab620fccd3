// See bam_reader.h. BGZF: RFC 1952 members with a 'BC' extra subfield holding the block size (SAM/BAM specification §4.1).
#include "bam_reader.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <functional>
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

void read_bam_seeds(const std::string& path, const std::vector<std::string>& contig_names, const std::vector<int64_t>& contig_len, int threads, BamBatch& out, bool long_reads) {
    // long_reads (extractSeeds2 with a longReadMode, processBAM.cpp:732-738, 781-784, 816-819): only primary records are kept, records need not be paired, and every record
    // counts as "mate 1" of its read name; a read is complete when it holds a primary record (protoSeeds::isComplete_unpaired). The batch keeps its pair layout:
    // read 2p is the long read, read 2p+1 an empty mate (no bases, no chains).
    auto mate_of = [long_reads](uint16_t flag) { return long_reads ? 0 : ((flag & 0x40) ? 0 : 1); };
    out = BamBatch();
    const bool trace = getenv("HLALA_BAM_TRACE") != nullptr; auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (!trace) return; auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[bam] %-10s %6.2f s\n", what, std::chrono::duration<double>(t1 - t0).count()); t0 = t1; };
    const std::vector<uint8_t> raw = read_file(path); lap("read");
    const std::vector<uint8_t> d = inflate_bgzf(raw, threads, path); lap("inflate");
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
    // ---- records: boundaries in one cheap sequential walk, then parsed and filtered on the thread pool (what extractSeeds2 keeps)
    std::vector<size_t> rec_at;
    while (p < d.size()) { need(4); const uint32_t bs = le32(&d[p]); if (bs < 32) throw std::runtime_error(path + ": BAM record shorter than its fixed part"); need(4 + (size_t)bs); rec_at.push_back(p + 4); p += 4 + (size_t)bs; }
    out.records = (int64_t)rec_at.size();
    const size_t NR = rec_at.size();
    std::vector<Rec> recs(NR); std::vector<uint8_t> keep(NR, 0);
    std::atomic<size_t> next_blk(0); std::atomic<int> err_kind(0);   // 1 fields exceed record, 2 unpaired, 3 aux malformed, 4 no AS
    auto parse = [&]() {
        const size_t blk = 16384;
        for (;;) {
            const size_t i0 = next_blk.fetch_add(1) * blk; if (i0 >= NR) break; const size_t i1 = std::min(NR, i0 + blk);
            for (size_t i = i0; i < i1; i++) {
                const uint8_t* r = &d[rec_at[i]]; const uint32_t bs = le32(r - 4);
                const int32_t ref = (int32_t)le32(r), pos = (int32_t)le32(r + 4); const uint32_t l_read_name = r[8]; const uint32_t n_cigar = le16(r + 12); const uint16_t flag = le16(r + 14); const int32_t l_seq = (int32_t)le32(r + 16);
                const size_t at_name = 32, at_cigar = at_name + l_read_name, at_seq = at_cigar + 4ull * n_cigar, at_qual = at_seq + (size_t)(l_seq + 1) / 2, at_aux = at_qual + (size_t)l_seq;
                if (l_seq < 0 || at_aux > bs) { err_kind = 1; continue; }
                if (flag & 0x4) continue;                                   // IsMapped (processBAM.cpp:727)
                if (long_reads && (flag & 0x100)) continue;                 // long-read mode: primary records only (:732-738)
                if (ref < 0 || ref >= (int32_t)n_ref || ref2contig[ref] < 0) continue;   // not an interesting contig (:739)
                if (n_cigar == 0) continue;                                 // :755
                int64_t reflen = 0; for (uint32_t k = 0; k < n_cigar; k++) { const uint32_t c = le32(r + at_cigar + 4 * k); const int op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += c >> 4; }
                const int contig = ref2contig[ref]; const int64_t stop = (int64_t)pos + reflen - 1;
                if (!(pos >= 0 && pos <= contig_len[contig] - 1 && stop >= 0 && stop <= contig_len[contig] - 1)) continue;   // inside the interval = the whole contig (:764-766)
                if (!long_reads && !(flag & 0x1)) { err_kind = 2; continue; }
                int32_t as = 0; bool have_as = false; bool bad = false;
                for (size_t a = at_aux; a + 3 <= bs && !bad;) {   // aux fields: tag[2] type value
                    const char t0 = (char)r[a], t1 = (char)r[a + 1], ty = (char)r[a + 2]; a += 3; size_t len = 0; int64_t v = 0; bool is_int = true;
                    switch (ty) {
                    case 'A': len = 1; is_int = false; break; case 'c': len = 1; if (a < bs) v = (int8_t)r[a]; break; case 'C': len = 1; if (a < bs) v = r[a]; break;
                    case 's': len = 2; if (a + 2 <= bs) v = (int16_t)le16(r + a); break; case 'S': len = 2; if (a + 2 <= bs) v = le16(r + a); break;
                    case 'i': len = 4; if (a + 4 <= bs) v = (int32_t)le32(r + a); break; case 'I': len = 4; if (a + 4 <= bs) v = le32(r + a); break;
                    case 'f': len = 4; is_int = false; break;
                    case 'Z': case 'H': { size_t e = a; while (e < bs && r[e]) e++; len = e - a + 1; is_int = false; break; }
                    case 'B': { if (a + 5 > bs) { bad = true; break; } const char st = (char)r[a]; const uint32_t cnt = le32(r + a + 1); const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; len = 5 + es * cnt; is_int = false; break; }
                    default: bad = true; break;
                    }
                    if (bad || a + len > bs) { bad = true; break; }
                    if (t0 == 'A' && t1 == 'S' && is_int) { as = (int32_t)v; have_as = true; }
                    a += len;
                }
                if (bad) { err_kind = 3; continue; }
                if (!have_as) { err_kind = 4; continue; }
                Rec& rec = recs[i]; rec.ref = contig; rec.pos = pos; rec.as = as; rec.flag = flag; rec.cigar_at = (uint32_t)0; rec.n_cigar = n_cigar; rec.seq_at = rec_at[i] + at_seq; rec.l_seq = l_seq; rec.order = i;
                keep[i] = 1;
            }
        }
    };
    { const int nt = std::max(1, std::min(threads, (int)std::max<size_t>(1, NR / 65536))); std::vector<std::thread> th; for (int t = 1; t < nt; t++) th.emplace_back(parse); parse(); for (auto& t : th) t.join(); }
    lap("parse");
    switch (err_kind.load()) {
    case 1: throw std::runtime_error(path + ": BAM record fields exceed the record");
    case 2: throw std::runtime_error(path + ": record of an unpaired read (the reference asserts IsPaired, processBAM.cpp:780)");
    case 3: throw std::runtime_error(path + ": malformed aux field in BAM record");
    case 4: throw std::runtime_error(path + ": record without an integer AS tag (the reference asserts it, processBAM.cpp:4314-4336)");
    default: break;
    }
    // ---- group by read name (names are compared as bytes, like the keys of the reference's std::map<std::string, protoSeeds>)
    struct NameRef { const char* p; uint32_t n; };
    struct NameHash { size_t operator()(const NameRef& x) const { uint64_t h = 1469598103934665603ull; for (uint32_t i = 0; i < x.n; i++) h = (h ^ (uint8_t)x.p[i]) * 1099511628211ull; return (size_t)h; } };
    struct NameEq { bool operator()(const NameRef& a, const NameRef& b) const { return a.n == b.n && memcmp(a.p, b.p, a.n) == 0; } };
    // open-addressing table of group ids (linear probing, load <= 0.5); the name hashes are computed on the thread pool first, the insertion itself runs in
    // file order so that group ids follow the first appearance of a name
    std::vector<uint64_t> name_hash(NR, 0);
    { std::atomic<size_t> nb(0); auto hash_all = [&]() { const size_t blk = 65536; NameHash H;
          for (;;) { const size_t i0 = nb.fetch_add(1) * blk; if (i0 >= NR) break; const size_t i1 = std::min(NR, i0 + blk);
              for (size_t i = i0; i < i1; i++) if (keep[i]) { const uint8_t* r = &d[rec_at[i]]; const uint32_t ln = r[8]; name_hash[i] = (uint64_t)H(NameRef{(const char*)r + 32, ln ? ln - 1 : 0}); } } };
      std::vector<std::thread> th; const int nt = std::max(1, std::min(threads, 16)); for (int t = 1; t < nt; t++) th.emplace_back(hash_all); hash_all(); for (auto& t : th) t.join(); }
    size_t n_keep = 0; for (size_t i = 0; i < NR; i++) n_keep += keep[i] ? 1 : 0;
    size_t tcap = 64; while (tcap < 2 * n_keep + 16) tcap <<= 1;
    std::vector<int32_t> table(tcap, -1);
    std::vector<NameRef> gname; std::vector<int32_t> rec_gid(NR, -1); std::vector<int32_t> gcount; gname.reserve(n_keep / 8 + 16); gcount.reserve(n_keep / 8 + 16);
    { NameEq EQ;
      for (size_t i = 0; i < NR; i++) if (keep[i]) {
        const uint8_t* r = &d[rec_at[i]]; const uint32_t ln = r[8]; NameRef nm{(const char*)r + 32, ln ? ln - 1 : 0};
        size_t h = (size_t)(name_hash[i] ^ (name_hash[i] >> 29)) & (tcap - 1); int32_t g = -1;
        for (;;) { const int32_t v = table[h]; if (v < 0) break; if (EQ(gname[(size_t)v], nm)) { g = v; break; } h = (h + 1) & (tcap - 1); }
        if (g < 0) { g = (int32_t)gname.size(); table[h] = g; gname.push_back(nm); gcount.push_back(0); }
        rec_gid[i] = g; gcount[(size_t)g]++; out.records_used++;
      } }
    // ---- the insert-size sample (see bam_reader.h): replay of extractSeeds(4000) over the kept records
    if (!long_reads) {     // the insert-size sample (paired reads only)
        std::vector<uint32_t> scan; scan.reserve((size_t)out.records_used); for (size_t i = 0; i < NR; i++) if (keep[i]) scan.push_back((uint32_t)i);
        std::vector<int32_t> contig_rank(contig_names.size());      // byte order of the contig names = iteration order of the reference's interval map
        { std::vector<int32_t> o(contig_names.size()); for (size_t i = 0; i < o.size(); i++) o[i] = (int32_t)i; std::sort(o.begin(), o.end(), [&](int32_t x, int32_t y) { return contig_names[(size_t)x] < contig_names[(size_t)y]; }); for (size_t k = 0; k < o.size(); k++) contig_rank[(size_t)o[k]] = (int32_t)k; }
        {   // stable counting sort by contig rank: file order inside a contig
            std::vector<size_t> at(contig_names.size() + 1, 0); for (uint32_t i : scan) at[(size_t)contig_rank[(size_t)recs[i].ref] + 1]++;
            for (size_t k = 1; k < at.size(); k++) at[k] += at[k - 1];
            std::vector<uint32_t> sorted(scan.size()); for (uint32_t i : scan) sorted[at[(size_t)contig_rank[(size_t)recs[i].ref]]++] = i;
            scan.swap(sorted);
        }
        struct St { int32_t first1 = -1, first2 = -1; bool prim1 = false, prim2 = false; };   // first primary record of either mate, in scan order
        std::vector<St> state(gname.size()); std::vector<uint8_t> seen(gname.size(), 0); std::vector<uint8_t> loaded(contig_names.size(), 0);
        // The reference leaves the record loop of the CURRENT interval when both thresholds are met (processBAM.cpp:672-676) but then moves on to the next
        // interval (:691-693 set tryAgain again): once the thresholds hold, every later contig still contributes its first record (and gets its translation
        // loaded) before the test fires again. Reproduced, because those records can complete a pair and those translations anchor the distances.
        long long included = 0, included_complete = 0; const long long want = 4000; int32_t skip_contig = -1;
        for (uint32_t i : scan) {
            const Rec& r = recs[i];
            if (r.ref == skip_contig) continue;
            const int32_t g = rec_gid[i]; St& s = state[(size_t)g];
            const bool first_mate = (r.flag & 0x40) != 0, primary = !(r.flag & 0x100);
            if (first_mate) { if (primary && s.first1 < 0) s.first1 = (int32_t)i; s.prim1 = s.prim1 || primary; } else { if (primary && s.first2 < 0) s.first2 = (int32_t)i; s.prim2 = s.prim2 || primary; }
            loaded[(size_t)r.ref] = 1;
            if (!seen[(size_t)g]) { seen[(size_t)g] = 1; included++; }
            if (s.prim1 && s.prim2) included_complete++;
            if (included >= want && included_complete >= want / 2) skip_contig = r.ref;     // the rest of this contig is not read
        }
        BamBatch::Sample& S = out.is_sample; S = BamBatch::Sample();
        for (size_t c = 0; c < loaded.size(); c++) if (loaded[c]) S.loaded_contigs.push_back((int32_t)c);
        std::vector<int32_t> order; for (size_t g = 0; g < gname.size(); g++) if (seen[g]) { S.names_seen++; if (state[g].prim1 && state[g].prim2) order.push_back((int32_t)g); else S.incomplete++; }
        std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { const NameRef& a = gname[(size_t)x]; const NameRef& b = gname[(size_t)y]; const int c = memcmp(a.p, b.p, std::min(a.n, b.n)); return c < 0 || (c == 0 && a.n < b.n); });
        static const char SEQ16s[] = "=ACMGRSVTWYHKDBN";
        S.read_off.assign(1, 0); S.chain_off.assign(1, 0); S.cigar_off.assign(1, 0);
        for (int32_t g : order) {
            S.pair_name.emplace_back(gname[(size_t)g].p, gname[(size_t)g].n);
            for (int m = 0; m < 2; m++) {
                const Rec& c = recs[(size_t)(m ? state[(size_t)g].first2 : state[(size_t)g].first1)];
                for (int32_t k = 0; k < c.l_seq; k++) { const uint8_t b = d[c.seq_at + (size_t)k / 2]; S.bases.push_back((uint8_t)SEQ16s[(k & 1) ? (b & 15) : (b >> 4)]); }
                const uint8_t* q = &d[c.seq_at + (size_t)(c.l_seq + 1) / 2]; for (int32_t k = 0; k < c.l_seq; k++) S.quals.push_back((uint8_t)(q[k] + 33));
                S.read_off.push_back((int64_t)S.bases.size());
                S.chain_contig.push_back(c.ref); S.chain_pos.push_back(c.pos); S.chain_flag.push_back(c.flag); S.chain_as.push_back(c.as);
                for (uint32_t k = 0; k < c.n_cigar; k++) S.cigar.push_back(le32(&d[c.seq_at] - 4ull * c.n_cigar + 4ull * k));
                S.cigar_off.push_back((int32_t)S.cigar.size()); S.chain_off.push_back((int32_t)S.chain_contig.size());
            }
        }
    }
    const size_t NG = gname.size(); out.names_seen = (int64_t)NG;
    std::vector<int64_t> goff(NG + 1, 0); for (size_t g = 0; g < NG; g++) goff[g + 1] = goff[g] + gcount[g];
    std::vector<uint32_t> grec((size_t)goff[NG]); { std::vector<int64_t> at(goff.begin(), goff.end() - 1); for (size_t i = 0; i < NR; i++) if (keep[i]) grec[(size_t)at[(size_t)rec_gid[i]]++] = (uint32_t)i; }   // file order inside a group
    std::vector<int32_t> gorder(NG); for (size_t g = 0; g < NG; g++) gorder[g] = (int32_t)g;
    std::sort(gorder.begin(), gorder.end(), [&](int32_t x, int32_t y) { const NameRef& a = gname[(size_t)x]; const NameRef& b = gname[(size_t)y]; const int c = memcmp(a.p, b.p, std::min(a.n, b.n)); return c < 0 || (c == 0 && a.n < b.n); });
    lap("group");
    auto cigar_of = [&](const Rec& c, uint32_t k) { return le32(&d[c.seq_at] - 4ull * c.n_cigar + 4ull * k); };   // the CIGAR sits right before SEQ
    // ---- complete pairs -> flat batch: per group the mates' records and the record that stands for each read (thread pool), offsets by
    //      one sequential pass over the groups in name order, then the arrays are filled in parallel
    static const char SEQ16[] = "=ACMGRSVTWYHKDBN";
    struct GInfo { int32_t n0 = 0, prim0 = -1, prim1 = -1; uint8_t complete = 0; };
    std::vector<GInfo> ginfo(NG); std::vector<uint32_t> gm(grec.size());   // gm: a group's records, first mates first, file order inside each
    const int nt_pool = std::max(1, std::min(threads, (int)std::max<size_t>(1, NG / 4096)));
    auto pool = [&](const std::function<void()>& fn) { std::vector<std::thread> th; for (int t = 1; t < nt_pool; t++) th.emplace_back(fn); fn(); for (auto& t : th) t.join(); };
    std::atomic<size_t> nextg(0);
    pool([&]() { std::vector<int> idx; const size_t blk = 2048;
        for (;;) { const size_t g0 = nextg.fetch_add(1) * blk; if (g0 >= NG) break; const size_t g1 = std::min(NG, g0 + blk);
            for (size_t g = g0; g < g1; g++) {
                GInfo& I = ginfo[g]; int64_t w = goff[g];
                for (int m = 0; m < 2; m++) { for (int64_t j = goff[g]; j < goff[g + 1]; j++) { const uint32_t i = grec[(size_t)j]; if (mate_of(recs[i].flag) == m) gm[(size_t)w++] = i; } if (m == 0) I.n0 = (int32_t)(w - goff[g]); }
                for (int m = 0; m < 2; m++) {
                    // the record whose SEQ/QUAL stand for the read: the first primary record after sortChainsInSeeds (processBAM.cpp:1945: std::sort ascending by AS, then std::reverse)
                    const uint32_t* mr = gm.data() + goff[g] + (m ? I.n0 : 0); const int nm = m ? (int)(goff[g + 1] - goff[g]) - I.n0 : I.n0;
                    idx.resize((size_t)nm); for (int i = 0; i < nm; i++) idx[(size_t)i] = i;
                    std::sort(idx.begin(), idx.end(), [&](int x, int y) { return recs[mr[x]].as < recs[mr[y]].as; }); std::reverse(idx.begin(), idx.end());
                    int prim = -1; for (int i : idx) if (!(recs[mr[i]].flag & 0x100)) { prim = i; break; }
                    (m ? I.prim1 : I.prim0) = prim;
                }
                I.complete = long_reads ? (I.prim0 >= 0 ? 1 : 0) : ((I.prim0 >= 0 && I.prim1 >= 0) ? 1 : 0);   // protoSeeds::isComplete / isComplete_unpaired
            } } });
    std::vector<int32_t> pair_group; pair_group.reserve(NG);
    out.read_off.assign(1, 0); out.chain_off.assign(1, 0);
    std::vector<int64_t> pair_cigar_at(1, 0);
    for (size_t go = 0; go < NG; go++) {
        const size_t g = (size_t)gorder[go]; const GInfo& I = ginfo[g];
        if (!I.complete) { out.pairs_incomplete++; continue; }
        pair_group.push_back((int32_t)g);
        int64_t ncig = 0;
        for (int m = 0; m < 2; m++) {
            const uint32_t* mr = gm.data() + goff[g] + (m ? I.n0 : 0); const int nm = m ? (int)(goff[g + 1] - goff[g]) - I.n0 : I.n0;
            out.read_off.push_back(out.read_off.back() + ((long_reads && m) ? 0 : recs[mr[m ? I.prim1 : I.prim0]].l_seq));
            out.chain_off.push_back(out.chain_off.back() + nm);
            for (int i = 0; i < nm; i++) ncig += recs[mr[i]].n_cigar;
        }
        pair_cigar_at.push_back(pair_cigar_at.back() + ncig);
    }
    const size_t NPAIR = pair_group.size(); const size_t NCH = (size_t)out.chain_off.back();
    if (pair_cigar_at.back() > 0x7fffffffll) throw std::runtime_error(path + ": more than 2^31 CIGAR operations");
    out.pair_name.resize(NPAIR); out.bases.resize((size_t)out.read_off.back()); out.quals.resize((size_t)out.read_off.back());
    out.chain_contig.resize(NCH); out.chain_pos.resize(NCH); out.chain_flag.resize(NCH); out.chain_as.resize(NCH); out.cigar_off.assign(NCH + 1, 0); out.cigar.resize((size_t)pair_cigar_at.back());
    std::atomic<size_t> nextp(0); std::vector<int64_t> part_s1((size_t)nt_pool, 0), part_s2((size_t)nt_pool, 0), part_n((size_t)nt_pool, 0); std::atomic<int> tid(0);   // integer sums: the estimate does not depend on the thread count std::atomic<int> tid(0);
    pool([&]() { const int me = tid.fetch_add(1); const size_t blk = 1024; int64_t s1 = 0, s2 = 0, sn = 0;
        for (;;) { const size_t p0 = nextp.fetch_add(1) * blk; if (p0 >= NPAIR) break; const size_t p1 = std::min(NPAIR, p0 + blk);
            for (size_t pi = p0; pi < p1; pi++) {
                const size_t g = (size_t)pair_group[pi]; const GInfo& I = ginfo[g];
                out.pair_name[pi].assign(gname[g].p, gname[g].n);
                int64_t cg = pair_cigar_at[pi];
                for (int m = 0; m < 2; m++) {
                    const uint32_t* mr = gm.data() + goff[g] + (m ? I.n0 : 0); const int nm = m ? (int)(goff[g + 1] - goff[g]) - I.n0 : I.n0;
                    if (long_reads && m) continue;     // the empty mate
                    const Rec& P = recs[mr[m ? I.prim1 : I.prim0]]; const int64_t b0 = out.read_off[2 * pi + (size_t)m];
                    for (int32_t i = 0; i < P.l_seq; i++) { const uint8_t b = d[P.seq_at + (size_t)i / 2]; out.bases[(size_t)b0 + (size_t)i] = (uint8_t)SEQ16[(i & 1) ? (b & 15) : (b >> 4)]; }
                    const uint8_t* q = &d[P.seq_at + (size_t)(P.l_seq + 1) / 2];
                    for (int32_t i = 0; i < P.l_seq; i++) out.quals[(size_t)b0 + (size_t)i] = (uint8_t)(q[i] + 33);
                    int32_t c0 = out.chain_off[2 * pi + (size_t)m];
                    for (int i = 0; i < nm; i++) { const Rec& c = recs[mr[i]]; const size_t ci = (size_t)c0 + (size_t)i;
                        out.chain_contig[ci] = c.ref; out.chain_pos[ci] = c.pos; out.chain_flag[ci] = c.flag; out.chain_as[ci] = c.as;
                        for (uint32_t k = 0; k < c.n_cigar; k++) out.cigar[(size_t)cg++] = cigar_of(c, k);
                        out.cigar_off[ci + 1] = (int32_t)cg; }
                }
                // insert-size sample: both primaries on one contig, opposite strands, forward mate upstream; gap = start of the downstream mate - end of the upstream mate - 1
                if (long_reads) continue;
                const Rec& A = recs[gm[(size_t)goff[g] + (size_t)I.prim0]]; const Rec& B = recs[gm[(size_t)goff[g] + (size_t)I.n0 + (size_t)I.prim1]];
                if (A.ref == B.ref && ((A.flag ^ B.flag) & 0x10)) {
                    auto end_of = [&](const Rec& c) { int64_t rl = 0; for (uint32_t k = 0; k < c.n_cigar; k++) { const uint32_t cgv = cigar_of(c, k); const int op = cgv & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += cgv >> 4; } return (int64_t)c.pos + rl - 1; };
                    const Rec& F = (A.flag & 0x10) ? B : A; const Rec& Rv = (A.flag & 0x10) ? A : B;
                    if (F.pos <= Rv.pos) { const int64_t gap = (int64_t)Rv.pos - end_of(F) - 1; if (gap > -2000 && gap < 2000) { s1 += gap; s2 += gap * gap; sn++; } }
                }
            } }
        part_s1[(size_t)me] = s1; part_s2[(size_t)me] = s2; part_n[(size_t)me] = sn; });
    int64_t s1i = 0, s2i = 0; for (int t = 0; t < nt_pool; t++) { s1i += part_s1[(size_t)t]; s2i += part_s2[(size_t)t]; out.tlen_n += part_n[(size_t)t]; }
    const double s1 = (double)s1i, s2 = (double)s2i;
    lap("assemble");
    if (out.tlen_n > 1) { out.tlen_mean = s1 / (double)out.tlen_n; const double var = s2 / (double)out.tlen_n - out.tlen_mean * out.tlen_mean; out.tlen_sd = var > 0 ? sqrt(var) : 0; }
}

} // namespace hlala
