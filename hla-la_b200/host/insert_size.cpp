// See insert_size.h.
#include "insert_size.h"

#include <cmath>
#include <cstdlib>
#include <map>
#include <set>
#include <stdexcept>

namespace hlala {

namespace {
// sequence id -> position along it, from the first of `levels` that knows the sequence (verboseSeedChain.h:231-283)
std::map<int, int> anchors_of(const FlatGraph& g, const int* levels, int n, const std::vector<uint8_t>& id_loaded) {
    std::map<int, int> a;
    for (int i = 0; i < n; i++) {
        const int l = levels[i];
        for (int32_t k = g.anchor_off[(size_t)l]; k < g.anchor_off[(size_t)l + 1]; k++) {
            const int id = g.anchor_prg_id[(size_t)k];
            if (id < 0 || (size_t)id >= id_loaded.size() || !id_loaded[(size_t)id]) continue;
            if (!a.count(id)) a[id] = g.anchor_pos[(size_t)k];
        }
    }
    return a;
}
}

InsertSizeEstimate estimate_insert_size(const FlatGraph& g, int64_t n_pairs, const int32_t* first_level, const int32_t* last_level, const uint8_t* reverse,
                                        const int32_t* loaded_contigs, int32_t n_loaded) {
    int max_id = 0; for (int32_t id : g.contig_prg_id) if (id > max_id) max_id = id;
    std::vector<uint8_t> id_loaded((size_t)max_id + 1, 0);
    for (int32_t i = 0; i < n_loaded; i++) { const int32_t c = loaded_contigs[i]; if (c < 0 || c >= g.n_contigs) throw std::runtime_error("insert size: contig index out of range"); id_loaded[(size_t)g.contig_prg_id[(size_t)c]] = 1; }
    InsertSizeEstimate r; std::map<int, double> hist;
    for (int64_t p = 0; p < n_pairs; p++) {
        const int f1 = first_level[2 * p], l1 = last_level[2 * p], f2 = first_level[2 * p + 1], l2 = last_level[2 * p + 1]; const bool r1 = reverse[2 * p] != 0, r2 = reverse[2 * p + 1] != 0;
        r.used++;
        bool valid = false;                                     // alignedReadPair_strandsValid
        if (f1 != -1 && f2 != -1 && r1 != r2) valid = !r1 ? (f1 < f2) : (l1 > l2);
        if (!valid) { r.skipped++; continue; }
        // the last <= 2 levels of the upstream alignment against the first <= 2 of the downstream one (levels of an alignment are consecutive)
        const bool first_up = f1 < f2; const int uf = first_up ? f1 : f2, ul = first_up ? l1 : l2, df = first_up ? f2 : f1, dl = first_up ? l2 : l1;
        int endl[2], begl[2]; int ne = 0, nb = 0;
        endl[ne++] = ul; if (ul - 1 >= uf) endl[ne++] = ul - 1;
        begl[nb++] = df; if (df + 1 <= dl) begl[nb++] = df + 1;
        const std::map<int, int> ea = anchors_of(g, endl, ne, id_loaded), ba = anchors_of(g, begl, nb, id_loaded);
        std::set<int> dist; for (const auto& kv : ea) { auto it = ba.find(kv.first); if (it != ba.end()) dist.insert(it->second - kv.second - 1); }
        for (int d : dist) { if (!hist.count(d)) hist[d] = 0; hist[d] += 1.0 / (double)dist.size(); }
    }
    // calculateInsertSizeFromHistogram
    double total = 0; for (const auto& kv : hist) total += kv.second;
    double cum = 0, med = 0, p20 = 0, p80 = 0; bool sm = false, s20 = false, s80 = false;
    for (const auto& kv : hist) {
        cum += kv.second;
        if (!sm && cum >= total * 0.5) { med = kv.first; sm = true; }
        if (!s20 && cum >= total * 0.2) { p20 = kv.first; s20 = true; }
        if (!s80 && cum >= total * 0.8) { p80 = kv.first; s80 = true; }
    }
    if (!(sm && s20 && s80)) throw std::runtime_error("insert size: no read pair of the sample yields a distance (the reference asserts set_weighted_80 && set_weighted_20 && set_median)");
    r.mean = med; const double a = std::abs(med - p20), b = std::abs(med - p80); r.sd = a > b ? a : b;
    return r;
}

} // namespace hlala
