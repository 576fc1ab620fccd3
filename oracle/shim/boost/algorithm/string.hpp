// Stand-in for the three Boost.StringAlgo entry points the reference uses
// (iter_split with first_finder, join, is_iequal).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_iequal {};
struct first_finder_shim { std::string sep; };
inline first_finder_shim first_finder(const std::string& s, is_iequal) { return first_finder_shim{s}; }
inline first_finder_shim first_finder(const std::string& s) { return first_finder_shim{s}; }
template <class Seq>
inline void iter_split(Seq& out, const std::string& in, const first_finder_shim& f) {
    out.clear();
    size_t pos = 0;
    while (true) {
        size_t hit = in.find(f.sep, pos);   // separators in graph.txt contain no letters: case-insensitive == exact
        if (hit == std::string::npos) { out.push_back(in.substr(pos)); break; }
        out.push_back(in.substr(pos, hit - pos));
        pos = hit + f.sep.size();
    }
}
namespace algorithm {
template <class Seq>
inline std::string join(const Seq& v, const std::string& sep) {
    std::string r; bool first = true;
    for (const auto& s : v) { if (!first) r += sep; r += s; first = false; }
    return r;
}
}
using algorithm::join;
}
