#pragma once
#include <string>
#include <sstream>
#include <stdexcept>
#include <typeinfo>
namespace boost {
class bad_lexical_cast : public std::bad_cast { public: const char* what() const noexcept override { return "bad lexical cast"; } };
namespace detail_shim {
template <class T> struct lc {
    static T from(const std::string& s) {
        std::istringstream is(s); is.unsetf(std::ios::skipws); T v; is >> v;
        if (is.fail() || is.get() != std::char_traits<char>::eof()) throw bad_lexical_cast();
        return v;
    }
};
template <> struct lc<std::string> { static std::string from(const std::string& s) { return s; } };
template <> struct lc<unsigned char> { static unsigned char from(const std::string& s) { if (s.size() != 1) throw bad_lexical_cast(); return (unsigned char)s[0]; } };
template <> struct lc<char> { static char from(const std::string& s) { if (s.size() != 1) throw bad_lexical_cast(); return s[0]; } };
template <> struct lc<bool> { static bool from(const std::string& s) { if (s == "1") return true; if (s == "0") return false; throw bad_lexical_cast(); } };
template <class S> inline std::string to_str(const S& v) { std::ostringstream os; os << v; return os.str(); }
inline std::string to_str(const std::string& v) { return v; }
inline std::string to_str(const char* v) { return std::string(v); }
}
template <class T, class S> inline T lexical_cast(const S& v) { return detail_shim::lc<T>::from(detail_shim::to_str(v)); }
}
