// Test-infrastructure stand-in for Boost.Serialization (absent from this image).
// The reference's serialize() members only need these symbols to exist; the
// oracle harness never reads or writes an archive.
#pragma once
#include <iostream>
#include <fstream>
#include <sstream>
namespace boost { namespace serialization { class access {}; } }
namespace boost { namespace archive {
class text_oarchive {
public:
    explicit text_oarchive(std::ostream&) {}
    template <class T> text_oarchive& operator<<(const T&) { return *this; }
    template <class T> text_oarchive& operator&(const T&) { return *this; }
};
} }
