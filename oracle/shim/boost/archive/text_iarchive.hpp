#pragma once
#include <iostream>
#include <fstream>
#include <sstream>
#include <stdexcept>
namespace boost { namespace serialization { class access; } }
namespace boost { namespace archive {
class text_iarchive {
public:
    explicit text_iarchive(std::istream&) {}
    template <class T> text_iarchive& operator>>(T&) { throw std::runtime_error("shim: boost archives are not readable in the oracle harness"); }
    template <class T> text_iarchive& operator&(T&) { return *this; }
};
} }
