#pragma once
#include <cmath>
namespace boost { namespace math {
class chi_squared { double df_; public: explicit chi_squared(double df) : df_(df) {} double degrees_of_freedom() const { return df_; } };
// only df = 1 is used by the reference (hla/HLATyper.cpp:4316): cdf = erf(sqrt(x/2))
inline double cdf(const chi_squared& d, double x) { (void)d; if (x <= 0) return 0; return std::erf(std::sqrt(x / 2.0)); }
} }
