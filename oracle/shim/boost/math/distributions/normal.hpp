// Stand-in for boost::math::normal: closed forms as in Boost.Math 1.5x normal.hpp
// (pdf: exp(-(x-mean)^2 / (2 sd^2)) / (sd * sqrt(2 pi))). Parity note: unpinned but analytic.
#pragma once
#include <cmath>
namespace boost { namespace math {
class normal {
    double m_, s_;
public:
    normal(double mean = 0, double sd = 1) : m_(mean), s_(sd) {}
    double mean() const { return m_; }
    double standard_deviation() const { return s_; }
};
inline double pdf(const normal& d, double x) {
    double sd = d.standard_deviation();
    double mean = d.mean();
    if (std::isinf(x)) return 0;
    double exponent = x - mean;
    exponent *= -exponent;
    exponent /= 2 * sd * sd;
    double result = std::exp(exponent);
    result /= sd * std::sqrt(2 * 3.14159265358979323846264338327950288);
    return result;
}
} }
