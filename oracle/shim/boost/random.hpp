// Stand-in for Boost.Random: only feeds the reference's simulators, which the oracle never runs.
#pragma once
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
namespace random {
template <class IntType = int, class RealType = double>
class poisson_distribution : public std::poisson_distribution<IntType> {
public:
    explicit poisson_distribution(RealType mean = 1) : std::poisson_distribution<IntType>(mean) {}
};
template <class Engine, class RealType = double>
class uniform_01 {
    Engine eng_;
public:
    explicit uniform_01(Engine e) : eng_(e) {}
    RealType operator()() { return std::generate_canonical<RealType, 53>(eng_); }
};
template <class IntType = int>
class uniform_int_distribution : public std::uniform_int_distribution<IntType> {
public:
    uniform_int_distribution(IntType a, IntType b) : std::uniform_int_distribution<IntType>(a, b) {}
};
template <class RealType = double>
class normal_distribution : public std::normal_distribution<RealType> {
public:
    normal_distribution(RealType m = 0, RealType s = 1) : std::normal_distribution<RealType>(m, s) {}
};
} // random
using random::uniform_01;
template <class Engine, class Dist>
class variate_generator {
    Engine e_; Dist d_;
public:
    variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
    typename Dist::result_type operator()() { return d_(e_); }
};
typedef std::normal_distribution<double> normal_distribution_d;
}
