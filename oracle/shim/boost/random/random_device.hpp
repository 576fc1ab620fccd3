#pragma once
#include <random>
namespace boost { namespace random { typedef std::random_device random_device; } using random::random_device; }
