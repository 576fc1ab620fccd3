// Hash of a k-mer (raw emission bytes), shared by the host index builder and the seeding kernel.
#pragma once
#include <cstdint>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif
namespace hlala {
__host__ __device__ inline uint64_t kmer_hash(const uint8_t* s, int k) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < k; i++) h = (h ^ s[i]) * 0x100000001B3ull + 0x632BE59BD9B4E019ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}
} // namespace hlala
