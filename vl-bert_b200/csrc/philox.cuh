// vlbert_b200 -- Philox4x32-10 counter-based RNG and the dropout-mask contract (oracle/philox.py restates both on the CPU):
//   key = (seed lo, seed hi); counter = (group lo, group hi, site, step), group = linear element index / 4;
//   element 4*group + j is KEPT iff word j >= floor(p * 2^32).
// Every fused kernel that applies dropout derives its mask from this function, so forward and backward agree without
// storing masks and the CPU oracle can reproduce them bit for bit.
#pragma once
#include <cstdint>

namespace vlb {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

// the four random words of element group `group` (elements 4*group .. 4*group+3)
__host__ __device__ __forceinline__ Philox4 dropout_words(uint64_t group, uint64_t seed, uint32_t site, uint32_t step) {
  return philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), site, step, (uint32_t)seed, (uint32_t)(seed >> 32));
}

}  // namespace vlb
