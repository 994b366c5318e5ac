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

// ---- dropout configuration handed to the kernels -----------------------------------------------------------------
// `rng` is a DEVICE pointer to {seed, step}: the kernels read it at run time, so a CUDA-graph replay sees the current
// training step without re-capturing.  thresh == 0 means "no dropout at this site".
// 2-D contract (every fused site): element (r, c) of a row-major [rows, cols] tensor belongs to
//   group = r * ceil(cols / 4) + c / 4, word = c % 4
// which is the linear contract above whenever cols % 4 == 0 (all hidden-size tensors); attention probabilities
// (rows = (b, head, query), cols = S keys) pad each row to a multiple of four so a row never shares a Philox call.
#include "../../include/vlbert_b200.h"
namespace vlb {
// `bits` (optional): keep flags precomputed by dropout_bits_kernel (vlb_dropout_bits) -- bit (c % 32) of word
// r * ceil(cols / 32) + c / 32 is the keep flag of element (r, c).  The attention / GEMM-epilogue / LayerNorm-backward
// consumers read these instead of running Philox themselves: the ten rounds per four elements made the attention kernels
// instruction-bound (+0.45 ms per step), one bit-generation launch per layer costs ~10 us and serves forward AND backward.
struct DropCfg {
  uint32_t thresh;
  float scale;
  uint32_t site;
  const uint64_t* rng;
  const uint32_t* bits;
};
inline DropCfg make_drop(const VlbDropout* d) {
  DropCfg c{0u, 1.0f, 0u, nullptr, nullptr};
  if (d != nullptr && d->p > 0.0f && (d->rng != nullptr || d->keep_bits != nullptr)) {
    c.thresh = dropout_threshold(d->p);
    c.scale = 1.0f / (1.0f - d->p);
    c.site = d->site;
    c.rng = d->rng;
    c.bits = d->keep_bits;
  }
  return c;
}
inline bool drop_valid(const VlbDropout* d) {
  return d == nullptr || (d->p >= 0.0f && d->p < 1.0f && (d->p == 0.0f || d->rng != nullptr || d->keep_bits != nullptr));
}
#if defined(__CUDACC__)
struct DropState { uint64_t seed; uint32_t step; };
__device__ __forceinline__ DropState drop_state(const DropCfg& d) {
  DropState s{0ull, 0u};
  if (d.thresh != 0u && d.rng != nullptr) { s.seed = __ldg(d.rng); s.step = (uint32_t)__ldg(d.rng + 1); }
  return s;
}
// keep flags of the four consecutive elements (r, c .. c+3), c % 4 == 0, as a 4-bit value; wpr = words per row of `bits`
__device__ __forceinline__ uint32_t keep4_bits(const uint32_t* __restrict__ bits, size_t r, int wpr, int c) {
  return (__ldg(bits + r * (size_t)wpr + (size_t)(c >> 5)) >> (c & 31)) & 0xFu;
}
__device__ __forceinline__ void drop4_bits(float (&x)[4], uint32_t k4, float scale) {
  x[0] = (k4 & 1u) ? x[0] * scale : 0.0f;
  x[1] = (k4 & 2u) ? x[1] * scale : 0.0f;
  x[2] = (k4 & 4u) ? x[2] * scale : 0.0f;
  x[3] = (k4 & 8u) ? x[3] * scale : 0.0f;
}
// apply the mask of one Philox group to four consecutive values
__device__ __forceinline__ void drop4(float (&x)[4], uint64_t group, const DropCfg& d, const DropState& s) {
  const Philox4 r = dropout_words(group, s.seed, d.site, s.step);
  x[0] = r.x >= d.thresh ? x[0] * d.scale : 0.0f;
  x[1] = r.y >= d.thresh ? x[1] * d.scale : 0.0f;
  x[2] = r.z >= d.thresh ? x[2] * d.scale : 0.0f;
  x[3] = r.w >= d.thresh ? x[3] * d.scale : 0.0f;
}
#endif
}  // namespace vlb
