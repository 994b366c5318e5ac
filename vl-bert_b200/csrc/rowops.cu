// HBM-bound row kernels around the GEMMs: TF-style LayerNorm forward/backward (reference:
// external/pytorch_pretrained_bert/modeling.py:231-235 -- biased variance, eps INSIDE the sqrt, eps = 1e-12),
// column sums for bias gradients, fp32 -> bf16 casts.  One warp per row, 128-bit loads, warp-shuffle
// reductions, statistics in fp32.
#include <cstdlib>

#include "common.cuh"
#include "philox.cuh"

namespace vlb {

namespace {

constexpr int LN_WARPS = 4;  // rows per block per iteration

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: x fp32 [M, H] -> y bf16 [M, H] (+ optional fp32 copy), mean/rstd fp32 [M]
// ITERS = ceil(H / 256): each lane holds ITERS vectors of 8 values.  Warps stride over rows and prefetch the next
// row's vectors before reducing the current one, so every warp keeps 2 x ITERS x 32 B x 32 lanes in flight.
// ------------------------------------------------------------------------------------------------
template <int ITERS>
struct RowRaw {
  float4 a[ITERS], b[ITERS];
};

template <int ITERS>
__device__ __forceinline__ void load_row_f32(RowRaw<ITERS>& r, const float* __restrict__ xr, int lane, int nvec) {
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      r.a[i] = __ldg(reinterpret_cast<const float4*>(xr + vi * 8));
      r.b[i] = __ldg(reinterpret_cast<const float4*>(xr + vi * 8 + 4));
    } else {
      r.a[i] = r.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int ITERS>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y, float* __restrict__ y32, float* __restrict__ mean,
                     float* __restrict__ rstd, int M, int H, int ldx, float eps, const DropCfg drop) {
  const int lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  const int wstride = gridDim.x * LN_WARPS;
  int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();
  if (row >= M) return;
  const DropState dstate = drop_state(drop);
  RowRaw<ITERS> cur, nxt;
  load_row_f32<ITERS>(cur, x + (size_t)row * ldx, lane, nvec);
  for (; row < M; row += wstride) {
    const int nrow = row + wstride;
    if (nrow < M) load_row_f32<ITERS>(nxt, x + (size_t)nrow * ldx, lane, nvec);
    float v[ITERS][8];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      v[i][0] = cur.a[i].x; v[i][1] = cur.a[i].y; v[i][2] = cur.a[i].z; v[i][3] = cur.a[i].w;
      v[i][4] = cur.b[i].x; v[i][5] = cur.b[i].y; v[i][6] = cur.b[i].z; v[i][7] = cur.b[i].w;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float mu = warp_sum(s) / (float)H;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mu;
          q += d * d;
        }
      }
    }
    const float var = warp_sum(q) / (float)H;
    const float rs = 1.0f / sqrtf(var + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = gg[j] * ((v[i][j] - mu) * rs) + bb[j];
        if (drop.thresh != 0u) {   // dropout on the LayerNorm output (embedding, visual_linguistic_bert.py:239)
          const uint64_t g0i = ((uint64_t)row * (uint64_t)H + (uint64_t)vi * 8) >> 2;
          float lo[4] = {o[0], o[1], o[2], o[3]}, hi[4] = {o[4], o[5], o[6], o[7]};
          drop4(lo, g0i, drop, dstate);
          drop4(hi, g0i + 1, drop, dstate);
          o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = lo[3]; o[4] = hi[0]; o[5] = hi[1]; o[6] = hi[2]; o[7] = hi[3];
        }
        uint4 pk;
        pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
        pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
        if (y) *reinterpret_cast<uint4*>(y + (size_t)row * H + vi * 8) = pk;
        if (y32) {
          *reinterpret_cast<float4*>(y32 + (size_t)row * H + vi * 8) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(y32 + (size_t)row * H + vi * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    cur = nxt;
  }
}

// Lean forward for the encoder's own LayerNorms (no output dropout, bf16 output only): one warp per row and no row loop, few
// enough registers that almost every row of the token matrix has its warp resident at once (6464 rows at config 2 against
// 148 x 40 warp slots), so all of the 20 MB input is requested in the first microsecond and the kernel is one HBM burst
// instead of a chain of dependent row iterations per warp.  Same arithmetic as layernorm_fwd_kernel (two-pass variance).
constexpr int LN_LEAN_WARPS = 8;
template <int ITERS>
__global__ void __launch_bounds__(LN_LEAN_WARPS * 32, 5)
layernorm_fwd_lean_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int M, int H, int ldx,
                          float eps) {
  const int lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  const int row = blockIdx.x * LN_LEAN_WARPS + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();
  if (row >= M) return;
  const float* xr = x + (size_t)row * ldx;
  float4 a[ITERS], b[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      a[i] = __ldg(reinterpret_cast<const float4*>(xr + vi * 8));
      b[i] = __ldg(reinterpret_cast<const float4*>(xr + vi * 8 + 4));
    } else {
      a[i] = b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) s += (a[i].x + a[i].y) + (a[i].z + a[i].w) + (b[i].x + b[i].y) + (b[i].z + b[i].w);
  const float mu = warp_sum(s) / (float)H;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    if (lane + i * 32 < nvec) {
      const float d0 = a[i].x - mu, d1 = a[i].y - mu, d2 = a[i].z - mu, d3 = a[i].w - mu;
      const float d4 = b[i].x - mu, d5 = b[i].y - mu, d6 = b[i].z - mu, d7 = b[i].w - mu;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) + (d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7);
    }
  }
  const float var = warp_sum(q) / (float)H;
  const float rs = 1.0f / sqrtf(var + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      uint4 pk;
      pk.x = pack_bf16x2(g0.x * ((a[i].x - mu) * rs) + b0.x, g0.y * ((a[i].y - mu) * rs) + b0.y);
      pk.y = pack_bf16x2(g0.z * ((a[i].z - mu) * rs) + b0.z, g0.w * ((a[i].w - mu) * rs) + b0.w);
      pk.z = pack_bf16x2(g1.x * ((b[i].x - mu) * rs) + b1.x, g1.y * ((b[i].y - mu) * rs) + b1.y);
      pk.w = pack_bf16x2(g1.z * ((b[i].z - mu) * rs) + b1.z, g1.w * ((b[i].w - mu) * rs) + b1.w);
      *reinterpret_cast<uint4*>(y + (size_t)row * H + vi * 8) = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  dy = dy_bf16 (optional) + dy_f32 (optional).
//   xhat = (x - mean) * rstd ; g = dy * gamma
//   dx = rstd * (g - mean_H(g) - xhat * mean_H(g * xhat))          -> bf16 (and/or fp32)
//   dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; dbias_prev += sum_rows dx   (fp32 atomics)
// Persistent warps stride over rows, keep per-lane column partials in registers and prefetch the next row.
// ------------------------------------------------------------------------------------------------
template <int ITERS>
struct RowRawB {
  float4 xa[ITERS], xb[ITERS];
  uint4 d16[ITERS];
  float4 da[ITERS], db[ITERS];
  float mu, rs;
};

template <int ITERS, bool HAS16, bool HAS32>
__device__ __forceinline__ void load_row_bwd(RowRawB<ITERS>& r, const float* __restrict__ xr, const __nv_bfloat16* __restrict__ d16r,
                                             const float* __restrict__ d32r, const float* __restrict__ mean_p,
                                             const float* __restrict__ rstd_p, int lane, int nvec) {
  r.mu = __ldg(mean_p);
  r.rs = __ldg(rstd_p);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      r.xa[i] = __ldg(reinterpret_cast<const float4*>(xr + vi * 8));
      r.xb[i] = __ldg(reinterpret_cast<const float4*>(xr + vi * 8 + 4));
      if (HAS16) r.d16[i] = __ldg(reinterpret_cast<const uint4*>(d16r + vi * 8));
      if (HAS32) {
        r.da[i] = __ldg(reinterpret_cast<const float4*>(d32r + vi * 8));
        r.db[i] = __ldg(reinterpret_cast<const float4*>(d32r + vi * 8 + 4));
      }
    }
  }
}

template <int ITERS, bool HAS16, bool HAS32>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32,
                     const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx16, float* __restrict__ dx32,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum, int M, int H,
                     int ldx, int ld_dx, const DropCfg in_drop, __nv_bfloat16* __restrict__ dx16_drop, const DropCfg out_drop) {
  __shared__ float red[3][LN_WARPS][32 * 8 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = H >> 3;
  const int wstride = gridDim.x * LN_WARPS;
  pdl_trigger();
  pdl_wait();
  const DropState in_state = drop_state(in_drop);
  const int out_wpr = (H + 31) >> 5;
  float gam[ITERS][8];
  float acc_g[ITERS][8], acc_b[ITERS][8], acc_c[ITERS][8];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int vi = lane + i * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gam[i][j] = (vi < nvec) ? __ldg(gamma + vi * 8 + j) : 0.0f;
      acc_g[i][j] = acc_b[i][j] = acc_c[i][j] = 0.0f;
    }
  }
  int row = blockIdx.x * LN_WARPS + warp;
  RowRawB<ITERS> cur, nxt;
  if (row < M)
    load_row_bwd<ITERS, HAS16, HAS32>(cur, x + (size_t)row * ldx, dy16 + (size_t)row * H, dy32 + (size_t)row * H, mean + row,
                                      rstd + row, lane, nvec);
  for (; row < M; row += wstride) {
    const int nrow = row + wstride;
    if (nrow < M)
      load_row_bwd<ITERS, HAS16, HAS32>(nxt, x + (size_t)nrow * ldx, dy16 + (size_t)nrow * H, dy32 + (size_t)nrow * H,
                                        mean + nrow, rstd + nrow, lane, nvec);
    const float mu = cur.mu, rs = cur.rs;
    float dy[ITERS][8], xh[ITERS][8];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float xv[8] = {cur.xa[i].x, cur.xa[i].y, cur.xa[i].z, cur.xa[i].w, cur.xb[i].x, cur.xb[i].y, cur.xb[i].z, cur.xb[i].w};
        float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (HAS16) {
          const uint4 u = cur.d16[i];
          d[0] = bf16lo(u.x); d[1] = bf16hi(u.x); d[2] = bf16lo(u.y); d[3] = bf16hi(u.y);
          d[4] = bf16lo(u.z); d[5] = bf16hi(u.z); d[6] = bf16lo(u.w); d[7] = bf16hi(u.w);
        }
        if (HAS32) {
          d[0] += cur.da[i].x; d[1] += cur.da[i].y; d[2] += cur.da[i].z; d[3] += cur.da[i].w;
          d[4] += cur.db[i].x; d[5] += cur.db[i].y; d[6] += cur.db[i].z; d[7] += cur.db[i].w;
        }
        if (in_drop.thresh != 0u) {   // the LayerNorm output was dropped in forward: its gradient passes through the same mask
          const uint64_t g0i = ((uint64_t)row * (uint64_t)H + (uint64_t)vi * 8) >> 2;
          float lo[4] = {d[0], d[1], d[2], d[3]}, hi[4] = {d[4], d[5], d[6], d[7]};
          drop4(lo, g0i, in_drop, in_state);
          drop4(hi, g0i + 1, in_drop, in_state);
          d[0] = lo[0]; d[1] = lo[1]; d[2] = lo[2]; d[3] = lo[3]; d[4] = hi[0]; d[5] = hi[1]; d[6] = hi[2]; d[7] = hi[3];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dy[i][j] = d[j];
          xh[i][j] = (xv[j] - mu) * rs;
          const float g = d[j] * gam[i][j];
          s1 += g;
          s2 += g * xh[i][j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[i][j] = xh[i][j] = 0.0f;
      }
    }
    const float m1 = warp_sum(s1) / (float)H;
    const float m2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = rs * (dy[i][j] * gam[i][j] - m1 - xh[i][j] * m2);
          acc_g[i][j] += dy[i][j] * xh[i][j];
          acc_b[i][j] += dy[i][j];
        }
        const size_t off = (size_t)row * H + vi * 8;
        if (out_drop.thresh != 0u) {
          // x = dropout(dense(.)) + residual: the dense branch (next GEMM operand, bias gradient) sees dx o mask / (1-p)
          float lo[4] = {o[0], o[1], o[2], o[3]}, hi[4] = {o[4], o[5], o[6], o[7]};
          drop4_bits(lo, keep4_bits(out_drop.bits, (size_t)row, out_wpr, vi * 8), out_drop.scale);
          drop4_bits(hi, keep4_bits(out_drop.bits, (size_t)row, out_wpr, vi * 8 + 4), out_drop.scale);
          uint4 pk;
          pk.x = pack_bf16x2(lo[0], lo[1]); pk.y = pack_bf16x2(lo[2], lo[3]);
          pk.z = pack_bf16x2(hi[0], hi[1]); pk.w = pack_bf16x2(hi[2], hi[3]);
          if (dx16_drop) *reinterpret_cast<uint4*>(dx16_drop + off) = pk;
          acc_c[i][0] += lo[0]; acc_c[i][1] += lo[1]; acc_c[i][2] += lo[2]; acc_c[i][3] += lo[3];
          acc_c[i][4] += hi[0]; acc_c[i][5] += hi[1]; acc_c[i][6] += hi[2]; acc_c[i][7] += hi[3];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_c[i][j] += o[j];
        }
        if (dx16) {
          uint4 pk;
          pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
          pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(dx16 + off) = pk;
        }
        if (dx32) {
          const size_t doff = (size_t)row * ld_dx + vi * 8;
          *reinterpret_cast<float4*>(dx32 + doff) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dx32 + doff + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    cur = nxt;
  }
  // block reduction of the column partials (over the LN_WARPS warps), then one atomic per column per block
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[0][warp][lane * 8 + j] = acc_g[i][j];
      red[1][warp][lane * 8 + j] = acc_b[i][j];
      red[2][warp][lane * 8 + j] = acc_c[i][j];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 256; e += LN_WARPS * 32) {
      const int col = i * 256 + e;
      if (col < H) {
        float a = 0.0f, b = 0.0f, c = 0.0f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) {
          a += red[0][w][e];
          b += red[1][w][e];
          c += red[2][w][e];
        }
        if (dgamma) atomicAdd(dgamma + col, a);
        if (dbeta) atomicAdd(dbeta + col, b);
        if (dcolsum) atomicAdd(dcolsum + col, c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix: out[n] += sum_m x[m, n]      (bias gradients)
// block = 256 threads = 32 column-groups (8 cols each) x 8 row lanes ; grid.x over column chunks of 256,
// grid.y over row slabs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int ld, float* __restrict__ out, int M, int N, int rows_per_block) {
  __shared__ float red[8][256 + 8];
  pdl_trigger();
  pdl_wait();
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (int r = r0 + rl; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (size_t)r * ld + col);
      acc[0] += bf16lo(u.x); acc[1] += bf16hi(u.x); acc[2] += bf16lo(u.y); acc[3] += bf16hi(u.y);
      acc[4] += bf16lo(u.z); acc[5] += bf16hi(u.z); acc[6] += bf16lo(u.w); acc[7] += bf16hi(u.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int e = threadIdx.x;
  const int c = blockIdx.x * 256 + e;
  if (c < N) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][e];
    atomicAdd(out + c, s);
  }
}

// fp32 -> bf16 (weights, inputs); n multiple of 8 handled vectorised, tail scalar.
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t nv = n >> 3;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(in + i * 8);
    const float4 b = *reinterpret_cast<const float4*>(in + i * 8 + 4);
    uint4 pk;
    pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
    pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + i * 8) = pk;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = (nv << 3) + threadIdx.x;
    out[i] = __float2bfloat16(in[i]);
  }
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

// many tensors, one launch: descriptor table lives in device memory (grid.y = tensor)
__global__ void multi_cast_kernel(const VlbCastDesc* __restrict__ descs) {
  pdl_trigger();
  pdl_wait();
  const VlbCastDesc d = descs[blockIdx.y];
  const float* in = static_cast<const float*>(d.src);
  const size_t n = (size_t)d.n;
  const size_t nv = n >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (d.dst_is_bf16) {
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(d.dst);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nv; i += stride) {
      const float4 a = *reinterpret_cast<const float4*>(in + i * 8);
      const float4 b = *reinterpret_cast<const float4*>(in + i * 8 + 4);
      uint4 pk;
      pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
      pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w);
      *reinterpret_cast<uint4*>(out + i * 8) = pk;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) out[(nv << 3) + threadIdx.x] = __float2bfloat16(in[(nv << 3) + threadIdx.x]);
  } else {
    float* out = static_cast<float*>(d.dst);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
  }
}

}  // namespace

int multi_cast(const VlbCastDesc* descs_device, int count, int blocks_per_tensor, cudaStream_t stream) {
  VLB_REQUIRE(descs_device && count > 0, "multi_cast: bad arguments");
  if (blocks_per_tensor < 1) blocks_per_tensor = 16;
  VLB_CHECK_CUDA(launch_pdl(multi_cast_kernel, dim3(blocks_per_tensor, count), dim3(256), 0, stream, descs_device));
  return VLB_OK;
}

#define VLB_LN_DISPATCH(ITERS_EXPR, CALL)                     \
  switch (ITERS_EXPR) {                                       \
    case 1: { constexpr int IT = 1; CALL; } break;            \
    case 2: { constexpr int IT = 2; CALL; } break;            \
    case 3: { constexpr int IT = 3; CALL; } break;            \
    case 4: { constexpr int IT = 4; CALL; } break;            \
    case 5: case 6: { constexpr int IT = 6; CALL; } break;    \
    default: { constexpr int IT = 8; CALL; } break;           \
  }

int layernorm_forward(const float* x, int ldx, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean,
                      float* rstd, int M, int H, float eps, cudaStream_t stream, const VlbDropout* out_drop) {
  VLB_REQUIRE(drop_valid(out_drop), "layernorm_forward: bad dropout configuration");
  VLB_REQUIRE(out_drop == nullptr || out_drop->p == 0.0f || out_drop->rng != nullptr, "layernorm_forward: the output dropout is evaluated inline and needs the rng state");
  const DropCfg dcfg = make_drop(out_drop);
  VLB_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "layernorm_forward: null pointer");
  VLB_REQUIRE(ldx % 4 == 0 && ldx >= H, "layernorm_forward: bad ldx %d", ldx);
  VLB_REQUIRE(H % 8 == 0 && H >= 8 && H <= 2048, "layernorm: H=%d must be a multiple of 8 in [8, 2048]", H);
  if (M <= 0) return VLB_OK;
  const int iters = (H + 255) / 256;
  int grid = (M + LN_WARPS - 1) / LN_WARPS;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  ProfScope prof(PROF_LN_FWD, (double)M * H * (4.0 + (y_bf16 ? 2.0 : 0.0) + (y_f32 ? 4.0 : 0.0)), stream);
  cudaError_t lerr = cudaSuccess;
  static const int lean = [] { const char* v = getenv("VLB_LN_LEAN"); return v ? atoi(v) : 1; }();
  if (lean && dcfg.thresh == 0u && y_bf16 != nullptr && y_f32 == nullptr && iters <= 4) {
    const int lgrid = (M + LN_LEAN_WARPS - 1) / LN_LEAN_WARPS;
    VLB_LN_DISPATCH(iters, (lerr = launch_pdl(layernorm_fwd_lean_kernel<IT>, dim3(lgrid), dim3(LN_LEAN_WARPS * 32), 0, stream, x, gamma, beta,
                                              static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, M, H, ldx, eps)));
    VLB_CHECK_CUDA(lerr);
    return VLB_OK;
  }
  VLB_LN_DISPATCH(iters, (lerr = launch_pdl(layernorm_fwd_kernel<IT>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, x, gamma, beta,
                                            static_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd, M, H, ldx, eps, dcfg)));
  VLB_CHECK_CUDA(lerr);
  return VLB_OK;
}

int layernorm_backward(const void* dy_bf16, const float* dy_f32, const float* x, int ldx, const float* mean, const float* rstd,
                       const float* gamma, void* dx_bf16, float* dx_f32, int ld_dx, float* dgamma, float* dbeta, float* dcolsum,
                       int M, int H, cudaStream_t stream, const VlbDropout* in_drop, void* dx_bf16_drop, const VlbDropout* out_drop) {
  VLB_REQUIRE(drop_valid(in_drop) && drop_valid(out_drop), "layernorm_backward: bad dropout configuration");
  const DropCfg din = make_drop(in_drop), dout = make_drop(out_drop);
  VLB_REQUIRE(dout.thresh == 0u || dx_bf16_drop != nullptr || dcolsum != nullptr, "layernorm_backward: out_drop without a consumer");
  VLB_REQUIRE(dout.thresh == 0u || dout.bits != nullptr, "layernorm_backward: out_drop needs precomputed keep bits (vlb_dropout_bits over [M, H])");
  VLB_REQUIRE(din.thresh == 0u || din.rng != nullptr, "layernorm_backward: in_drop is evaluated inline and needs the rng state");
  VLB_REQUIRE(ldx % 4 == 0 && ldx >= H && (dx_f32 == nullptr || (ld_dx % 4 == 0 && ld_dx >= H)), "layernorm_backward: bad ld");
  VLB_REQUIRE((dy_bf16 || dy_f32) && x && mean && rstd && gamma, "layernorm_backward: null pointer");
  VLB_REQUIRE(H % 8 == 0 && H >= 8 && H <= 2048, "layernorm: H=%d must be a multiple of 8 in [8, 2048]", H);
  if (M <= 0) return VLB_OK;
  const int iters = (H + 255) / 256;
  int grid = (M + LN_WARPS - 1) / LN_WARPS;
  static const int per_sm = [] { const char* v = getenv("VLB_LN_BWD_BLOCKS_PER_SM"); return v ? atoi(v) : 2; }();
  const int cap = num_sms() * per_sm;
  if (grid > cap) grid = cap;
  ProfScope prof(PROF_LN_BWD, (double)M * H * (4.0 + (dy_bf16 ? 2.0 : 0.0) + (dy_f32 ? 4.0 : 0.0) + (dx_bf16 ? 2.0 : 0.0) + (dx_f32 ? 4.0 : 0.0) +
                                               (dx_bf16_drop ? 2.0 : 0.0)), stream);
#define VLB_LN_BWD(H16, H32)                                                                                    \
  VLB_LN_DISPATCH(iters, (lerr = launch_pdl(layernorm_bwd_kernel<IT, H16, H32>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, \
                             static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, mean, rstd, gamma,             \
                             static_cast<__nv_bfloat16*>(dx_bf16), dx_f32, dgamma, dbeta, dcolsum, M, H, ldx, ld_dx, din,   \
                             static_cast<__nv_bfloat16*>(dx_bf16_drop), dout)))
  cudaError_t lerr = cudaSuccess;
  if (dy_bf16 && dy_f32) { VLB_LN_BWD(true, true) }
  else if (dy_bf16) { VLB_LN_BWD(true, false) }
  else { VLB_LN_BWD(false, true) }
#undef VLB_LN_BWD
  VLB_CHECK_CUDA(lerr);
  return VLB_OK;
}

int colsum_bf16(const void* x, int ld, float* out, int M, int N, cudaStream_t stream) {
  VLB_REQUIRE(x && out, "colsum: null pointer");
  VLB_REQUIRE(N % 8 == 0 && ld % 8 == 0, "colsum: N and ld must be multiples of 8");
  if (M <= 0) return VLB_OK;
  const int col_blocks = (N + 255) / 256;
  int slabs = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  int rows_per = (M + slabs - 1) / slabs;
  if (rows_per < 32) rows_per = 32;
  slabs = (M + rows_per - 1) / rows_per;
  VLB_CHECK_CUDA(launch_pdl(colsum_bf16_kernel, dim3(col_blocks, slabs), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(x), ld,
                            out, M, N, rows_per));
  return VLB_OK;
}

int cast_f32_to_bf16(const float* in, void* out, size_t n, cudaStream_t stream) {
  VLB_REQUIRE(in && out, "cast: null pointer");
  if (n == 0) return VLB_OK;
  VLB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "cast: pointers must be 16-byte aligned");
  size_t blocks = ((n >> 3) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), n);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int cast_bf16_to_f32(const void* in, float* out, size_t n, cudaStream_t stream) {
  VLB_REQUIRE(in && out, "cast: null pointer");
  if (n == 0) return VLB_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  cast_bf16_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, n);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

}  // namespace vlb
