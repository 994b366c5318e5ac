// vlbert_b200 -- the optimizer step that follows the hot path: multi-tensor AdamW with decoupled weight
// decay exactly as the reference's AdamW.step (common/nlp/bert/optimization.py:129-187) and the global-norm gradient clip of
// the trainer (common/trainer.py:139-147, torch.nn.utils.clip_grad_norm_), in two launches for the whole model instead of
// ~10 elementwise kernels per parameter tensor.  HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace vlb {

namespace {

// sum of squares of every gradient -> sq[0] (fp32 atomics of per-block partial sums; sq must be zero on entry)
__global__ void grad_sqnorm_kernel(const VlbAdamWTensor* __restrict__ descs, float* __restrict__ sq) {
  const VlbAdamWTensor d = descs[blockIdx.y];
  const float* __restrict__ g = d.grad;
  const size_t n = (size_t)d.n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.0f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const size_t nv = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nv; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g) + i);
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float a = g[(nv << 2) + threadIdx.x]; acc += a * a; }
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) { const float a = g[i]; acc += a * a; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_sum[8];
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_sum[w];
    if (t != 0.0f) atomicAdd(sq, t);
  }
}

// grid.x: blocks that stride over one tensor.  16 left the [30522, 768] word-embedding table (21 % of all parameters) to 16 SMs:
// 2.7 ms per step, 0.20 of the HBM roofline (BENCH r02 v4); small tensors simply finish after one iteration.
constexpr int kBlocksPerTensor = 128;

// hyper[t] = (lr, lr * weight_decay, step_size, unused) of tensor t for THIS step (step_size carries the bias correction)
struct AdamConst { float beta1, beta2, omb1, omb2, eps; };   // omb = 1 - beta, rounded from the double-precision difference like torch does

__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, const AdamConst& c, float step_size, float lr_wd) {
  // optimization.py:160-169: m <- m*b1 + (1-b1) g ; v <- v*b2 + (1-b2) g*g ; denom = sqrt(v) + eps ; p <- p - step_size * m / denom
  const float eps = c.eps;
  m = __fadd_rn(__fmul_rn(m, c.beta1), __fmul_rn(c.omb1, g));
  v = __fadd_rn(__fmul_rn(v, c.beta2), __fmul_rn(__fmul_rn(c.omb2, g), g));
  const float denom = __fadd_rn(__fsqrt_rn(v), eps);
  p = __fadd_rn(p, __fmul_rn(-step_size, __fdiv_rn(m, denom)));
  // :180-181 decoupled decay AFTER the Adam update, on the updated parameter: p <- p - lr*wd*p
  if (lr_wd > 0.0f) p = __fadd_rn(p, __fmul_rn(-lr_wd, p));
  return p;
}

__global__ void adamw_kernel(const VlbAdamWTensor* __restrict__ descs, const float4* __restrict__ hyper, AdamConst c,
                             const float* __restrict__ sq, float max_norm) {
  const VlbAdamWTensor d = descs[blockIdx.y];
  const float4 h = hyper[blockIdx.y];
  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), applied only when < 1
  float coef = 1.0f;
  if (sq != nullptr && max_norm > 0.0f) {
    const float c = max_norm / (sqrtf(*sq) + 1e-6f);
    coef = c < 1.0f ? c : 1.0f;
  }
  const size_t n = (size_t)d.n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float* __restrict__ p = d.param;
  const float* __restrict__ g = d.grad;
  float* __restrict__ m = d.exp_avg;
  float* __restrict__ v = d.exp_avg_sq;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const size_t nv = vec ? (n >> 2) : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_one(pp.x, gg.x * coef, mm.x, vv.x, c, h.z, h.y);
    adamw_one(pp.y, gg.y * coef, mm.y, vv.y, c, h.z, h.y);
    adamw_one(pp.z, gg.z * coef, mm.z, vv.z, c, h.z, h.y);
    adamw_one(pp.w, gg.w * coef, mm.w, vv.w, c, h.z, h.y);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (size_t i = (nv << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float pp = p[i], mm = m[i], vv = v[i];
    adamw_one(pp, g[i] * coef, mm, vv, c, h.z, h.y);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace

int grad_sqnorm(const VlbAdamWTensor* descs_device, int count, float* sq, cudaStream_t stream) {
  VLB_REQUIRE(descs_device && sq && count > 0, "grad_sqnorm: bad arguments");
  VLB_CHECK_CUDA(cudaMemsetAsync(sq, 0, sizeof(float), stream));
  grad_sqnorm_kernel<<<dim3(kBlocksPerTensor, count), 256, 0, stream>>>(descs_device, sq);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int adamw_step(const VlbAdamWTensor* descs_device, const float* hyper_device, int count, double beta1, double beta2, double eps,
               const float* sq, float max_norm, cudaStream_t stream) {
  VLB_REQUIRE(descs_device && hyper_device && count > 0, "adamw_step: bad arguments");
  VLB_REQUIRE((reinterpret_cast<uintptr_t>(hyper_device) & 15) == 0, "adamw_step: hyper table must be 16-byte aligned");
  VLB_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "adamw_step: bad hyper-parameters");
  const AdamConst c{(float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps};
  adamw_kernel<<<dim3(kBlocksPerTensor, count), 256, 0, stream>>>(descs_device, reinterpret_cast<const float4*>(hyper_device), c, sq, max_norm);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

}  // namespace vlb

// ---- dropout contract on the device (csrc/philox.cuh): the mask itself, and a stand-alone bf16/f32 dropout -------------
#include "philox.cuh"

namespace vlb {
namespace {

__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, size_t n, uint32_t thresh, uint64_t seed, uint32_t site, uint32_t step) {
  const size_t groups = (n + 3) >> 2;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const Philox4 r = dropout_words(g, seed, site, step);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (g * 4 + j < n) keep[g * 4 + j] = w[j] >= thresh ? 1 : 0;
  }
}

template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, size_t n, uint32_t thresh, float scale, uint64_t seed,
                               uint32_t site, uint32_t step) {
  const size_t groups = (n + 3) >> 2;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const Philox4 r = dropout_words(g, seed, site, step);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = g * 4 + j;
      if (i < n) y[i] = w[j] >= thresh ? (T)((float)x[i] * scale) : (T)0.0f;
    }
  }
}

__global__ void dropout_mask_2d_kernel(uint8_t* __restrict__ keep, size_t rows, int cols, uint32_t thresh, uint64_t seed, uint32_t site,
                                       uint32_t step) {
  const size_t gpr = (size_t)((cols + 3) >> 2);
  const size_t groups = rows * gpr;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const size_t r = g / gpr;
    const int c0 = (int)(g - r * gpr) * 4;
    const Philox4 w4 = dropout_words(g, seed, site, step);
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c0 + j < cols) keep[r * (size_t)cols + c0 + j] = w[j] >= thresh ? 1 : 0;
  }
}

// y[r, c] = x[r, c] * keep(r, col_offset + c) / (1 - p), keep indexed over a [rows, total_cols] tensor (2-D contract);
// (seed, step) come from device memory.  cols, col_offset, total_cols multiples of 4.
template <typename T>
__global__ void dropout_2d_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, size_t rows, int cols, int col_offset,
                                  int total_cols, const DropCfg drop) {
  const DropState st = drop_state(drop);
  const size_t gpr = (size_t)(cols >> 2), tg = (size_t)(total_cols >> 2);
  const size_t groups = rows * gpr;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const size_t r = g / gpr;
    const int c = (int)(g - r * gpr) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (float)x[r * (size_t)ldx + c + j];
    if (drop.thresh != 0u) drop4(v, r * tg + (size_t)((col_offset + c) >> 2), drop, st);
#pragma unroll
    for (int j = 0; j < 4; ++j) y[r * (size_t)ldy + c + j] = (T)v[j];
  }
}

// keep flags as bits: one thread = one 32-bit word = eight Philox groups of row r (2-D contract)
__global__ void dropout_bits_kernel(uint32_t* __restrict__ bits, size_t rows, int cols, int wpr, const DropCfg drop) {
  pdl_trigger();
  pdl_wait();     // (seed, step) are written by the caller's preceding stream work
  const DropState st = drop_state(drop);
  const size_t gpr = (size_t)((cols + 3) >> 2);
  const size_t words = rows * (size_t)wpr;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) {
    const size_t r = w / (size_t)wpr;
    const int c0 = (int)(w - r * (size_t)wpr) * 32;
    uint32_t out = 0u;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int c = c0 + g * 4;
      if (c < cols) {
        const Philox4 q = dropout_words(r * gpr + (size_t)(c >> 2), st.seed, drop.site, st.step);
        uint32_t k = (q.x >= drop.thresh ? 1u : 0u) | (q.y >= drop.thresh ? 2u : 0u) | (q.z >= drop.thresh ? 4u : 0u) | (q.w >= drop.thresh ? 8u : 0u);
        if (c + 4 > cols) k &= (1u << (cols - c)) - 1u;
        out |= k << (4 * g);
      }
    }
    bits[w] = out;
  }
}

// up to three mask tensors in one launch (the three dropout sites of a BertLayer): blockIdx.y selects the job
struct BitsJob { uint32_t* bits; size_t rows; int cols; int wpr; uint32_t site; uint32_t thresh; };
struct BitsJobs { BitsJob j[3]; const uint64_t* rng; };
__global__ void dropout_bits_multi_kernel(const BitsJobs jobs) {
  pdl_trigger();
  pdl_wait();
  const BitsJob jb = jobs.j[blockIdx.y];
  if (jb.bits == nullptr) return;
  DropCfg d{jb.thresh, 1.0f, jb.site, jobs.rng, nullptr};
  const DropState st = drop_state(d);
  const size_t gpr = (size_t)((jb.cols + 3) >> 2);
  const size_t words = jb.rows * (size_t)jb.wpr;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) {
    const size_t r = w / (size_t)jb.wpr;
    const int c0 = (int)(w - r * (size_t)jb.wpr) * 32;
    uint32_t out = 0u;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int c = c0 + g * 4;
      if (c < jb.cols) {
        const Philox4 q = dropout_words(r * gpr + (size_t)(c >> 2), st.seed, jb.site, st.step);
        uint32_t k = (q.x >= jb.thresh ? 1u : 0u) | (q.y >= jb.thresh ? 2u : 0u) | (q.z >= jb.thresh ? 4u : 0u) | (q.w >= jb.thresh ? 8u : 0u);
        if (c + 4 > jb.cols) k &= (1u << (jb.cols - c)) - 1u;
        out |= k << (4 * g);
      }
    }
    jb.bits[w] = out;
  }
}

int dropout_grid(size_t n) {
  const size_t groups = (n + 3) >> 2;
  size_t g = (groups + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint32_t site, uint32_t step, cudaStream_t stream) {
  VLB_REQUIRE(keep && n >= 0 && p >= 0.0f && p < 1.0f, "dropout_mask: bad arguments");
  if (n == 0) return VLB_OK;
  dropout_mask_kernel<<<dropout_grid((size_t)n), 256, 0, stream>>>(keep, (size_t)n, dropout_threshold(p), seed, site, step);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int dropout_mask_2d(uint8_t* keep, int64_t rows, int cols, float p, uint64_t seed, uint32_t site, uint32_t step, cudaStream_t stream) {
  VLB_REQUIRE(keep && rows >= 0 && cols > 0 && p >= 0.0f && p < 1.0f, "dropout_mask_2d: bad arguments");
  if (rows == 0) return VLB_OK;
  const size_t n = (size_t)rows * (size_t)(((cols + 3) >> 2) << 2);
  dropout_mask_2d_kernel<<<dropout_grid(n), 256, 0, stream>>>(keep, (size_t)rows, cols, dropout_threshold(p), seed, site, step);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int64_t dropout_bits_words(int64_t rows, int cols) { return rows * (int64_t)((cols + 31) >> 5); }

int dropout_bits(uint32_t* bits, int64_t rows, int cols, const VlbDropout* drop, cudaStream_t stream) {
  VLB_REQUIRE(bits && rows >= 0 && cols > 0 && drop && drop->p > 0.0f && drop->p < 1.0f && drop->rng, "dropout_bits: bad arguments");
  if (rows == 0) return VLB_OK;
  const int wpr = (cols + 31) >> 5;
  const size_t words = (size_t)rows * wpr;
  size_t g = (words + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  DropCfg d = make_drop(drop);
  d.bits = nullptr;
  VLB_CHECK_CUDA(launch_pdl(dropout_bits_kernel, dim3((unsigned)(g > cap ? cap : g)), dim3(256), 0, stream, bits, (size_t)rows, cols, wpr, d));
  return VLB_OK;
}

// the three sites of one BertLayer in ONE launch: attention probabilities [B*heads*S, S] and the two dense outputs [M, H]
int layer_dropout_bits(uint32_t* keep_attn, uint32_t* keep_self_out, uint32_t* keep_out, int B, int S, int H, int heads,
                       const VlbLayerDropout& d, cudaStream_t stream) {
  BitsJobs jobs{};
  jobs.rng = d.rng;
  const size_t M = (size_t)B * S;
  if (d.p_attn > 0.0f && keep_attn)
    jobs.j[0] = BitsJob{keep_attn, (size_t)B * heads * S, S, (S + 31) >> 5, d.site_attn, dropout_threshold(d.p_attn)};
  if (d.p_hidden > 0.0f && keep_self_out)
    jobs.j[1] = BitsJob{keep_self_out, M, H, (H + 31) >> 5, d.site_self_out, dropout_threshold(d.p_hidden)};
  if (d.p_hidden > 0.0f && keep_out)
    jobs.j[2] = BitsJob{keep_out, M, H, (H + 31) >> 5, d.site_out, dropout_threshold(d.p_hidden)};
  size_t words = 0;
  for (int i = 0; i < 3; ++i) { const size_t w = jobs.j[i].rows * (size_t)jobs.j[i].wpr; if (w > words) words = w; }
  if (words == 0) return VLB_OK;
  size_t g = (words + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  VLB_CHECK_CUDA(launch_pdl(dropout_bits_multi_kernel, dim3((unsigned)(g > cap ? cap : g), 3), dim3(256), 0, stream, jobs));
  return VLB_OK;
}

int dropout_2d(const void* x, int ldx, void* y, int ldy, int64_t rows, int cols, int col_offset, int total_cols, int is_bf16,
               const VlbDropout* drop, cudaStream_t stream) {
  VLB_REQUIRE(x && y && rows >= 0 && cols > 0 && cols % 4 == 0 && col_offset % 4 == 0 && total_cols % 4 == 0 &&
              col_offset + cols <= total_cols && ldx >= cols && ldy >= cols, "dropout_2d: bad arguments");
  VLB_REQUIRE(drop_valid(drop), "dropout_2d: bad dropout configuration");
  if (rows == 0) return VLB_OK;
  const DropCfg d = make_drop(drop);
  const int grid = dropout_grid((size_t)rows * cols);
  if (is_bf16)
    dropout_2d_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy,
                                                               (size_t)rows, cols, col_offset, total_cols, d);
  else
    dropout_2d_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), ldx, static_cast<float*>(y), ldy, (size_t)rows, cols,
                                                       col_offset, total_cols, d);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int dropout_apply(const void* x, void* y, int64_t n, int is_bf16, float p, uint64_t seed, uint32_t site, uint32_t step, cudaStream_t stream) {
  VLB_REQUIRE(x && y && n >= 0 && p >= 0.0f && p < 1.0f, "dropout: bad arguments");
  if (n == 0) return VLB_OK;
  const float scale = 1.0f / (1.0f - p);
  if (is_bf16)
    dropout_kernel<__nv_bfloat16><<<dropout_grid((size_t)n), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
                                                                             (size_t)n, dropout_threshold(p), scale, seed, site, step);
  else
    dropout_kernel<float><<<dropout_grid((size_t)n), 256, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(y), (size_t)n,
                                                                     dropout_threshold(p), scale, seed, site, step);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

}  // namespace vlb
