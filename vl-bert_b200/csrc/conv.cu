// Convolution front end of the region-feature path (ResNet-101 C4 backbone + res5 RoI head; reference:
// common/backbone/resnet/resnet.py:74-118 Bottleneck, :175-186 stem/stages; common/fast_rcnn.py:74-86 head) in NHWC bf16.
//
// A convolution is lowered onto the tcgen05 GEMM of gemm_sm100.cu:
//   forward   y[P, Cout]   = col[P, K] W[Cout, K]^T          (NT;  K = kh*kw*Cin, tap-major / channel-minor)
//   dgrad     dcol[P, K]   = dy[P, Cout] W[Cout, K]          (NN)  -> col2im
//   wgrad     dW[Cout, K]  = dy[P, Cout]^T col[P, K]         (TN)
// with frozen-BatchNorm scale/shift (resnet.py frozen_bn / common/fast_rcnn.py:88-92), ReLU and the residual add fused in
// the GEMM epilogue.  1x1 stride-1 convolutions need no lowering at all (col == x).  This file holds the HBM-bound
// data-movement kernels around those GEMMs: im2col / col2im (explicit lowering; the implicit-GEMM variant that lets TMA
// gather the taps is the next step), the elementwise ReLU/BN backward, 3x3/2 max-pool, the 14x14 average pool of the RoI
// head, NCHW fp32 <-> NHWC bf16 layout changes, and an NHWC bf16 RoIAlign (lanes over channels, 128-bit accesses).
#include "common.cuh"
#include "gemm_sm100.cuh"

namespace vlb {

namespace {

// col[(n,ho,wo), (r,s,c)] = x[n, ho*stride - pad + r*dil, wo*stride - pad + s*dil, c]   (0 outside) ; 8 channels per thread
__global__ void im2col_v8_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, ConvGeom g) {
  const int c8n = g.C >> 3;
  const int taps = g.kh * g.kw;
  const long total = (long)g.N * g.Ho * g.Wo * taps * c8n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    long t = i / c8n;
    const int tap = (int)(t % taps);
    t /= taps;
    const int wo = (int)(t % g.Wo);
    t /= g.Wo;
    const int ho = (int)(t % g.Ho);
    const int n = (int)(t / g.Ho);
    const int r = tap / g.kw, s = tap - r * g.kw;
    const int h = ho * g.stride - g.pad + r * g.dil, w = wo * g.stride - g.pad + s * g.dil;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (h >= 0 && h < g.H && w >= 0 && w < g.W)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * g.H + h) * g.W + w) * g.C + c8 * 8));
    const size_t p = ((size_t)n * g.Ho + ho) * g.Wo + wo;
    *reinterpret_cast<uint4*>(col + p * g.Kp + (size_t)tap * g.C + c8 * 8) = v;
  }
}

// generic (any C, e.g. the 3-channel stem): one thread per 8 consecutive columns of one row (one 16-byte store), scalar
// reads of x (small and cache-resident: neighbouring rows re-read the same pixels); also zero-fills the K padding
__global__ void im2col_scalar_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, ConvGeom g) {
  const int k8n = g.Kp >> 3;
  const long total = (long)g.N * g.Ho * g.Wo * k8n;
  const int K = g.kh * g.kw * g.C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k8 = (int)(i % k8n);
    const long p = i / k8n;
    const int wo = (int)(p % g.Wo);
    const long t = p / g.Wo;
    const int ho = (int)(t % g.Ho), n = (int)(t / g.Ho);
    const int h0 = ho * g.stride - g.pad, w0 = wo * g.stride - g.pad;
    int k = k8 * 8;
    int tap = k / g.C, c = k - tap * g.C;
    int r = tap / g.kw, sx = tap - r * g.kw;
    const __nv_bfloat16* xn = x + (size_t)n * g.H * g.W * g.C;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j, ++k) {
      const int h = h0 + r * g.dil, w = w0 + sx * g.dil;
      v[j] = (k < K && h >= 0 && h < g.H && w >= 0 && w < g.W) ? xn[((size_t)h * g.W + w) * g.C + c] : __float2bfloat16(0.0f);
      if (++c == g.C) { c = 0; if (++sx == g.kw) { sx = 0; ++r; } }
    }
    *reinterpret_cast<uint4*>(col + p * g.Kp + k8 * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// gather form of the im2col adjoint: dx[n,h,w,c] = sum over taps (r,s) of dcol[(n,ho,wo),(r,s,c)] with
// ho*stride - pad + r*dil == h (and the same for w).  No atomics; fp32 accumulation; optional "+ add" (residual gradient).
__global__ void col2im_v8_kernel(const __nv_bfloat16* __restrict__ dcol, const __nv_bfloat16* __restrict__ add,
                                 __nv_bfloat16* __restrict__ dx, ConvGeom g) {
  const int c8n = g.C >> 3;
  const long total = (long)g.N * g.H * g.W * c8n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    long t = i / c8n;
    const int w = (int)(t % g.W);
    t /= g.W;
    const int h = (int)(t % g.H);
    const int n = (int)(t / g.H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < g.kh; ++r) {
      const int hn = h + g.pad - r * g.dil;
      if (hn < 0 || hn % g.stride) continue;
      const int ho = hn / g.stride;
      if (ho >= g.Ho) continue;
      for (int s = 0; s < g.kw; ++s) {
        const int wn = w + g.pad - s * g.dil;
        if (wn < 0 || wn % g.stride) continue;
        const int wo = wn / g.stride;
        if (wo >= g.Wo) continue;
        const size_t p = ((size_t)n * g.Ho + ho) * g.Wo + wo;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dcol + p * g.Kp + (size_t)(r * g.kw + s) * g.C + c8 * 8));
        acc[0] += bf16lo(u.x); acc[1] += bf16hi(u.x); acc[2] += bf16lo(u.y); acc[3] += bf16hi(u.y);
        acc[4] += bf16lo(u.z); acc[5] += bf16hi(u.z); acc[6] += bf16lo(u.w); acc[7] += bf16hi(u.w);
      }
    }
    const size_t o = (((size_t)n * g.H + h) * g.W + w) * g.C + c8 * 8;
    if (add != nullptr) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(add + o));
      acc[0] += bf16lo(u.x); acc[1] += bf16hi(u.x); acc[2] += bf16lo(u.y); acc[3] += bf16hi(u.y);
      acc[4] += bf16lo(u.z); acc[5] += bf16hi(u.z); acc[6] += bf16lo(u.w); acc[7] += bf16hi(u.w);
    }
    uint4 pk;
    pk.x = pack_bf16x2(acc[0], acc[1]); pk.y = pack_bf16x2(acc[2], acc[3]);
    pk.z = pack_bf16x2(acc[4], acc[5]); pk.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dx + o) = pk;
  }
}

// Backward through "y = relu(...)" followed by the frozen-BN scale of the producing convolution:
//   d_pre = dy * [y > 0]      (optional output, the gradient that also flows into the identity branch)
//   d_conv = d_pre * scale[c] (optional output, gradient wrt the convolution's raw output)
// dy may be the sum of two bf16 tensors (dy + dy2).  mask == nullptr: no ReLU (plain scale).
__global__ void relu_bn_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                                   const __nv_bfloat16* __restrict__ y_mask, const float* __restrict__ scale,
                                   __nv_bfloat16* __restrict__ d_pre, __nv_bfloat16* __restrict__ d_conv, long rows, int C) {
  const int c8n = C >> 3;
  const long total = rows * c8n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const size_t o = (size_t)i * 8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + o));
    float v[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
    if (dy2 != nullptr) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(dy2 + o));
      v[0] += bf16lo(q.x); v[1] += bf16hi(q.x); v[2] += bf16lo(q.y); v[3] += bf16hi(q.y);
      v[4] += bf16lo(q.z); v[5] += bf16hi(q.z); v[6] += bf16lo(q.w); v[7] += bf16hi(q.w);
    }
    if (y_mask != nullptr) {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(y_mask + o));
      const float mm[8] = {bf16lo(m.x), bf16hi(m.x), bf16lo(m.y), bf16hi(m.y), bf16lo(m.z), bf16hi(m.z), bf16lo(m.w), bf16hi(m.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = mm[j] > 0.0f ? v[j] : 0.0f;
    }
    if (d_pre != nullptr) {
      uint4 pk;
      pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
      pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(d_pre + o) = pk;
    }
    if (d_conv != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c8 * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + c8 * 8 + 4));
      uint4 pk;
      pk.x = pack_bf16x2(v[0] * s0.x, v[1] * s0.y); pk.y = pack_bf16x2(v[2] * s0.z, v[3] * s0.w);
      pk.z = pack_bf16x2(v[4] * s1.x, v[5] * s1.y); pk.w = pack_bf16x2(v[6] * s1.z, v[7] * s1.w);
      *reinterpret_cast<uint4*>(d_conv + o) = pk;
    }
  }
}

// 3x3 stride-2 pad-1 max pooling, NHWC bf16 (forward only: the stem is frozen, IMAGE_FROZEN_BACKBONE_STAGES [1,2])
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H, int W,
                                    int C, int Ho, int Wo) {
  const int c8n = C >> 3;
  const long total = (long)N * Ho * Wo * c8n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    long t = i / c8n;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = wo * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + h) * W + w) * C + c8 * 8));
        const float v[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    }
    uint4 pk;
    pk.x = pack_bf16x2(m[0], m[1]); pk.y = pack_bf16x2(m[2], m[3]);
    pk.z = pack_bf16x2(m[4], m[5]); pk.w = pack_bf16x2(m[6], m[7]);
    *reinterpret_cast<uint4*>(y + (((size_t)n * Ho + ho) * Wo + wo) * C + c8 * 8) = pk;
  }
}

// mean over the HW positions of each sample: x bf16 [K, HW, C] -> y f32 [K, C]   (AvgPool2d(14) + Flattener, fast_rcnn.py:80-84)
__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int K, int HW, int C) {
  const int k = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f;
    for (int p = 0; p < HW; ++p) s += __bfloat162float(x[((size_t)k * HW + p) * C + c]);
    y[(size_t)k * C + c] = s / (float)HW;
  }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int K, int HW, int C) {
  const long total = (long)K * HW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int k = (int)(i / ((long)HW * C));
    dx[i] = __float2bfloat16(dy[(size_t)k * C + c] / (float)HW);
  }
}

// layout changes: NCHW fp32 -> NHWC bf16 (images) and NHWC bf16 -> NCHW fp32 (features for the NCHW RoIAlign ABI)
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H, int W) {
  const long total = (long)N * C * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    y[i] = __float2bfloat16(x[(((size_t)n * C + c) * H + h) * W + w]);
  }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W) {
  const long total = (long)N * C * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long t = i / W;
    const int h = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    y[i] = __bfloat162float(x[(((size_t)n * H + h) * W + w) * C + c]);
  }
}

// ------------------------------------------------------------------------------------------------
// RoIAlign on an NHWC bf16 feature map: one warp per (roi, bin), lanes over channels (8 per lane per step, 128-bit loads),
// same sampling conventions as roi_align.cu (ROIAlign_cuda.cu:15-122).  Output NHWC bf16 [K, ph, pw, C].
// ------------------------------------------------------------------------------------------------
struct BinSample { int i0, i1, i2, i3; float w0, w1, w2, w3; };
__device__ __forceinline__ bool bin_sample(int height, int width, float y, float x, BinSample& s) {
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) return false;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = y - (float)y_low, lx = x - (float)x_low, hy = 1.0f - ly, hx = 1.0f - lx;
  s.w0 = hy * hx; s.w1 = hy * lx; s.w2 = ly * hx; s.w3 = ly * lx;
  s.i0 = y_low * width + x_low; s.i1 = y_low * width + x_high; s.i2 = y_high * width + x_low; s.i3 = y_high * width + x_high;
  return true;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(128)
roi_align_nhwc_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ rois, __nv_bfloat16* __restrict__ out,
                      const __nv_bfloat16* __restrict__ grad_out, float* __restrict__ grad_feat, int K, int C, int H, int W,
                      int ph_n, int pw_n, float scale, int sampling_ratio) {
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);  // (roi, bin)
  const int lane = threadIdx.x & 31;
  const int nbins = ph_n * pw_n;
  if (gw >= K * nbins) return;
  const int n = gw / nbins, bin = gw - n * nbins;
  const int ph = bin / pw_n, pw = bin - ph * pw_n;
  const float* roi = rois + 5 * n;
  const int b = (int)roi[0];
  const float sw = roi[1] * scale, sh = roi[2] * scale, ew = roi[3] * scale, eh = roi[4] * scale;
  const float rw = fmaxf(ew - sw, 1.0f), rh = fmaxf(eh - sh, 1.0f);
  const float bh = rh / (float)ph_n, bw = rw / (float)pw_n;
  const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph_n);
  const int gwd = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw_n);
  const float count = (float)(gh * gwd);
  const size_t img = (size_t)b * H * W;
  for (int c0 = lane * 8; c0 < C; c0 += 256) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float g[8];
    if (BACKWARD) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(grad_out + ((size_t)n * nbins + bin) * C + c0));
      g[0] = bf16lo(u.x); g[1] = bf16hi(u.x); g[2] = bf16lo(u.y); g[3] = bf16hi(u.y);
      g[4] = bf16lo(u.z); g[5] = bf16hi(u.z); g[6] = bf16lo(u.w); g[7] = bf16hi(u.w);
    }
    for (int iy = 0; iy < gh; ++iy) {
      const float y = sh + ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
      for (int ix = 0; ix < gwd; ++ix) {
        const float x = sw + pw * bw + ((float)ix + 0.5f) * bw / (float)gwd;
        BinSample s;
        if (!bin_sample(H, W, y, x, s)) continue;
        const int idx[4] = {s.i0, s.i1, s.i2, s.i3};
        const float wt[4] = {s.w0, s.w1, s.w2, s.w3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!BACKWARD) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(feat + (img + idx[q]) * C + c0));
            acc[0] += wt[q] * bf16lo(u.x); acc[1] += wt[q] * bf16hi(u.x); acc[2] += wt[q] * bf16lo(u.y); acc[3] += wt[q] * bf16hi(u.y);
            acc[4] += wt[q] * bf16lo(u.z); acc[5] += wt[q] * bf16hi(u.z); acc[6] += wt[q] * bf16lo(u.w); acc[7] += wt[q] * bf16hi(u.w);
          } else {
            float* dst = grad_feat + (img + idx[q]) * C + c0;
            const float f = wt[q] / count;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g[0] * f), "f"(g[1] * f), "f"(g[2] * f), "f"(g[3] * f) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(g[4] * f), "f"(g[5] * f), "f"(g[6] * f), "f"(g[7] * f) : "memory");
          }
        }
      }
    }
    if (!BACKWARD) {
      uint4 pk;
      pk.x = pack_bf16x2(acc[0] / count, acc[1] / count); pk.y = pack_bf16x2(acc[2] / count, acc[3] / count);
      pk.z = pack_bf16x2(acc[4] / count, acc[5] / count); pk.w = pack_bf16x2(acc[6] / count, acc[7] / count);
      *reinterpret_cast<uint4*>(out + ((size_t)n * nbins + bin) * C + c0) = pk;
    }
  }
}

inline int grid_for(long total, int block) {
  long g = (total + block - 1) / block;
  const long cap = (long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int im2col_nhwc(const void* x, void* col, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int dil, int Ho, int Wo,
                int Kp, cudaStream_t stream) {
  VLB_REQUIRE(x && col, "im2col: null pointer");
  VLB_REQUIRE(Kp >= kh * kw * C && Kp % 8 == 0, "im2col: bad padded K %d", Kp);
  ConvGeom g{N, H, W, C, Ho, Wo, kh, kw, stride, pad, dil, Kp};
  ProfScope prof(PROF_IM2COL, 2.0 * N * Ho * Wo * Kp + 2.0 * N * H * W * C, stream);   // bytes: col written + x read once
  if (C % 8 == 0 && Kp == kh * kw * C) {
    const long total = (long)N * Ho * Wo * kh * kw * (C / 8);
    im2col_v8_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(col), g);
  } else {
    const long total = (long)N * Ho * Wo * (Kp / 8);
    im2col_scalar_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(col), g);
  }
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int col2im_nhwc(const void* dcol, const void* add, void* dx, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int dil,
                int Ho, int Wo, int Kp, cudaStream_t stream) {
  VLB_REQUIRE(dcol && dx, "col2im: null pointer");
  VLB_REQUIRE(C % 8 == 0 && Kp == kh * kw * C, "col2im: channels must be a multiple of 8");
  ConvGeom g{N, H, W, C, Ho, Wo, kh, kw, stride, pad, dil, Kp};
  ProfScope prof(PROF_COL2IM, 2.0 * N * Ho * Wo * Kp + 2.0 * N * H * W * C * (add ? 2 : 1), stream);
  const long total = (long)N * H * W * (C / 8);
  col2im_v8_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dcol), static_cast<const __nv_bfloat16*>(add),
                                                            static_cast<__nv_bfloat16*>(dx), g);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int relu_bn_backward(const void* dy, const void* dy2, const void* y_mask, const float* scale, void* d_pre, void* d_conv, int64_t rows,
                     int C, cudaStream_t stream) {
  VLB_REQUIRE(dy && (d_pre || d_conv) && (d_conv == nullptr || scale != nullptr), "relu_bn_backward: null pointer");
  VLB_REQUIRE(C % 8 == 0, "relu_bn_backward: channels must be a multiple of 8");
  const long total = rows * (C / 8);
  if (total <= 0) return VLB_OK;
  ProfScope prof(PROF_CONV_ELT, 2.0 * rows * C * (1 + (dy2 ? 1 : 0) + (y_mask ? 1 : 0) + (d_pre ? 1 : 0) + (d_conv ? 1 : 0)), stream);
  relu_bn_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(dy2),
                                                              static_cast<const __nv_bfloat16*>(y_mask), scale,
                                                              static_cast<__nv_bfloat16*>(d_pre), static_cast<__nv_bfloat16*>(d_conv), rows, C);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int maxpool3x3s2_nhwc(const void* x, void* y, int N, int H, int W, int C, cudaStream_t stream) {
  VLB_REQUIRE(x && y && C % 8 == 0, "maxpool: bad arguments");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long total = (long)N * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, H, W, C, Ho, Wo);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int avgpool_forward(const void* x, float* y, int K, int HW, int C, cudaStream_t stream) {
  VLB_REQUIRE(x && y, "avgpool: null pointer");
  if (K == 0) return VLB_OK;
  avgpool_fwd_kernel<<<K, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), y, K, HW, C);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}
int avgpool_backward(const float* dy, void* dx, int K, int HW, int C, cudaStream_t stream) {
  VLB_REQUIRE(dy && dx, "avgpool: null pointer");
  const long total = (long)K * HW * C;
  if (total == 0) return VLB_OK;
  avgpool_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(dy, static_cast<__nv_bfloat16*>(dx), K, HW, C);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, cudaStream_t stream) {
  VLB_REQUIRE(x && y, "layout: null pointer");
  const long total = (long)N * C * H * W;
  nchw_f32_to_nhwc_bf16_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y), N, C, H, W);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}
int nhwc_bf16_to_nchw_f32(const void* x, float* y, int N, int C, int H, int W, cudaStream_t stream) {
  VLB_REQUIRE(x && y, "layout: null pointer");
  const long total = (long)N * C * H * W;
  nhwc_bf16_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), y, N, C, H, W);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int roi_align_nhwc_forward(const void* feat, const float* rois, void* out, int K, int C, int H, int W, int ph, int pw, float scale,
                           int sampling_ratio, cudaStream_t stream) {
  if (K == 0) return VLB_OK;
  VLB_REQUIRE(feat && rois && out && C % 8 == 0, "roi_align_nhwc_forward: bad arguments");
  const int warps = K * ph * pw;
  ProfScope prof(PROF_ROI_NHWC, 2.0 * K * ph * pw * C * 5, stream);      // 4 taps read + 1 write per output (sampling_ratio 1)
  roi_align_nhwc_kernel<false><<<(warps + 3) / 4, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(feat), rois,
                                                                   static_cast<__nv_bfloat16*>(out), nullptr, nullptr, K, C, H, W, ph, pw,
                                                                   scale, sampling_ratio);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}
int roi_align_nhwc_backward(const void* grad_out, const float* rois, float* grad_feat, int K, int N, int C, int H, int W, int ph, int pw,
                            float scale, int sampling_ratio, cudaStream_t stream) {
  VLB_REQUIRE(grad_feat, "roi_align_nhwc_backward: null pointer");
  ProfScope prof(PROF_ROI_NHWC, 4.0 * N * C * H * W + 2.0 * K * ph * pw * C + 16.0 * K * ph * pw * C, stream);
  VLB_CHECK_CUDA(cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)N * C * H * W, stream));
  if (K == 0) return VLB_OK;
  VLB_REQUIRE(grad_out && rois && C % 8 == 0, "roi_align_nhwc_backward: bad arguments");
  const int warps = K * ph * pw;
  roi_align_nhwc_kernel<true><<<(warps + 3) / 4, 128, 0, stream>>>(nullptr, rois, nullptr, static_cast<const __nv_bfloat16*>(grad_out),
                                                                  grad_feat, K, C, H, W, ph, pw, scale, sampling_ratio);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

}  // namespace vlb
