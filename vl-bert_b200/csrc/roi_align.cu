// Region-feature front end kernels.
//
//  * RoIAlign forward / backward with the reference's exact conventions
//    (common/lib/roi_pooling/cuda/ROIAlign_cuda.cu:15-62 sampling, :64-122 forward, :125-175/:177-254 backward;
//    rois = (batch_idx, x1, y1, x2, y2) scaled by spatial_scale, no rounding, no half-pixel shift, roi size
//    clamped >= 1, mean over sampling_ratio^2 (or ceil(roi/pooled)^2) bilinear samples).
//    The reference launches one thread per output scalar and re-derives the bin geometry per channel; here a
//    thread owns one output bin (ph, pw) of one RoI, computes the sample weights once and streams over a slab
//    of channels, so writes are coalesced along the 196 contiguous bins of a channel and the geometry cost is
//    amortised over the channels.  Arithmetic is ordered (and FMA-contraction disabled) so that the forward is
//    bit-identical to the reference CPU kernel / the C oracle.
//  * coordinate embedding + feature concat (common/utils/bbox.py:33-65, common/fast_rcnn.py:170-174): builds the
//    bf16 A operand [B*R, 4096] of the obj_downsample GEMM for the precomputed-feature path.
#include "common.cuh"
#include "philox.cuh"

namespace vlb {

namespace {

struct Sample {
  int i0, i1, i2, i3;   // plane offsets, i0 < 0 => outside
  float w0, w1, w2, w3;
};

__device__ __forceinline__ Sample make_sample(int height, int width, float y, float x) {
  Sample s;
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) {
    s.i0 = s.i1 = s.i2 = s.i3 = -1;
    s.w0 = s.w1 = s.w2 = s.w3 = 0.0f;
    return s;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = __fsub_rn(y, (float)y_low), lx = __fsub_rn(x, (float)x_low);
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  s.w0 = __fmul_rn(hy, hx); s.w1 = __fmul_rn(hy, lx); s.w2 = __fmul_rn(ly, hx); s.w3 = __fmul_rn(ly, lx);
  s.i0 = y_low * width + x_low;  s.i1 = y_low * width + x_high;
  s.i2 = y_high * width + x_low; s.i3 = y_high * width + x_high;
  return s;
}

struct RoiGeom {
  int b;
  float start_w, start_h, bin_w, bin_h;
  int grid_h, grid_w;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* roi, float scale, int ph_n, int pw_n, int sampling_ratio) {
  RoiGeom g;
  g.b = (int)roi[0];
  const float sw = __fmul_rn(roi[1], scale), sh = __fmul_rn(roi[2], scale);
  const float ew = __fmul_rn(roi[3], scale), eh = __fmul_rn(roi[4], scale);
  const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
  g.start_w = sw; g.start_h = sh;
  g.bin_h = __fdiv_rn(rh, (float)ph_n);
  g.bin_w = __fdiv_rn(rw, (float)pw_n);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)ph_n));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)pw_n));
  return g;
}

__device__ __forceinline__ float sample_y(const RoiGeom& g, int ph, int iy) {
  // roi_start_h + ph * bin_size_h + (iy + .5f) * bin_size_h / grid_h     (left-to-right, unfused)
  return __fadd_rn(__fadd_rn(g.start_h, __fmul_rn((float)ph, g.bin_h)),
                   __fdiv_rn(__fmul_rn((float)iy + 0.5f, g.bin_h), (float)g.grid_h));
}
__device__ __forceinline__ float sample_x(const RoiGeom& g, int pw, int ix) {
  return __fadd_rn(__fadd_rn(g.start_w, __fmul_rn((float)pw, g.bin_w)),
                   __fdiv_rn(__fmul_rn((float)ix + 0.5f, g.bin_w), (float)g.grid_w));
}

// grid = (K rois, channel slabs); block = bins rounded up to a warp multiple.
__global__ void roi_align_fwd_kernel(const float* __restrict__ input, const float* __restrict__ rois, float* __restrict__ out,
                                     int C, int H, int W, int ph_n, int pw_n, float scale, int sampling_ratio, int c_per_block) {
  const int n = blockIdx.x;
  const int bin = threadIdx.x;
  const int nbins = ph_n * pw_n;
  if (bin >= nbins) return;
  const int ph = bin / pw_n, pw = bin - ph * pw_n;
  const RoiGeom g = roi_geometry(rois + 5 * n, scale, ph_n, pw_n, sampling_ratio);
  const float count = (float)(g.grid_h * g.grid_w);
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(C, c0 + c_per_block);
  const size_t plane = (size_t)H * W;
  const float* in_b = input + (size_t)g.b * C * plane;
  float* out_n = out + (size_t)n * C * nbins + bin;
  if (g.grid_h == 1 && g.grid_w == 1) {
    // the reference configuration (sampling_ratio = 1, common/lib/roi_pooling/roi_align.py:51)
    const Sample s = make_sample(H, W, sample_y(g, ph, 0), sample_x(g, pw, 0));
    for (int c = c0; c < c1; ++c) {
      float v = 0.0f;
      if (s.i0 >= 0) {
        const float* p = in_b + c * plane;
        v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.w0, __ldg(p + s.i0)), __fmul_rn(s.w1, __ldg(p + s.i1))),
                                __fmul_rn(s.w2, __ldg(p + s.i2))),
                      __fmul_rn(s.w3, __ldg(p + s.i3)));
      }
      out_n[(size_t)c * nbins] = __fdiv_rn(__fadd_rn(0.0f, v), count);
    }
  } else {
    for (int c = c0; c < c1; ++c) {
      const float* p = in_b + c * plane;
      float acc = 0.0f;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const float y = sample_y(g, ph, iy);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const Sample s = make_sample(H, W, y, sample_x(g, pw, ix));
          if (s.i0 >= 0) {
            const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.w0, __ldg(p + s.i0)), __fmul_rn(s.w1, __ldg(p + s.i1))),
                                                __fmul_rn(s.w2, __ldg(p + s.i2))),
                                      __fmul_rn(s.w3, __ldg(p + s.i3)));
            acc = __fadd_rn(acc, v);
          }
        }
      }
      out_n[(size_t)c * nbins] = __fdiv_rn(acc, count);
    }
  }
}

__global__ void roi_align_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois,
                                     float* __restrict__ grad_in, int C, int H, int W, int ph_n, int pw_n, float scale,
                                     int sampling_ratio, int c_per_block) {
  const int n = blockIdx.x;
  const int bin = threadIdx.x;
  const int nbins = ph_n * pw_n;
  if (bin >= nbins) return;
  const int ph = bin / pw_n, pw = bin - ph * pw_n;
  const RoiGeom g = roi_geometry(rois + 5 * n, scale, ph_n, pw_n, sampling_ratio);
  const float count = (float)(g.grid_h * g.grid_w);
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(C, c0 + c_per_block);
  const size_t plane = (size_t)H * W;
  float* gin_b = grad_in + (size_t)g.b * C * plane;
  const float* go_n = grad_out + (size_t)n * C * nbins + bin;
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const float y = sample_y(g, ph, iy);
    for (int ix = 0; ix < g.grid_w; ++ix) {
      const Sample s = make_sample(H, W, y, sample_x(g, pw, ix));
      if (s.i0 < 0) continue;
      for (int c = c0; c < c1; ++c) {
        const float gv = go_n[(size_t)c * nbins];
        float* p = gin_b + c * plane;
        atomicAdd(p + s.i0, __fdiv_rn(__fmul_rn(gv, s.w0), count));
        atomicAdd(p + s.i1, __fdiv_rn(__fmul_rn(gv, s.w1), count));
        atomicAdd(p + s.i2, __fdiv_rn(__fmul_rn(gv, s.w2), count));
        atomicAdd(p + s.i3, __fdiv_rn(__fmul_rn(gv, s.w3), count));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// A operand of obj_downsample for the precomputed-feature path.  One block per (b, r) slot:
//   cols [0, 2048)    : 4 coordinates (xc/W, yc/H, w/W, h/H) * 100 -> sin(pos / 1000^(i/256)) | cos(...), i < 256
//   cols [2048, 4096) : the 2048-d precomputed feature boxes[b, r, 4:]  (or mask_visual_embed where mvrc_ops == 1)
// Slots with box_mask == 0 produce an all-zero row (their GEMM output is discarded by the re-padding gather).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
region_operand_kernel(const float* __restrict__ boxes, int ld_box, const uint8_t* __restrict__ box_mask,
                      const float* __restrict__ im_info, int ld_info, const int64_t* __restrict__ mvrc_ops,
                      const float* __restrict__ mask_visual_embed, __nv_bfloat16* __restrict__ A, int R, int feat_dim,
                      const DropCfg drop) {
  const int slot = blockIdx.x;
  const int b = slot / R;
  __nv_bfloat16* arow = A + (size_t)slot * (2048 + feat_dim);
  const int ncol = 2048 + feat_dim;
  if (!box_mask[slot]) {
    for (int c = threadIdx.x * 8; c < ncol; c += 256 * 8) *reinterpret_cast<uint4*>(arow + c) = make_uint4(0, 0, 0, 0);
    return;
  }
  const float* bx = boxes + (size_t)slot * ld_box;
  const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
  const float Wd = im_info[(size_t)b * ld_info], Hd = im_info[(size_t)b * ld_info + 1];
  float pos[4];
  pos[0] = (x1 + x2) / 2.0f / Wd * 100.0f;
  pos[1] = (y1 + y2) / 2.0f / Hd * 100.0f;
  pos[2] = (x2 - x1) / Wd * 100.0f;
  pos[3] = (y2 - y1) / Hd * 100.0f;
  // Dropout(0.1) at the head of obj_downsample (common/fast_rcnn.py:104-109): mask over the [B*R, ncol] slot layout
  const DropState dstate = drop_state(drop);
  const uint64_t row_group0 = ((uint64_t)slot * (uint64_t)ncol) >> 2;   // ncol % 8 == 0
  // 4 coordinates x (sin | cos) x 256 frequencies (common/utils/bbox.py:33-65)
  for (int e = threadIdx.x; e < 512; e += 256) {   // e -> (coordinate c, sin/cos half, group of four frequencies q)
    const int c = e >> 7, hc = (e >> 6) & 1, q = e & 63;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = q * 4 + k;
      const float dim = powf(1000.0f, (float)i / 256.0f);
      float sv, cv;
      sincosf(pos[c] / dim, &sv, &cv);
      v[k] = hc ? cv : sv;
    }
    const int col = c * 512 + hc * 256 + q * 4;
    if (drop.thresh != 0u) drop4(v, row_group0 + (uint64_t)(col >> 2), drop, dstate);
    uint2 pk;
    pk.x = pack_bf16x2(v[0], v[1]);
    pk.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(arow + col) = pk;
  }
  const float* f = bx + 4;
  if (mvrc_ops != nullptr && mask_visual_embed != nullptr && mvrc_ops[slot] == 1) f = mask_visual_embed;
  for (int c = threadIdx.x * 4; c < feat_dim; c += 256 * 4) {
    const float4 vv = *reinterpret_cast<const float4*>(f + c);
    float v[4] = {vv.x, vv.y, vv.z, vv.w};
    if (drop.thresh != 0u) drop4(v, row_group0 + (uint64_t)((2048 + c) >> 2), drop, dstate);
    uint2 pk;
    pk.x = pack_bf16x2(v[0], v[1]);
    pk.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(arow + 2048 + c) = pk;
  }
}

// k-th valid box of sample b -> slot k (pad_sequence, common/utils/pad_sequence.py:4-17):
//   gather_idx[b*R + k] = b*R + r_k  for k < count_b else -1
__global__ void box_slot_index_kernel(const uint8_t* __restrict__ box_mask, int32_t* __restrict__ gather_idx, int R) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int k = 0;
    for (int r = 0; r < R; ++r)
      if (box_mask[(size_t)b * R + r]) gather_idx[(size_t)b * R + k++] = b * R + r;
    for (; k < R; ++k) gather_idx[(size_t)b * R + k] = -1;
  }
}

}  // namespace

int roi_align_forward(const float* input, const float* rois, float* out, int K, int C, int H, int W, int ph, int pw,
                      float spatial_scale, int sampling_ratio, cudaStream_t stream) {
  if (K == 0) return VLB_OK;  // zero-size output short-circuits (ROIAlign_cuda.cu:278-281)
  VLB_REQUIRE(input && rois && out, "roi_align_forward: null pointer");
  VLB_REQUIRE(ph > 0 && pw > 0 && ph * pw <= 1024 && C > 0 && H > 0 && W > 0, "roi_align_forward: bad sizes");
  const int threads = ((ph * pw + 31) / 32) * 32;
  int slabs = (num_sms() * 8 + K - 1) / K;
  if (slabs > C) slabs = C;
  if (slabs < 1) slabs = 1;
  const int c_per = (C + slabs - 1) / slabs;
  slabs = (C + c_per - 1) / c_per;
  roi_align_fwd_kernel<<<dim3(K, slabs), threads, 0, stream>>>(input, rois, out, C, H, W, ph, pw, spatial_scale, sampling_ratio,
                                                               c_per);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int roi_align_backward(const float* grad_out, const float* rois, float* grad_in, int K, int N, int C, int H, int W, int ph,
                       int pw, float spatial_scale, int sampling_ratio, cudaStream_t stream) {
  VLB_REQUIRE(grad_in, "roi_align_backward: null pointer");
  VLB_CHECK_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * (size_t)N * C * H * W, stream));
  if (K == 0) return VLB_OK;
  VLB_REQUIRE(grad_out && rois, "roi_align_backward: null pointer");
  VLB_REQUIRE(ph > 0 && pw > 0 && ph * pw <= 1024, "roi_align_backward: bad sizes");
  const int threads = ((ph * pw + 31) / 32) * 32;
  int slabs = (num_sms() * 8 + K - 1) / K;
  if (slabs > C) slabs = C;
  if (slabs < 1) slabs = 1;
  const int c_per = (C + slabs - 1) / slabs;
  slabs = (C + c_per - 1) / c_per;
  roi_align_bwd_kernel<<<dim3(K, slabs), threads, 0, stream>>>(grad_out, rois, grad_in, C, H, W, ph, pw, spatial_scale,
                                                               sampling_ratio, c_per);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int region_operand(const float* boxes, int ld_box, const uint8_t* box_mask, const float* im_info, int ld_info,
                   const int64_t* mvrc_ops, const float* mask_visual_embed, void* A, int32_t* gather_idx, int B, int R,
                   int feat_dim, cudaStream_t stream, const VlbDropout* drop) {
  VLB_REQUIRE(boxes && box_mask && im_info && A && gather_idx, "region_operand: null pointer");
  VLB_REQUIRE(drop_valid(drop), "region_operand: bad dropout configuration");
  VLB_REQUIRE(feat_dim % 8 == 0 && ld_box % 4 == 0 && ld_box >= 4 + feat_dim, "region_operand: bad feature layout");
  region_operand_kernel<<<B * R, 256, 0, stream>>>(boxes, ld_box, box_mask, im_info, ld_info, mvrc_ops, mask_visual_embed,
                                                   static_cast<__nv_bfloat16*>(A), R, feat_dim, make_drop(drop));
  VLB_CHECK_LAUNCH();
  box_slot_index_kernel<<<B, 32, 0, stream>>>(box_mask, gather_idx, R);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

}  // namespace vlb
