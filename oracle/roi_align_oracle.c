/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference RoIAlign (fp32, NCHW).
 *
 * Follows common/lib/roi_pooling/cuda/ROIAlign_cuda.cu: sampling of one bilinear point :15-62,
 * forward per output element :64-122, weights/indices of one point for backward :125-175, backward
 * scatter :177-254 (the CUDA kernel's atomicAdd order is unspecified; here the sum is sequential in
 * (roi, channel, ph, pw) order, so backward parity is "equal up to fp32 summation order").
 * Conventions kept from the reference: roi = (batch_idx, x1, y1, x2, y2) in image pixels scaled by
 * spatial_scale with NO rounding and NO half-pixel shift (:81-89), roi width/height clamped to >= 1
 * (:92-93), sampling grid = sampling_ratio or ceil(roi_size / pooled_size) (:99-100), mean over the
 * grid (:103,118).
 *
 * Pinned by tests/test_oracle_roialign.py against torchvision.ops.roi_align(aligned=False), which
 * SURVEY.md section 8(c) showed bit-identical to the reference's own CPU kernel
 * (common/lib/roi_pooling/cpu/ROIAlign_cpu.cpp), and against tests/golden/roi_align_debug.npz.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libroi_align_oracle.so oracle/roi_align_oracle.c -lm
 */
#include <math.h>
#include <string.h>

static void point_weights(int height, int width, float y, float x, float w[4], int idx[4]) {
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) {
    w[0] = w[1] = w[2] = w[3] = 0.0f;
    idx[0] = idx[1] = idx[2] = idx[3] = -1;
    return;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  float ly = y - (float)y_low, lx = x - (float)x_low;
  float hy = 1.0f - ly, hx = 1.0f - lx;
  w[0] = hy * hx; w[1] = hy * lx; w[2] = ly * hx; w[3] = ly * lx;
  idx[0] = y_low * width + x_low;  idx[1] = y_low * width + x_high;
  idx[2] = y_high * width + x_low; idx[3] = y_high * width + x_high;
}

static void roi_geometry(const float* roi, float scale, int ph_n, int pw_n, int sampling_ratio,
                         float* start_w, float* start_h, float* bin_w, float* bin_h, int* grid_h, int* grid_w) {
  float sw = roi[1] * scale, sh = roi[2] * scale, ew = roi[3] * scale, eh = roi[4] * scale;
  float rw = fmaxf(ew - sw, 1.0f), rh = fmaxf(eh - sh, 1.0f);
  *start_w = sw; *start_h = sh;
  *bin_h = rh / (float)ph_n; *bin_w = rw / (float)pw_n;
  *grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph_n);
  *grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw_n);
}

/* input [N,C,H,W], rois [K,5], output [K,C,ph,pw] */
void roi_align_forward_oracle(const float* input, const float* rois, int K, int C, int H, int W, int ph_n, int pw_n,
                              float spatial_scale, int sampling_ratio, float* output) {
  for (int n = 0; n < K; ++n) {
    const float* roi = rois + 5 * n;
    int b = (int)roi[0];
    float sw, sh, bw, bh; int gh, gw;
    roi_geometry(roi, spatial_scale, ph_n, pw_n, sampling_ratio, &sw, &sh, &bw, &bh, &gh, &gw);
    const float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      const float* plane = input + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < ph_n; ++ph) for (int pw = 0; pw < pw_n; ++pw) {
        float acc = 0.0f;
        for (int iy = 0; iy < gh; ++iy) {
          const float y = sh + ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
          for (int ix = 0; ix < gw; ++ix) {
            const float x = sw + pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
            float w[4]; int idx[4];
            point_weights(H, W, y, x, w, idx);
            if (idx[0] >= 0)
              acc += w[0] * plane[idx[0]] + w[1] * plane[idx[1]] + w[2] * plane[idx[2]] + w[3] * plane[idx[3]];
          }
        }
        output[(((size_t)n * C + c) * ph_n + ph) * pw_n + pw] = acc / count;
      }
    }
  }
}

/* grad_out [K,C,ph,pw] -> grad_in [N,C,H,W] (zero-initialised here) */
void roi_align_backward_oracle(const float* grad_out, const float* rois, int K, int N, int C, int H, int W, int ph_n,
                               int pw_n, float spatial_scale, int sampling_ratio, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < K; ++n) {
    const float* roi = rois + 5 * n;
    int b = (int)roi[0];
    float sw, sh, bw, bh; int gh, gw;
    roi_geometry(roi, spatial_scale, ph_n, pw_n, sampling_ratio, &sw, &sh, &bw, &bh, &gh, &gw);
    const float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      float* plane = grad_in + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < ph_n; ++ph) for (int pw = 0; pw < pw_n; ++pw) {
        const float g = grad_out[(((size_t)n * C + c) * ph_n + ph) * pw_n + pw];
        for (int iy = 0; iy < gh; ++iy) {
          const float y = sh + ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
          for (int ix = 0; ix < gw; ++ix) {
            const float x = sw + pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
            float w[4]; int idx[4];
            point_weights(H, W, y, x, w, idx);
            if (idx[0] >= 0) {
              plane[idx[0]] += g * w[0] / count; plane[idx[1]] += g * w[1] / count;
              plane[idx[2]] += g * w[2] / count; plane[idx[3]] += g * w[3] / count;
            }
          }
        }
      }
    }
  }
}
