// One BertLayer forward / backward as a fixed kernel sequence on one stream.
// Reference: external/pytorch_pretrained_bert/modeling.py:388-397 (BertLayer) =
//   BertSelfAttention :290-315, BertSelfOutput :329-333, BertIntermediate :361-364, BertOutput :374-378.
//
// Precision contract (what tests/test_gpu_parity_bf16.py pins): every GEMM takes bf16 operands, accumulates in fp32 and -- where
// its result is an activation that feeds another GEMM -- stores bf16; the RESIDUAL STREAM stays in fp32, forward and
// backward, like the reference under torch.autocast: the residual operand of the two "dense + residual" epilogues is the fp32
// LayerNorm output, recomputed in the epilogue from the LayerNorm's stored fp32 input and row statistics (no second copy of
// any activation is written), and the gradient of the stream travels as (bf16 GEMM result) + (fp32 LayerNorm gradient), summed
// in fp32 inside the next LayerNorm backward.  x_resid == NULL / dx_f32 == NULL select the older all-bf16 stream.
#include <cstdlib>

#include "ops.cuh"

namespace vlb {

namespace {
// the three dropout sites of a layer as VlbDropout records (p == 0 -> the site is off); keep flags come from acts.keep_*
struct LayerDrop {
  VlbDropout attn, self_out, out;
  LayerDrop(const VlbLayerDropout* d, const VlbLayerActs& a) {
    attn = self_out = out = VlbDropout{0.0f, 0u, nullptr, nullptr};
    if (d != nullptr && d->rng != nullptr) {
      if (d->p_attn > 0.0f) attn = VlbDropout{d->p_attn, d->site_attn, d->rng, a.keep_attn};
      if (d->p_hidden > 0.0f) {
        self_out = VlbDropout{d->p_hidden, d->site_self_out, d->rng, a.keep_self_out};
        out = VlbDropout{d->p_hidden, d->site_out, d->rng, a.keep_out};
      }
    }
  }
};
bool layer_drop_valid(const VlbLayerDropout* d, const VlbLayerActs& a) {
  if (d == nullptr) return true;
  if (!(d->p_attn >= 0.0f && d->p_attn < 1.0f && d->p_hidden >= 0.0f && d->p_hidden < 1.0f)) return false;
  if (d->p_attn == 0.0f && d->p_hidden == 0.0f) return true;
  if (d->rng == nullptr) return false;
  if (d->p_attn > 0.0f && a.keep_attn == nullptr) return false;
  if (d->p_hidden > 0.0f && (a.keep_self_out == nullptr || a.keep_out == nullptr)) return false;
  return true;
}
}  // namespace

// Forward: 7 launches (+ 1 that writes the layer's dropout keep flags).
int bert_layer_forward(const VlbLayerWeights& w, const void* x, const VlbResidual* x_resid, const float* add_mask, const VlbLayerActs& a,
                       int B, int S, int H, int heads, int I, float eps, const VlbLayerDropout* drop, cudaStream_t st) {
  const int M = B * S;
  VLB_REQUIRE(layer_drop_valid(drop, a), "bert_layer_forward: bad dropout configuration (probabilities, rng state or keep-flag buffers)");
  VLB_REQUIRE(x_resid == nullptr || x_resid->x_f32 != nullptr, "bert_layer_forward: x_resid without a tensor");
  const LayerDrop ld(drop, a);
  const bool f32_stream = x_resid != nullptr;
  int rc;
  int launches = 7;
  // 0. keep flags of the layer's three dropout sites (read by the forward kernels below and by the whole backward)
  if ((ld.attn.p > 0.0f || ld.out.p > 0.0f) && !drop->keep_bits_ready) {
    if ((rc = layer_dropout_bits(a.keep_attn, a.keep_self_out, a.keep_out, B, S, H, heads, *drop, st))) return rc;
    ++launches;
  }
  GemmEpilogue e;
  // 1. fused QKV projection: [M,H] x [3H,H]^T + b
  e = GemmEpilogue();
  e.out = a.qkv; e.ldo = 3 * H; e.out_kind = OUT_BF16; e.bias = w.b_qkv;
  if ((rc = gemm_bf16(GEMM_NT, M, 3 * H, H, x, H, w.w_qkv, H, e, 1, 0, st))) return rc;
  // 2. attention
  if ((rc = mhsa_forward(a.qkv, add_mask, a.ctx, a.lse, B, S, H, heads, st, &ld.attn))) return rc;
  // 3. attention output dense + bias -> dropout -> + residual -> fp32
  e = GemmEpilogue();
  e.out = a.a; e.ldo = H; e.out_kind = OUT_F32; e.bias = w.b_o;
  if (f32_stream) {
    e.resid = x_resid->x_f32; e.ldr = H; e.resid_kind = RESID_LN_F32;
    e.ln_mean = x_resid->mean; e.ln_rstd = x_resid->rstd; e.ln_gamma = x_resid->gamma; e.ln_beta = x_resid->beta;
  } else {
    e.resid = x; e.ldr = H; e.resid_kind = RESID_BF16;
  }
  e.drop = make_drop(&ld.self_out);
  if ((rc = gemm_bf16(GEMM_NT, M, H, H, a.ctx, H, w.w_o, H, e, 1, 0, st))) return rc;
  // 4. LayerNorm 1
  if ((rc = layernorm_forward(a.a, H, w.ln1_g, w.ln1_b, a.h, nullptr, a.ln1_mean, a.ln1_rstd, M, H, eps, st))) return rc;
  // 5. intermediate dense + bias + erf-GELU (GELU' of the pre-activation kept for backward)
  e = GemmEpilogue();
  e.out = a.u; e.ldo = I; e.out_kind = OUT_BF16; e.bias = w.b_1; e.act = ACT_GELU; e.aux = a.z; e.ld_aux = I;
  if ((rc = gemm_bf16(GEMM_NT, M, I, H, a.h, H, w.w_1, H, e, 1, 0, st))) return rc;
  // 6. output dense + bias -> dropout -> + residual -> fp32
  e = GemmEpilogue();
  e.out = a.y0; e.ldo = H; e.out_kind = OUT_F32; e.bias = w.b_2;
  if (f32_stream) {   // residual = LayerNorm 1 output in fp32, recomputed from its input a.a
    e.resid = a.a; e.ldr = H; e.resid_kind = RESID_LN_F32;
    e.ln_mean = a.ln1_mean; e.ln_rstd = a.ln1_rstd; e.ln_gamma = w.ln1_g; e.ln_beta = w.ln1_b;
  } else {
    e.resid = a.h; e.ldr = H; e.resid_kind = RESID_BF16;
  }
  e.drop = make_drop(&ld.out);
  if ((rc = gemm_bf16(GEMM_NT, M, H, I, a.u, I, w.w_2, I, e, 1, 0, st))) return rc;
  // 7. LayerNorm 2
  if ((rc = layernorm_forward(a.y0, H, w.ln2_g, w.ln2_b, a.y, a.y_f32, a.ln2_mean, a.ln2_rstd, M, H, eps, st))) return rc;
  count_launch(launches);
  return VLB_OK;
}

int64_t bert_layer_backward_workspace(int M, int H, int I) {
  // d_y0 [M,H] | dz [M,I] | dh [M,H] | d_a [M,H] | dctx [M,H] | dqkv [M,3H]   (bf16, each 256B-aligned)
  // | d_y0 fp32 [M,H] (the residual-stream gradient between the two LayerNorm backwards)
  // | fp32 [M,3H] scratch for the multi-block attention backward (used when S > 128)
  auto al = [](int64_t v) { return (v + 255) & ~int64_t(255); };
  return al((int64_t)M * H * 2) * 4 + al((int64_t)M * I * 2) + al((int64_t)M * 3 * H * 2) + al((int64_t)M * H * 4) + al((int64_t)M * 3 * H * 4);
}

// Backward: 8 launches.
int bert_layer_backward(const VlbLayerWeights& w, const VlbLayerActs& a, const void* x, const float* add_mask, const void* dy16,
                        const float* dy32, void* dx, float* dx_f32, const VlbLayerGrads& g, void* workspace, int64_t ws_bytes, int B, int S,
                        int H, int heads, int I, const VlbLayerDropout* drop, cudaStream_t st) {
  const int M = B * S;
  VLB_REQUIRE(layer_drop_valid(drop, a), "bert_layer_backward: bad dropout configuration (probabilities, rng state or keep-flag buffers)");
  const LayerDrop ld(drop, a);
  const bool hdrop = ld.out.p > 0.0f;
  const bool f32_stream = dx_f32 != nullptr;
  VLB_REQUIRE(ws_bytes >= bert_layer_backward_workspace(M, H, I), "bert_layer_backward: workspace too small");
  VLB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "bert_layer_backward: workspace must be 256B aligned");
  auto al = [](int64_t v) { return (v + 255) & ~int64_t(255); };
  uint8_t* p = static_cast<uint8_t*>(workspace);
  void* d_y0 = p; p += al((int64_t)M * H * 2);   // bf16 operand of the dense branch: LayerNorm-2 gradient (x dropout mask of BertOutput)
  void* dz = p;   p += al((int64_t)M * I * 2);
  void* dh = p;   p += al((int64_t)M * H * 2);
  void* d_a = p;  p += al((int64_t)M * H * 2);   // bf16 operand: LayerNorm-1 gradient (x dropout mask of BertSelfOutput)
  void* dctx = p; p += al((int64_t)M * H * 2);
  void* dqkv = p; p += al((int64_t)M * 3 * H * 2);
  float* d_y0_32 = reinterpret_cast<float*>(p); p += al((int64_t)M * H * 4);
  float* attn_scratch = reinterpret_cast<float*>(p);
  int rc;
  GemmEpilogue e;
  // In the all-bf16 stream the residual branch needs the UNmasked bf16 gradient as well: with dropout that is a second
  // bf16 tensor; it borrows dctx / dqkv, which are written only after their last use as residuals.
  void* d_y0_plain = d_y0;
  void* d_a_plain = d_a;
  if (!f32_stream && hdrop) { d_y0_plain = dctx; d_a_plain = dqkv; }
  // LayerNorm 2 backward: dense-branch operand (bf16, masked), residual-branch gradient (fp32), dgamma2/dbeta2, db_2 = colsum(operand)
  if (f32_stream) {
    if ((rc = layernorm_backward(dy16, dy32, a.y0, H, a.ln2_mean, a.ln2_rstd, w.ln2_g, hdrop ? nullptr : d_y0, d_y0_32, H, g.dln2_g, g.dln2_b,
                                 g.db_2, M, H, st, nullptr, hdrop ? d_y0 : nullptr, hdrop ? &ld.out : nullptr))) return rc;
  } else {
    if ((rc = layernorm_backward(dy16, dy32, a.y0, H, a.ln2_mean, a.ln2_rstd, w.ln2_g, d_y0_plain, nullptr, 0, g.dln2_g, g.dln2_b,
                                 g.db_2, M, H, st, nullptr, hdrop ? d_y0 : nullptr, hdrop ? &ld.out : nullptr))) return rc;
  }
  // dz = (d_y0 W2) o gelu'(z) ; db_1 += colsum(dz) fused into the same epilogue
  e = GemmEpilogue(); e.out = dz; e.ldo = I; e.out_kind = OUT_BF16; e.act = ACT_DGELU_MUL; e.aux = a.z; e.ld_aux = I;
  e.colsum = g.db_1;
  if ((rc = gemm_bf16(GEMM_NN, M, I, H, d_y0, H, w.w_2, I, e, 1, 0, st))) return rc;
  // dh = dz W1 (+ d_y0: the residual gradient joins in fp32 inside the next LayerNorm backward, or here in bf16)
  e = GemmEpilogue(); e.out = dh; e.ldo = H; e.out_kind = OUT_BF16;
  if (!f32_stream) { e.resid = d_y0_plain; e.ldr = H; e.resid_kind = RESID_BF16; }
  if ((rc = gemm_bf16(GEMM_NN, M, H, I, dz, I, w.w_1, H, e, 1, 0, st))) return rc;
  // LayerNorm 1 backward: operand d_a (bf16, masked), residual-stream gradient (fp32, straight into the caller's dx_f32), db_o
  if (f32_stream) {
    if ((rc = layernorm_backward(dh, d_y0_32, a.a, H, a.ln1_mean, a.ln1_rstd, w.ln1_g, hdrop ? nullptr : d_a, dx_f32, H, g.dln1_g, g.dln1_b,
                                 g.db_o, M, H, st, nullptr, hdrop ? d_a : nullptr, hdrop ? &ld.self_out : nullptr))) return rc;
  } else {
    if ((rc = layernorm_backward(dh, nullptr, a.a, H, a.ln1_mean, a.ln1_rstd, w.ln1_g, d_a_plain, nullptr, 0, g.dln1_g, g.dln1_b,
                                 g.db_o, M, H, st, nullptr, hdrop ? d_a : nullptr, hdrop ? &ld.self_out : nullptr))) return rc;
  }
  if (!f32_stream && hdrop) {
    // the bf16 stream with dropout: d_a_plain lives in dqkv and must survive until the last GEMM; move it out of the way
    VLB_CHECK_CUDA(cudaMemcpyAsync(dh, d_a_plain, (size_t)M * H * 2, cudaMemcpyDeviceToDevice, st));
    d_a_plain = dh;
  }
  // dctx = d_a Wo
  e = GemmEpilogue(); e.out = dctx; e.ldo = H; e.out_kind = OUT_BF16;
  if ((rc = gemm_bf16(GEMM_NN, M, H, H, d_a, H, w.w_o, H, e, 1, 0, st))) return rc;
  // attention backward
  // attention backward; db_qkv += colsum(dqkv) is fused into its store phase (a transposing warp reduction + one atomic per lane)
  if ((rc = mhsa_backward(a.qkv, add_mask, a.ctx, a.lse, dctx, dqkv, attn_scratch, B, S, H, heads, st, &ld.attn, g.db_qkv))) return rc;
  // dx = dqkv Wqkv (+ d_a in the bf16 stream)
  e = GemmEpilogue(); e.out = dx; e.ldo = H; e.out_kind = OUT_BF16;
  if (!f32_stream) { e.resid = d_a_plain; e.ldr = H; e.resid_kind = RESID_BF16; }
  if ((rc = gemm_bf16(GEMM_NN, M, H, 3 * H, dqkv, 3 * H, w.w_qkv, H, e, 1, 0, st))) return rc;
  // all four weight gradients of the layer in ONE grouped launch (every operand is still live in the workspace):
  //   dW2 += d_y0^T u ; dW1 += dz^T h ; dWo += d_a^T ctx ; dWqkv += dqkv^T x     (reduction over the M token rows; d_y0 / d_a are
  //   the dense-branch operands, i.e. already multiplied by the dropout mask of that dense output)
  {
    GroupedProblem q[4] = {
        {H, I, d_y0, H, a.u, I, g.dw_2, I},
        {I, H, dz, I, a.h, H, g.dw_1, H},
        {3 * H, H, dqkv, 3 * H, x, H, g.dw_qkv, H},
        {H, H, d_a, H, a.ctx, H, g.dw_o, H},
    };
    static const int env_bn = [] { const char* v = getenv("VLB_WGRAD_BN"); return v ? atoi(v) : 256; }();
    static const int env_split = [] { const char* v = getenv("VLB_WGRAD_SPLIT"); return v ? atoi(v) : 2; }();
    if ((rc = gemm_grouped_tn(4, q, M, env_split, true, env_bn, st))) return rc;
  }
  count_launch(8);
  return VLB_OK;
}

}  // namespace vlb
