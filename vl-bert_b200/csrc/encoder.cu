// One BertLayer forward / backward as a fixed kernel sequence on one stream.
// Reference: external/pytorch_pretrained_bert/modeling.py:388-397 (BertLayer) =
//   BertSelfAttention :290-315, BertSelfOutput :329-333, BertIntermediate :361-364, BertOutput :374-378.
#include <cstdlib>

#include "ops.cuh"

namespace vlb {

namespace {
// split-K factor for a wgrad GEMM whose output has `tiles` 128x128 tiles: fill the machine once.
int wgrad_split(int n_out, int k_out, int kred) {
  const int tiles = ((n_out + 127) / 128) * ((k_out + 127) / 128);
  int s = num_sms() / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  const int kb = (kred + 63) / 64;
  if (s > kb) s = kb;
  if (s > 16) s = 16;
  return s;
}
}  // namespace

// Forward: 7 launches.
namespace {
// the three dropout sites of a layer as VlbDropout records (p == 0 -> the site is off)
struct LayerDrop {
  VlbDropout attn, self_out, out;
  explicit LayerDrop(const VlbLayerDropout* d) {
    attn = self_out = out = VlbDropout{0.0f, 0u, nullptr};
    if (d != nullptr && d->rng != nullptr) {
      attn = VlbDropout{d->p_attn, d->site_attn, d->rng};
      self_out = VlbDropout{d->p_hidden, d->site_self_out, d->rng};
      out = VlbDropout{d->p_hidden, d->site_out, d->rng};
    }
  }
};
bool layer_drop_valid(const VlbLayerDropout* d) {
  return d == nullptr || (d->p_attn >= 0.0f && d->p_attn < 1.0f && d->p_hidden >= 0.0f && d->p_hidden < 1.0f &&
                          ((d->p_attn == 0.0f && d->p_hidden == 0.0f) || d->rng != nullptr));
}
}  // namespace

int bert_layer_forward(const VlbLayerWeights& w, const void* x, const float* add_mask, const VlbLayerActs& a, int B, int S, int H,
                       int heads, int I, float eps, const VlbLayerDropout* drop, cudaStream_t st) {
  const int M = B * S;
  VLB_REQUIRE(layer_drop_valid(drop), "bert_layer_forward: bad dropout configuration");
  const LayerDrop ld(drop);
  int rc;
  GemmEpilogue e;
  // 1. fused QKV projection: [M,H] x [3H,H]^T + b
  e = GemmEpilogue();
  e.out = a.qkv; e.ldo = 3 * H; e.out_kind = OUT_BF16; e.bias = w.b_qkv;
  if ((rc = gemm_bf16(GEMM_NT, M, 3 * H, H, x, H, w.w_qkv, H, e, 1, 0, st))) return rc;
  // 2. attention
  if ((rc = mhsa_forward(a.qkv, add_mask, a.ctx, a.lse, B, S, H, heads, st, &ld.attn))) return rc;
  // 3. attention output dense + bias -> dropout -> + residual -> fp32
  e = GemmEpilogue();
  e.out = a.a; e.ldo = H; e.out_kind = OUT_F32; e.bias = w.b_o; e.resid = x; e.ldr = H; e.resid_kind = RESID_BF16;
  e.drop = make_drop(&ld.self_out);
  if ((rc = gemm_bf16(GEMM_NT, M, H, H, a.ctx, H, w.w_o, H, e, 1, 0, st))) return rc;
  // 4. LayerNorm 1
  if ((rc = layernorm_forward(a.a, H, w.ln1_g, w.ln1_b, a.h, nullptr, a.ln1_mean, a.ln1_rstd, M, H, eps, st))) return rc;
  // 5. intermediate dense + bias + erf-GELU (pre-activation kept for backward)
  e = GemmEpilogue();
  e.out = a.u; e.ldo = I; e.out_kind = OUT_BF16; e.bias = w.b_1; e.act = ACT_GELU; e.aux = a.z; e.ld_aux = I;
  if ((rc = gemm_bf16(GEMM_NT, M, I, H, a.h, H, w.w_1, H, e, 1, 0, st))) return rc;
  // 6. output dense + bias -> dropout -> + residual -> fp32
  e = GemmEpilogue();
  e.out = a.y0; e.ldo = H; e.out_kind = OUT_F32; e.bias = w.b_2; e.resid = a.h; e.ldr = H; e.resid_kind = RESID_BF16;
  e.drop = make_drop(&ld.out);
  if ((rc = gemm_bf16(GEMM_NT, M, H, I, a.u, I, w.w_2, I, e, 1, 0, st))) return rc;
  // 7. LayerNorm 2
  if ((rc = layernorm_forward(a.y0, H, w.ln2_g, w.ln2_b, a.y, a.y_f32, a.ln2_mean, a.ln2_rstd, M, H, eps, st))) return rc;
  count_launch(7);
  return VLB_OK;
}

int64_t bert_layer_backward_workspace(int M, int H, int I) {
  // d_y0 [M,H] | dz [M,I] | dh [M,H] | d_a [M,H] | dctx [M,H] | dqkv [M,3H] | d_y0' [M,H] | d_a' [M,H]   (bf16), each
  // 256B-aligned (the primed copies = gradient x dropout mask, the operand of the dense branch; used when p_hidden > 0),
  // + fp32 [M,3H] scratch for the multi-block attention backward (used when S > 128)
  auto al = [](int64_t v) { return (v + 255) & ~int64_t(255); };
  return al((int64_t)M * H * 2) * 6 + al((int64_t)M * I * 2) + al((int64_t)M * 3 * H * 2) + al((int64_t)M * 3 * H * 4);
}

// Backward: 9 launches.
int bert_layer_backward(const VlbLayerWeights& w, const VlbLayerActs& a, const void* x, const float* add_mask, const void* dy16,
                        const float* dy32, void* dx, const VlbLayerGrads& g, void* workspace, int64_t ws_bytes, int B, int S,
                        int H, int heads, int I, const VlbLayerDropout* drop, cudaStream_t st) {
  const int M = B * S;
  VLB_REQUIRE(layer_drop_valid(drop), "bert_layer_backward: bad dropout configuration");
  const LayerDrop ld(drop);
  const bool hdrop = ld.out.p > 0.0f;
  VLB_REQUIRE(ws_bytes >= bert_layer_backward_workspace(M, H, I), "bert_layer_backward: workspace too small");
  VLB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "bert_layer_backward: workspace must be 256B aligned");
  auto al = [](int64_t v) { return (v + 255) & ~int64_t(255); };
  uint8_t* p = static_cast<uint8_t*>(workspace);
  void* d_y0 = p; p += al((int64_t)M * H * 2);
  void* dz = p;   p += al((int64_t)M * I * 2);
  void* dh = p;   p += al((int64_t)M * H * 2);
  void* d_a = p;  p += al((int64_t)M * H * 2);
  void* dctx = p; p += al((int64_t)M * H * 2);
  void* dqkv = p; p += al((int64_t)M * 3 * H * 2);
  void* d_y0m = p; p += al((int64_t)M * H * 2);   // d_y0 o mask(site_out) / (1-p): gradient wrt dense(u) of BertOutput
  void* d_am = p;  p += al((int64_t)M * H * 2);   // d_a  o mask(site_self_out) / (1-p): gradient wrt dense(ctx) of BertSelfOutput
  if (!hdrop) { d_y0m = d_y0; d_am = d_a; }
  float* attn_scratch = reinterpret_cast<float*>(p);
  int rc;
  GemmEpilogue e;
  // LayerNorm 2 backward: d_y0 (bf16), dgamma2/dbeta2, db_2 = colsum(d_y0)
  if ((rc = layernorm_backward(dy16, dy32, a.y0, H, a.ln2_mean, a.ln2_rstd, w.ln2_g, d_y0, nullptr, 0, g.dln2_g, g.dln2_b,
                               g.db_2, M, H, st, nullptr, hdrop ? d_y0m : nullptr, hdrop ? &ld.out : nullptr))) return rc;
  // dz = (d_y0 W2) o gelu'(z) ; db_1 += colsum(dz) fused into the same epilogue
  e = GemmEpilogue(); e.out = dz; e.ldo = I; e.out_kind = OUT_BF16; e.act = ACT_DGELU_MUL; e.aux = a.z; e.ld_aux = I;
  e.colsum = g.db_1;
  if ((rc = gemm_bf16(GEMM_NN, M, I, H, d_y0m, H, w.w_2, I, e, 1, 0, st))) return rc;
  // dh = dz W1 + d_y0 (residual)
  e = GemmEpilogue(); e.out = dh; e.ldo = H; e.out_kind = OUT_BF16; e.resid = d_y0; e.ldr = H; e.resid_kind = RESID_BF16;
  if ((rc = gemm_bf16(GEMM_NN, M, H, I, dz, I, w.w_1, H, e, 1, 0, st))) return rc;
  // LayerNorm 1 backward: d_a, dgamma1/dbeta1, db_o = colsum(d_a)
  if ((rc = layernorm_backward(dh, nullptr, a.a, H, a.ln1_mean, a.ln1_rstd, w.ln1_g, d_a, nullptr, 0, g.dln1_g, g.dln1_b,
                               g.db_o, M, H, st, nullptr, hdrop ? d_am : nullptr, hdrop ? &ld.self_out : nullptr))) return rc;
  // dctx = d_a' Wo
  e = GemmEpilogue(); e.out = dctx; e.ldo = H; e.out_kind = OUT_BF16;
  if ((rc = gemm_bf16(GEMM_NN, M, H, H, d_am, H, w.w_o, H, e, 1, 0, st))) return rc;
  // attention backward
  if ((rc = mhsa_backward(a.qkv, add_mask, a.ctx, a.lse, dctx, dqkv, attn_scratch, B, S, H, heads, st, &ld.attn))) return rc;
  // db_qkv += colsum(dqkv) ; dx = dqkv Wqkv + d_a (residual)
  if ((rc = colsum_bf16(dqkv, 3 * H, g.db_qkv, M, 3 * H, st))) return rc;
  e = GemmEpilogue(); e.out = dx; e.ldo = H; e.out_kind = OUT_BF16; e.resid = d_a; e.ldr = H; e.resid_kind = RESID_BF16;
  if ((rc = gemm_bf16(GEMM_NN, M, H, 3 * H, dqkv, 3 * H, w.w_qkv, H, e, 1, 0, st))) return rc;
  // all four weight gradients of the layer in ONE grouped launch (every operand is still live in the workspace):
  //   dW2 += d_y0'^T u ; dW1 += dz^T h ; dWo += d_a'^T ctx ; dWqkv += dqkv^T x     (reduction over the M token rows;
  //   primed = masked by the dropout of that dense output)
  {
    GroupedProblem q[4] = {
        {H, I, d_y0m, H, a.u, I, g.dw_2, I},
        {I, H, dz, I, a.h, H, g.dw_1, H},
        {3 * H, H, dqkv, 3 * H, x, H, g.dw_qkv, H},
        {H, H, d_am, H, a.ctx, H, g.dw_o, H},
    };
    static const int env_bn = [] { const char* v = getenv("VLB_WGRAD_BN"); return v ? atoi(v) : 256; }();
    static const int env_split = [] { const char* v = getenv("VLB_WGRAD_SPLIT"); return v ? atoi(v) : 2; }();
    if ((rc = gemm_grouped_tn(4, q, M, env_split, true, env_bn, st))) return rc;
  }
  count_launch(9);
  return VLB_OK;
}

}  // namespace vlb
