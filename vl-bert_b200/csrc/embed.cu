// Token / segment / position embedding gathers and the left-packing of the mixed-modality sequence
// [text tokens ; visual regions ; END ; pad]  (reference: common/visual_linguistic_bert.py:173-241 forward
// packing, :146-159 output un-packing).  Integer index math is bit-exact with the reference's boolean-mask
// assignment semantics: the k-th True of text_mask[b] lands at packed position k, the k-th True of
// object_mask[b] at text_end[b] + k, END at object_end[b], everything after is padding.
// Warp-per-row gathers with 128-bit accesses.
#include "common.cuh"

namespace vlb {

namespace {

// ------------------------------------------------------------------------------------------------
// index kernel: one block per sample.
//   kind[b,s]   0 text, 1 region, 2 END, 3 pad        src[b,s]  source column (text: t, region: r)
//   pos_id[b,s] position id (already offset by position_padding_idx + 1)
//   type_id[b,s] token type (text: text_token_type_ids, region/END: 2, pad: 0)
//   add_mask[b,s] 0 for s <= object_end else -10000   (visual_linguistic_bert.py:119-127,235)
//   obj_row[b,r] packed row (b*S + s) holding region r, or -1 when object_mask[b,r] is false  (un-packing)
//   lens[b] = {text_end, object_end};  err |= 1 if object_end + 1 > S (caller's max_length too small)
// ------------------------------------------------------------------------------------------------
__global__ void pack_index_kernel(const uint8_t* __restrict__ text_mask, const uint8_t* __restrict__ object_mask,
                                  const int64_t* __restrict__ text_type_ids, int T, int R, int S, int pos_offset,
                                  int32_t* __restrict__ kind, int32_t* __restrict__ src, int32_t* __restrict__ pos_id,
                                  int32_t* __restrict__ type_id, float* __restrict__ add_mask,
                                  int32_t* __restrict__ obj_row, int32_t* __restrict__ lens, int32_t* __restrict__ err) {
  extern __shared__ int32_t sh[];  // [T] text source by rank, [R] object source by rank
  int32_t* t_src = sh;
  int32_t* o_src = sh + T;
  __shared__ int32_t s_te, s_oe;
  const int b = blockIdx.x;
  const uint8_t* tm = text_mask + (size_t)b * T;
  const uint8_t* om = object_mask + (size_t)b * R;
  if (threadIdx.x == 0) {
    int te = 0;
    for (int i = 0; i < T; ++i) if (tm[i]) t_src[te++] = i;
    int ne = 0;
    for (int i = 0; i < R; ++i) if (om[i]) o_src[ne++] = i;
    s_te = te;
    s_oe = te + ne;
    lens[2 * b] = te;
    lens[2 * b + 1] = te + ne;
    if (te + ne + 1 > S) atomicOr(err, 1);
  }
  __syncthreads();
  const int te = s_te, oe = s_oe;
  for (int r = threadIdx.x; r < R; r += blockDim.x) obj_row[(size_t)b * R + r] = -1;
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    int k, sr = 0, pid = s, ty = 0;
    if (s < te) {
      k = 0; sr = t_src[s]; ty = (int)text_type_ids[(size_t)b * T + sr];
    } else if (s < oe) {
      k = 1; sr = o_src[s - te]; pid = te; ty = 2;
      obj_row[(size_t)b * R + sr] = b * S + s;
    } else if (s == oe) {
      k = 2; pid = te + 1; ty = 2;
    } else {
      k = 3;
    }
    const size_t o = (size_t)b * S + s;
    kind[o] = k; src[o] = sr; pos_id[o] = pid + pos_offset; type_id[o] = ty;
    add_mask[o] = (s <= oe) ? 0.0f : -10000.0f;
  }
}

__device__ __forceinline__ void add8(float (&a)[8], const float* p) {
  const float4 x = *reinterpret_cast<const float4*>(p);
  const float4 y = *reinterpret_cast<const float4*>(p + 4);
  a[0] += x.x; a[1] += x.y; a[2] += x.z; a[3] += x.w;
  a[4] += y.x; a[5] += y.y; a[6] += y.z; a[7] += y.w;
}

// ------------------------------------------------------------------------------------------------
// pack forward: e[b,s,:] = vl + position_embeddings[pos_id] + token_type_embeddings[type_id]  (fp32, pre-LayerNorm)
//   text  : vl = word_embeddings[ids[b,src]] + text_vis_ln[b,src]
//   region: vl = object_vl[b,src,H:2H] (linguistic half) + obj_vis_ln[b,src]
//   END   : vl = end_embedding[0]        pad: vl = 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pack_fwd_kernel(const int32_t* __restrict__ kind, const int32_t* __restrict__ src, const int32_t* __restrict__ pos_id,
                const int32_t* __restrict__ type_id, const int64_t* __restrict__ ids, const float* __restrict__ word_emb,
                const float* __restrict__ end_emb, const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                const float* __restrict__ text_vis_ln, const float* __restrict__ obj_vis_ln,
                const float* __restrict__ object_vl, int ld_obj, int lin_off, float* __restrict__ e, int B, int T, int R, int S,
                int H, int vocab, int max_pos, int32_t* __restrict__ err) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * S) return;
  const int b = row / S;
  const int k = kind[row], sr = src[row];
  int pid = pos_id[row];
  if (pid < 0 || pid >= max_pos) { if (lane == 0) atomicOr(err, 2); pid = 0; }
  const float* pe = pos_emb + (size_t)pid * H;
  int ty = type_id[row];
  if (ty < 0 || ty > 2) { if (lane == 0) atomicOr(err, 8); ty = 0; }   // token types: 0/1 text segments, 2 regions and [END]
  const float* te = type_emb + (size_t)ty * H;
  const float* a0 = nullptr;
  const float* a1 = nullptr;
  if (k == 0) {
    int64_t id = ids[(size_t)b * T + sr];
    if (id < 0 || id >= vocab) { if (lane == 0) atomicOr(err, 4); id = 0; }
    a0 = word_emb + (size_t)id * H;
    a1 = text_vis_ln + ((size_t)b * T + sr) * H;
  } else if (k == 1) {
    a0 = object_vl + ((size_t)b * R + sr) * ld_obj + lin_off;
    a1 = obj_vis_ln + ((size_t)b * R + sr) * H;
  } else if (k == 2) {
    a0 = end_emb;
  }
  for (int c = lane * 8; c < H; c += 256) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (a0) add8(a, a0 + c);
    if (a1) add8(a, a1 + c);
    add8(a, pe + c);
    add8(a, te + c);
    float* o = e + (size_t)row * H + c;
    *reinterpret_cast<float4*>(o) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(a[4], a[5], a[6], a[7]);
  }
}

__device__ __forceinline__ void atomic_add8(float* dst, const float (&v)[8]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// ------------------------------------------------------------------------------------------------
// pack backward: de [B*S, H] fp32 (gradient wrt the pre-LayerNorm sum) is scattered to
//   d_word_emb[ids] / d_end_emb / d_pos_emb[pos_id] / d_type_emb[type_id]      (fp32 atomics, "+=")
//   d_text_vl[b, src, :]  (dense [B*T, H], written once per referenced row; caller zero-fills)
//   d_obj_vl [b, src, :]  (dense [B*R, H])
// One block per (sample, 256-column slab); a thread owns one column and walks the sample's S rows.  The heavily shared
// destinations (the 3 token-type rows; the position row shared by all regions of a sample) are first accumulated
// per block -- types in registers, positions in shared memory [S+2][256] -- and flushed with one atomic per
// (row, column) per block, instead of one contended atomic per token.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_bwd_kernel(const int32_t* __restrict__ kind, const int32_t* __restrict__ src, const int32_t* __restrict__ pos_id,
                const int32_t* __restrict__ type_id, const int64_t* __restrict__ ids, const float* __restrict__ de,
                float* __restrict__ d_word, float* __restrict__ d_end, float* __restrict__ d_pos, float* __restrict__ d_type,
                float* __restrict__ d_text_vl, float* __restrict__ d_obj_vl, int T, int R, int S, int H, int vocab,
                int max_pos, int pos_offset) {
  extern __shared__ float pos_acc[];  // [S + 2][W], W = blockDim.x columns per block
  const int W = blockDim.x;
  const int b = blockIdx.x;
  const int c = blockIdx.y * W + threadIdx.x;
  const bool col_ok = c < H;
  for (int i = threadIdx.x; i < (S + 2) * W; i += W) pos_acc[i] = 0.0f;
  __syncthreads();
  float ty0 = 0.0f, ty1 = 0.0f, ty2 = 0.0f;
  for (int s = 0; s < S; ++s) {
    const int row = b * S + s;
    const int k = kind[row], sr = src[row];
    const float v = col_ok ? __ldg(de + (size_t)row * H + c) : 0.0f;
    int li = pos_id[row] - pos_offset;
    li = li < 0 ? 0 : (li > S + 1 ? S + 1 : li);
    pos_acc[li * W + threadIdx.x] += v;
    const int ty = type_id[row];
    ty0 += (ty == 0) ? v : 0.0f;
    ty1 += (ty == 1) ? v : 0.0f;
    ty2 += (ty == 2) ? v : 0.0f;
    if (!col_ok) continue;
    if (k == 0) {
      int64_t id = ids[(size_t)b * T + sr];
      if (id < 0 || id >= vocab) id = 0;
      if (d_word) atomicAdd(d_word + (size_t)id * H + c, v);
      if (d_text_vl) d_text_vl[((size_t)b * T + sr) * H + c] = v;
    } else if (k == 1) {
      if (d_obj_vl) d_obj_vl[((size_t)b * R + sr) * H + c] = v;
    } else if (k == 2) {
      if (d_end) atomicAdd(d_end + c, v);
    }
  }
  if (!col_ok) return;
  if (d_type) {
    if (ty0 != 0.0f) atomicAdd(d_type + c, ty0);
    if (ty1 != 0.0f) atomicAdd(d_type + H + c, ty1);
    if (ty2 != 0.0f) atomicAdd(d_type + 2 * H + c, ty2);
  }
  if (d_pos) {
    for (int li = 0; li < S + 2; ++li) {
      const float v = pos_acc[li * W + threadIdx.x];
      const int pid = li + pos_offset;
      if (v != 0.0f && pid >= 0 && pid < max_pos) atomicAdd(d_pos + (size_t)pid * H + c, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// row gather / scatter (output un-packing, visual_linguistic_bert.py:146-159, and the pad_sequence
// re-padding of region features, common/utils/pad_sequence.py:4-17):
//   gather : out[i, :] = idx[i] >= 0 ? in[idx[i], :] : 0
//   scatter: out[idx[i], :] (+)= in[i, :] for idx[i] >= 0          (each target row referenced at most once)
// IN/OUT element types: bf16 or fp32 (converted on the fly).
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(128)
gather_rows_kernel(const TI* __restrict__ in, int ld_in, const int32_t* __restrict__ idx, TO* __restrict__ out, int ld_out,
                   int n_out, int H) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_out) return;
  const int s = idx[row];
  for (int c = lane * 8; c < H; c += 256) {
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (s >= 0) {
      if constexpr (sizeof(TI) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(in + (size_t)s * ld_in + c);
        v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
        v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
      } else {
        add8(v, reinterpret_cast<const float*>(in) + (size_t)s * ld_in + c);
      }
    }
    if constexpr (sizeof(TO) == 2) {
      uint4 pk;
      pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
      pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(out + (size_t)row * ld_out + c) = pk;
    } else {
      float* o = reinterpret_cast<float*>(out) + (size_t)row * ld_out + c;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

template <typename TI>
__global__ void __launch_bounds__(128)
scatter_rows_add_f32_kernel(const TI* __restrict__ in, int ld_in, const int32_t* __restrict__ idx, float* __restrict__ out,
                            int ld_out, int n_in, int H) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_in) return;
  const int d = idx[row];
  if (d < 0) return;
  for (int c = lane * 8; c < H; c += 256) {
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if constexpr (sizeof(TI) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (size_t)row * ld_in + c);
      v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
      v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
    } else {
      add8(v, reinterpret_cast<const float*>(in) + (size_t)row * ld_in + c);
    }
    float* o = out + (size_t)d * ld_out + c;
    add8(v, o);
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

}  // namespace

int pack_index(const uint8_t* text_mask, const uint8_t* object_mask, const int64_t* text_type_ids, int B, int T, int R, int S,
               int pos_offset, int32_t* kind, int32_t* src, int32_t* pos_id, int32_t* type_id, float* add_mask,
               int32_t* obj_row, int32_t* lens, int32_t* err, cudaStream_t stream) {
  VLB_REQUIRE(text_mask && object_mask && text_type_ids && kind && src && pos_id && type_id && add_mask && obj_row && lens && err,
              "pack_index: null pointer");
  VLB_REQUIRE(B > 0 && T > 0 && R >= 0 && S > 0, "pack_index: bad sizes");
  pack_index_kernel<<<B, 128, (T + R + 1) * sizeof(int32_t), stream>>>(text_mask, object_mask, text_type_ids, T, R, S, pos_offset,
                                                                    kind, src, pos_id, type_id, add_mask, obj_row, lens, err);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int pack_forward(const int32_t* kind, const int32_t* src, const int32_t* pos_id, const int32_t* type_id, const int64_t* ids,
                 const float* word_emb, const float* end_emb, const float* pos_emb, const float* type_emb,
                 const float* text_vis_ln, const float* obj_vis_ln, const float* object_vl, int ld_obj, int lin_off, float* e,
                 int B, int T, int R, int S, int H, int vocab, int max_pos, int32_t* err, cudaStream_t stream) {
  VLB_REQUIRE(kind && src && pos_id && type_id && ids && word_emb && end_emb && pos_emb && type_emb && text_vis_ln && e && err,
              "pack_forward: null pointer");
  VLB_REQUIRE(H % 8 == 0 && ld_obj % 4 == 0 && lin_off % 4 == 0, "pack_forward: H must be a multiple of 8");
  const int rows = B * S;
  pack_fwd_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(kind, src, pos_id, type_id, ids, word_emb, end_emb, pos_emb, type_emb,
                                                      text_vis_ln, obj_vis_ln, object_vl, ld_obj, lin_off, e, B, T, R, S, H, vocab,
                                                      max_pos, err);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int pack_backward(const int32_t* kind, const int32_t* src, const int32_t* pos_id, const int32_t* type_id, const int64_t* ids,
                  const float* de, float* d_word, float* d_end, float* d_pos, float* d_type, float* d_text_vl, float* d_obj_vl,
                  int B, int T, int R, int S, int H, int vocab, int max_pos, int pos_offset, cudaStream_t stream) {
  VLB_REQUIRE(kind && src && pos_id && type_id && ids && de, "pack_backward: null pointer");
  const int W = ((S + 2) * 256 * (int)sizeof(float) <= 160 * 1024) ? 256 : 128;   // columns per block
  const int smem = (S + 2) * W * (int)sizeof(float);
  VLB_REQUIRE(smem <= 200 * 1024, "pack_backward: packed length %d too large", S);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(pack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  pack_bwd_kernel<<<dim3(B, (H + W - 1) / W), W, smem, stream>>>(kind, src, pos_id, type_id, ids, de, d_word, d_end, d_pos,
                                                                d_type, d_text_vl, d_obj_vl, T, R, S, H, vocab, max_pos, pos_offset);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int gather_rows(const void* in, int in_is_bf16, int ld_in, const int32_t* idx, void* out, int out_is_bf16, int ld_out, int n_out,
                int H, cudaStream_t stream) {
  VLB_REQUIRE(in && idx && out, "gather_rows: null pointer");
  VLB_REQUIRE(H % 8 == 0 && ld_in % (in_is_bf16 ? 8 : 4) == 0 && ld_out % (out_is_bf16 ? 8 : 4) == 0,
              "gather_rows: H must be a multiple of 8, ld of 8 (bf16) / 4 (f32)");
  if (n_out <= 0) return VLB_OK;
  const int grid = (n_out + 3) / 4;
  if (in_is_bf16 && out_is_bf16)
    gather_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), ld_in, idx, static_cast<__nv_bfloat16*>(out), ld_out, n_out, H);
  else if (in_is_bf16)
    gather_rows_kernel<__nv_bfloat16, float><<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), ld_in, idx,
                                                                      static_cast<float*>(out), ld_out, n_out, H);
  else if (out_is_bf16)
    gather_rows_kernel<float, __nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<const float*>(in), ld_in, idx,
                                                                      static_cast<__nv_bfloat16*>(out), ld_out, n_out, H);
  else
    gather_rows_kernel<float, float><<<grid, 128, 0, stream>>>(static_cast<const float*>(in), ld_in, idx,
                                                              static_cast<float*>(out), ld_out, n_out, H);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

int scatter_rows_add(const void* in, int in_is_bf16, int ld_in, const int32_t* idx, float* out, int ld_out, int n_in, int H,
                     cudaStream_t stream) {
  VLB_REQUIRE(in && idx && out, "scatter_rows_add: null pointer");
  VLB_REQUIRE(H % 8 == 0 && ld_in % (in_is_bf16 ? 8 : 4) == 0 && ld_out % 4 == 0,
              "scatter_rows_add: H must be a multiple of 8, ld of 8 (bf16) / 4 (f32)");
  if (n_in <= 0) return VLB_OK;
  const int grid = (n_in + 3) / 4;
  if (in_is_bf16)
    scatter_rows_add_f32_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), ld_in, idx, out,
                                                                        ld_out, n_in, H);
  else
    scatter_rows_add_f32_kernel<float><<<grid, 128, 0, stream>>>(static_cast<const float*>(in), ld_in, idx, out, ld_out, n_in, H);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

}  // namespace vlb
