// tcgen05 GEMM for sm_100a: persistent, warp-specialised (TMA producer / single-thread MMA issuer /
// 4 epilogue warps), 128 x BN x 64 tiles, bf16 operands staged by TMA into 128B-swizzled shared
// memory, fp32 accumulators double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Fused epilogues cover everything the VL-BERT encoder layer needs around its GEMMs
// (reference: external/pytorch_pretrained_bert/modeling.py:291-293 QKV, :330-333 output dense +
// residual, :362-363 intermediate dense + erf-GELU, :375-378 output dense + residual) and their
// backward passes (dgrad with fused GELU' / residual-gradient add, wgrad with split-K fp32
// reduction).
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "gemm_sm100.cuh"

namespace vlb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 8;        // two warps per TMEM lane quarter, each takes every other 32-column chunk
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;  // warp0 = TMA, warp1 = MMA + TMEM alloc, warps 2..9 = epilogue
constexpr int STAGE_F32_PER_WARP = 32 * 32;        // 4 KB transposition buffer per epilogue warp

// Compile-time epilogue variants for the hot launches of the encoder layer (0 = generic, flags read at run time).
enum : int { EPI_GENERIC = 0, EPI_BIAS_BF16, EPI_BIAS_RESID16_F32, EPI_BIAS_GELU_AUX_BF16, EPI_DGELU_BF16, EPI_RESID16_BF16,
             EPI_ATOMIC_F32, EPI_CONV_RELU_BF16, EPI_CONV_RESID_RELU_BF16, EPI_PLAIN_BF16, EPI_BIAS_DROP_RESID16_F32, EPI_BIAS_RESIDLN_F32,
             EPI_BIAS_DROP_RESIDLN_F32, EPI_NUM };
struct EpiTraitsBase { static constexpr bool kStatic = true; static constexpr bool bias = false, cscale = false, drop = false; static constexpr int act = ACT_NONE, resid = RESID_NONE, out = OUT_BF16; };
template <int EPI> struct EpiTraits : EpiTraitsBase { static constexpr bool kStatic = false; };
template <> struct EpiTraits<EPI_BIAS_BF16>          : EpiTraitsBase { static constexpr bool bias = true; };
template <> struct EpiTraits<EPI_BIAS_RESID16_F32>   : EpiTraitsBase { static constexpr bool bias = true; static constexpr int resid = RESID_BF16, out = OUT_F32; };
// dense + bias -> dropout -> + residual (BertSelfOutput / BertOutput in training mode, modeling.py:330-333,375-378)
template <> struct EpiTraits<EPI_BIAS_DROP_RESID16_F32> : EpiTraitsBase { static constexpr bool bias = true, drop = true; static constexpr int resid = RESID_BF16, out = OUT_F32; };
// the same with the residual stream in fp32 (LayerNorm output recomputed from its stored input, or a plain fp32 tensor)
template <> struct EpiTraits<EPI_BIAS_RESIDLN_F32> : EpiTraitsBase { static constexpr bool bias = true; static constexpr int resid = RESID_LN_F32, out = OUT_F32; };
template <> struct EpiTraits<EPI_BIAS_DROP_RESIDLN_F32> : EpiTraitsBase { static constexpr bool bias = true, drop = true; static constexpr int resid = RESID_LN_F32, out = OUT_F32; };
template <> struct EpiTraits<EPI_BIAS_GELU_AUX_BF16> : EpiTraitsBase { static constexpr bool bias = true; static constexpr int act = ACT_GELU; };
template <> struct EpiTraits<EPI_DGELU_BF16>         : EpiTraitsBase { static constexpr int act = ACT_DGELU_MUL; };
template <> struct EpiTraits<EPI_RESID16_BF16>       : EpiTraitsBase { static constexpr int resid = RESID_BF16; };
template <> struct EpiTraits<EPI_ATOMIC_F32>         : EpiTraitsBase { static constexpr int out = OUT_F32_ATOMIC; };
// convolution + frozen BatchNorm: per-channel scale and shift, then ReLU / residual add + ReLU (Bottleneck, resnet.py:98-118)
template <> struct EpiTraits<EPI_CONV_RELU_BF16>       : EpiTraitsBase { static constexpr bool bias = true, cscale = true; static constexpr int act = ACT_RELU; };
template <> struct EpiTraits<EPI_CONV_RESID_RELU_BF16> : EpiTraitsBase { static constexpr bool bias = true, cscale = true; static constexpr int act = ACT_RELU_POST, resid = RESID_BF16; };
template <> struct EpiTraits<EPI_PLAIN_BF16>           : EpiTraitsBase {};

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int split_k, kb_per_split;
  int num_items;
  // descriptor geometry (bytes); see make_smem_desc_sw128
  uint32_t a_lbo, a_sbo, a_kadv;
  uint32_t b_lbo, b_sbo, b_kadv;
  GemmEpilogue e;
  // implicit convolution operand (TMA im2col mode): side 0 none; 1 = A of an NT GEMM (rows = output pixels, K = taps x channels);
  // 2 = B of a TN GEMM (reduction rows = output pixels, N = taps x channels)
  int cv_side;
  int cv_C, cv_Ho, cv_Wo, cv_stride, cv_lower, cv_dil, cv_kw;
  // stream-K tail: the tiles of the last, partial round of the persistent grid are split along K into sk_chunks work units
  // each; units add their partial accumulators into an fp32 scratch tile, the last unit to arrive applies the epilogue.
  int sk_full_items;      // items [0, sk_full_items) are whole tiles; 0 chunks = feature off
  int sk_chunks, sk_kb_per_chunk;
  float* sk_scratch;      // [tail tiles][BM][BN] fp32, all zero between launches
  int* sk_counters;       // [tail tiles], all zero between launches
  int sk_debug;           // timing experiments only (VLB_SK_DEBUG): 1 = skip the partial reds, 2 = skip the fix-up pass
  // Tail split: the tiles of the last, partial round of the persistent grid are cut along N into `tail_split` units of
  // BN / tail_split (= 64) columns each, so that the partial round is spread over every CTA instead of leaving most SMs idle
  // for a whole tile time.  Items [0, tail_first) are whole tiles; item tail_first + u is unit u % tail_split of tile
  // tail_first + u / tail_split.  0 = off.
  int tail_first, tail_split;
  int epi_prefetch;       // 1: epilogue warps L2-prefetch the tile's residual / aux input while its mainloop runs
  // Dynamic tile scheduling: sched -> {next-item counter, finished-CTA counter} in global memory (both zero between launches).
  // Every CTA starts with item = its index and then takes items nworkers + atomicAdd(counter, 1): a CTA whose start is delayed
  // (e.g. its SM is still held by an NCCL channel CTA of an overlapped gradient all-reduce) simply takes fewer items instead of
  // holding its statically assigned share back; the last CTA to finish resets both counters.  nullptr = static round robin.
  int* sched;
  unsigned long long* trace;   // timing aid (vlb_debug_gemm_trace): per CTA 8 items x 4 globaltimer stamps, nullptr = off
  int roles_high;              // 1: the TMA producer and the MMA issuer are the two HIGHEST hardware warps of the CTA (see gemm_body)
};

constexpr int SCHED_DEPTH = 4;   // ring of item indices handed from the producer warp to the MMA / epilogue warps

// linear output-pixel index -> base pixel (w, h, n) of the im2col traversal
struct PixelCoord { int w, h, n; };
__device__ __forceinline__ PixelCoord conv_base_pixel(const GemmParams& p, int pix) {
  const int hw = p.cv_Ho * p.cv_Wo;
  PixelCoord c;
  c.n = pix / hw;
  const int rem = pix - c.n * hw;
  const int ho = rem / p.cv_Wo;
  c.h = ho * p.cv_stride + p.cv_lower;
  c.w = (rem - ho * p.cv_Wo) * p.cv_stride + p.cv_lower;
  return c;
}

// CG2 = CTA-pair mode (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BN tile; each CTA stages its own 128
// rows of A and one HALF of the B tile, so the B operand crosses the L2->SM fabric once per pair instead of once per CTA.
// Grouped launch: up to MAX_GROUP independent problems with the same reduction length K share one persistent grid
// (the four weight-gradient GEMMs of a layer: 432 equal tiles -> 3 full rounds instead of 4 launches of 1-2 ragged rounds).
constexpr int MAX_GROUP = 4;
struct GroupTable {
  CUtensorMap ta[MAX_GROUP];
  CUtensorMap tb[MAX_GROUP];
  int count;
  int item_begin[MAX_GROUP + 1];   // prefix sums of per-problem items (= m_blocks * n_blocks * splits)
  int m_blocks[MAX_GROUP], n_blocks[MAX_GROUP];
  int M[MAX_GROUP], N[MAX_GROUP];
  void* out[MAX_GROUP];
  int ldo[MAX_GROUP];
};

// Stream-K tail (experimental, compiled out by default): measured on B200 it does not pay for this workload -- the partial
// reds + arrival fence + fix-up pass cost ~8-12 us per launch, more than the partial round they remove (profiles/r01_streamk_*.log).
#ifndef VLB_ENABLE_STREAMK
#define VLB_ENABLE_STREAMK 0
#endif

struct ItemCoord {
  int g, split, m_blk, n_blk;
  int kb_begin, kb_end;   // k-block range of this work unit
  int tail;               // >= 0: index of the stream-K tail tile this unit is a K-chunk of
  int n0, bn;             // first column and width of this work unit (bn < BN for the N-split units of the last round)
};
template <bool GROUPED, int BN>
__device__ __forceinline__ ItemCoord decode_item(int item, const GemmParams& p, const GroupTable& gt) {
  ItemCoord c;
  c.g = 0;
  c.bn = BN;
  if (!GROUPED && p.tail_split > 1 && item >= p.tail_first) {
    const int u = item - p.tail_first;
    const int tile = p.tail_first + u / p.tail_split;
    const int sub = u - (u / p.tail_split) * p.tail_split;
    c.tail = -1;
    c.split = 0;
    c.m_blk = tile / p.num_n_blocks;
    c.n_blk = tile - c.m_blk * p.num_n_blocks;
    c.kb_begin = 0;
    c.kb_end = p.num_k_blocks;
    c.bn = BN / p.tail_split;
    c.n0 = c.n_blk * BN + sub * c.bn;
    return c;
  }
  int local = item, mb = p.num_m_blocks, nb = p.num_n_blocks;
  if (GROUPED) {
#pragma unroll
    for (int i = 1; i < MAX_GROUP; ++i)
      if (i < gt.count && item >= gt.item_begin[i]) c.g = i;
    local = item - gt.item_begin[c.g];
    mb = gt.m_blocks[c.g];
    nb = gt.n_blocks[c.g];
  }
  c.tail = -1;
  if (VLB_ENABLE_STREAMK && !GROUPED && p.sk_chunks > 1 && item >= p.sk_full_items) {
    const int u = item - p.sk_full_items;
    c.tail = u / p.sk_chunks;
    c.split = u - c.tail * p.sk_chunks;
    const int tile = p.sk_full_items + c.tail;
    c.m_blk = tile / nb;
    c.n_blk = tile - c.m_blk * nb;
    c.kb_begin = c.split * p.sk_kb_per_chunk;
    c.kb_end = min(p.num_k_blocks, c.kb_begin + p.sk_kb_per_chunk);
    c.n0 = c.n_blk * BN;
    return c;
  }
  const int per_split = mb * nb;
  c.split = local / per_split;
  const int rem = local - c.split * per_split;
  c.m_blk = rem / nb;
  c.n_blk = rem - c.m_blk * nb;
  c.kb_begin = c.split * p.kb_per_split;
  c.kb_end = min(p.num_k_blocks, c.kb_begin + p.kb_per_split);
  c.n0 = c.n_blk * BN;
  return c;
}

// Cluster modes (CM): 0 = independent CTAs; 1 = CG2 pair MMA (above); 2 = MC2: a cluster of two CTAs works on two
// vertically adjacent 128 x BN tiles that need the SAME B tile; each CTA fetches one half of it and TMA-multicasts it into
// both CTAs' shared memory (B crosses the L2->SM fabric once per cluster), the MMAs stay independent cta_group::1.
template <int BN, int CM = 0, int EW = EPI_WARPS>
struct Cfg {
  static constexpr bool CG2 = CM == 1;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = CG2 ? BN / 2 : BN;   // rows (n) of the B tile present in this CTA's shared memory
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CG2 ? (BN == 256 ? 6 : 8) : ((BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8)));
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int EPI_STAGE_BYTES = EW * STAGE_F32_PER_WARP * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers + scheduler ring*/;
  static constexpr int THREADS = 64 + EW * 32;
  static_assert(SMEM_BYTES <= 227 * 1024, "tile configuration exceeds the shared memory of an SM");
};

// Epilogue of one 32-row x 32-column accumulator chunk owned by a warp.
//   tcgen05.ld gives thread t row t (32 consecutive columns).  Writing those rows straight to global memory would make
//   every warp store touch 32 different cache lines, so the chunk is first transposed through a 4 KB XOR-swizzled
//   shared-memory buffer: afterwards lane (r = lane/8, g = lane%8) owns 4 consecutive columns of rows r, r+4, ...,
//   and all global loads (bias, residual, saved pre-activation) and stores are row-contiguous.
//   FROM_SCRATCH (stream-K fix-up): the accumulator chunk is not in registers but in an fp32 scratch tile in global memory
//   (sk_src -> its first row / column, row stride sk_ld); it is read -- and reset to zero -- directly in the row-contiguous
//   lane layout of phase 2, so those accesses are coalesced too.
template <int EPI, bool FROM_SCRATCH = false>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& e, const uint32_t (&v)[32], float* stage, int lane,
                                               int row_base, int col0, int M, int N, float* sk_src = nullptr, int sk_ld = 0,
                                               DropState dstate = DropState{0ull, 0u}) {
  using T = EpiTraits<EPI>;
  const bool has_bias = T::kStatic ? T::bias : (e.bias != nullptr);
  const int act = T::kStatic ? T::act : e.act;
  const int resid_kind = T::kStatic ? T::resid : e.resid_kind;
  const int out_kind = T::kStatic ? T::out : e.out_kind;
  const bool need_aux_in = (act == ACT_DGELU_MUL || act == ACT_DRELU_MUL);
  const bool has_drop = T::kStatic ? T::drop : (e.drop.thresh != 0u);
  const int g = lane & 7;
  const int col = col0 + g * 4;
  const bool col_ok = col < N;
  const int r0 = lane >> 3;
  // ---- phase 0: issue every global load of this chunk up front (8 independent requests per lane in flight) ----
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has_bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  float4 scale4 = make_float4(e.alpha, e.alpha, e.alpha, e.alpha);
  if ((T::kStatic ? T::cscale : (e.colscale != nullptr)) && col_ok) {
    const float4 cs = __ldg(reinterpret_cast<const float4*>(e.colscale + col));
    scale4 = make_float4(cs.x * e.alpha, cs.y * e.alpha, cs.z * e.alpha, cs.w * e.alpha);
  }
  uint2 in16[8];   // bf16 residual or saved pre-activation
  float4 in32[8];  // fp32 residual (generic path only)
  if (resid_kind == RESID_BF16 || need_aux_in) {
    const __nv_bfloat16* __restrict__ src = reinterpret_cast<const __nv_bfloat16*>(need_aux_in ? e.aux : e.resid);
    const int ld = need_aux_in ? e.ld_aux : e.ldr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + r0;
      in16[i] = (row < M && col_ok) ? __ldg(reinterpret_cast<const uint2*>(src + (size_t)row * ld + col)) : make_uint2(0u, 0u);
    }
  }
  uint32_t kw[8];   // dropout keep flags: the word holding this lane's four columns, one per row (issued with the other loads)
  if (has_drop) {
    const int wpr = (N + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + r0;
      kw[i] = (row < M && col_ok) ? __ldg(e.drop.bits + (size_t)row * wpr + (col >> 5)) : 0u;
    }
  }
  float ln_mu[8], ln_rs[8];
  float4 ln_g4 = make_float4(1.f, 1.f, 1.f, 1.f), ln_b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool ln_resid = resid_kind == RESID_LN_F32 && e.ln_mean != nullptr;
  if (resid_kind == RESID_F32 || resid_kind == RESID_LN_F32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + r0;
      in32[i] = (row < M && col_ok) ? __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.resid) + (size_t)row * e.ldr + col))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      if (ln_resid) {
        ln_mu[i] = row < M ? __ldg(e.ln_mean + row) : 0.0f;
        ln_rs[i] = row < M ? __ldg(e.ln_rstd + row) : 0.0f;
      }
    }
    if (ln_resid && col_ok) {
      ln_g4 = __ldg(reinterpret_cast<const float4*>(e.ln_gamma + col));
      ln_b4 = __ldg(reinterpret_cast<const float4*>(e.ln_beta + col));
    }
  }
  // ---- phase 1: transpose the accumulator chunk through shared memory ----
  if (FROM_SCRATCH) {
    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)   // all eight loads in flight before anything depends on them (one L2 round trip per chunk)
      acc[i] = __ldcg(reinterpret_cast<const float4*>(sk_src + (size_t)(i * 4 + r0) * sk_ld + g * 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + r0;
      __stcg(reinterpret_cast<float4*>(sk_src + (size_t)r * sk_ld + g * 4), make_float4(0.f, 0.f, 0.f, 0.f));   // scratch returns to zero
      reinterpret_cast<float4*>(stage + r * 32)[g ^ (r & 7)] = acc[i];
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(stage + lane * 32);
#pragma unroll
    for (int gg = 0; gg < 8; ++gg)
      dst[gg ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * gg]), __uint_as_float(v[4 * gg + 1]),
                                         __uint_as_float(v[4 * gg + 2]), __uint_as_float(v[4 * gg + 3]));
  }
  __syncwarp();
  // ---- phase 2: math + row-contiguous stores ----
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + r0;
    const int row = row_base + r;
    const float4 a = reinterpret_cast<const float4*>(stage + r * 32)[g ^ (r & 7)];
    if (row < M && col_ok) {
      float x[4] = {fmaf(a.x, scale4.x, bias4.x), fmaf(a.y, scale4.y, bias4.y), fmaf(a.z, scale4.z, bias4.z), fmaf(a.w, scale4.w, bias4.w)};
      if (act == ACT_GELU) {
        // x = gelu(z); aux receives gelu'(z) (bf16), computed from the same erf / exp -- the backward dgrad epilogue then only
        // multiplies (ACT_DGELU_MUL) instead of re-evaluating erf and exp for every element.
        float gp[4];
#pragma unroll
        for (int j = 0; j < 4; j += 2) {   // two elements per packed-fp32 instruction
          float2 cdf, pdf;
          const float2 xx = make_float2(x[j], x[j + 1]);
          gelu_parts2(xx, cdf, pdf);
          const float2 g2 = __ffma2_rn(xx, pdf, cdf);
          const float2 o2 = __fmul2_rn(xx, cdf);
          gp[j] = g2.x; gp[j + 1] = g2.y;
          x[j] = o2.x; x[j + 1] = o2.y;
        }
        if (e.aux != nullptr) {
          uint2 z;
          z.x = pack_bf16x2(gp[0], gp[1]);
          z.y = pack_bf16x2(gp[2], gp[3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.aux) + (size_t)row * e.ld_aux + col) = z;
        }
      } else if (act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.0f);
      } else if (need_aux_in) {
        const float zz[4] = {bf16lo(in16[i].x), bf16hi(in16[i].x), bf16lo(in16[i].y), bf16hi(in16[i].y)};
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = (act == ACT_DGELU_MUL) ? x[j] * zz[j] : (zz[j] > 0.0f ? x[j] : 0.0f);
      }
      if (has_drop) drop4_bits(x, (kw[i] >> (col & 31)) & 0xFu, e.drop.scale);
      if (resid_kind == RESID_BF16) {
        x[0] += bf16lo(in16[i].x); x[1] += bf16hi(in16[i].x); x[2] += bf16lo(in16[i].y); x[3] += bf16hi(in16[i].y);
      } else if (resid_kind == RESID_F32) {
        x[0] += in32[i].x; x[1] += in32[i].y; x[2] += in32[i].z; x[3] += in32[i].w;
      } else if (resid_kind == RESID_LN_F32) {
        if (ln_resid) {   // residual = LayerNorm(resid row) in fp32, same formula as layernorm_fwd_kernel
          x[0] += fmaf(ln_g4.x, (in32[i].x - ln_mu[i]) * ln_rs[i], ln_b4.x);
          x[1] += fmaf(ln_g4.y, (in32[i].y - ln_mu[i]) * ln_rs[i], ln_b4.y);
          x[2] += fmaf(ln_g4.z, (in32[i].z - ln_mu[i]) * ln_rs[i], ln_b4.z);
          x[3] += fmaf(ln_g4.w, (in32[i].w - ln_mu[i]) * ln_rs[i], ln_b4.w);
        } else {
          x[0] += in32[i].x; x[1] += in32[i].y; x[2] += in32[i].z; x[3] += in32[i].w;
        }
      }
      if (act == ACT_RELU_POST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.0f);
      }
      if (out_kind == OUT_BF16) {
        uint2 o;
        o.x = pack_bf16x2(x[0], x[1]);
        o.y = pack_bf16x2(x[2], x[3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)row * e.ldo + col) = o;
        if (e.colsum != nullptr) {  // sum what was actually stored (bf16-rounded), like a separate column-sum pass would
          csum[0] += bf16lo(o.x); csum[1] += bf16hi(o.x); csum[2] += bf16lo(o.y); csum[3] += bf16hi(o.y);
        }
      } else {
        float* op = reinterpret_cast<float*>(e.out) + (size_t)row * e.ldo + col;
        if (out_kind == OUT_F32) {
          *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
        }
        if (e.colsum != nullptr) { csum[0] += x[0]; csum[1] += x[1]; csum[2] += x[2]; csum[3] += x[3]; }
      }
    }
  }
  if (e.colsum != nullptr) {
    // lanes with equal (lane & 7) own the same 4 columns: reduce over the 4 row groups, then one atomic per column
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], 8);
      csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], 16);
    }
    if (lane < 8 && col_ok)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(e.colsum + col), "f"(csum[0]), "f"(csum[1]), "f"(csum[2]), "f"(csum[3]) : "memory");
  }
  __syncwarp();
}

// Interior fast path of epilogue_chunk for the encoder's compile-time variants: the chunk lies completely inside the matrix
// (no per-row / per-column predicates), row addresses advance by a constant stride, and the fp32 arithmetic runs on the
// packed pipe (FFMA2 / FMUL2 / FADD2: two results per issue slot) -- the epilogue warps are issue-bound (two warps per
// scheduler, ~850 instructions per chunk in the general routine), and in a single-wave launch the epilogue is not hidden
// behind any mainloop.  Same layout and same results as epilogue_chunk up to fp32 rounding of re-associated scalings.
template <int EPI>
struct EpiFast {
  using T = EpiTraits<EPI>;
  static constexpr bool value = T::kStatic && !T::cscale && T::out != OUT_F32_ATOMIC &&
                                (T::act == ACT_NONE || T::act == ACT_GELU || T::act == ACT_DGELU_MUL) &&
                                (T::resid == RESID_NONE || T::resid == RESID_LN_F32);
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// Operands of one chunk that were staged in shared memory ahead of time (pair kernel, last tile of a CTA: the idle
// operand ring receives the whole tile's residual rows, keep-flag words and LayerNorm statistics with every load in flight at
// once -- otherwise each chunk starts with its own round trip to L2 / DRAM and the exposed epilogue is a chain of them).
//   in32: this lane's eight residual vectors, entry i at in32[i * 32];  kw / mu / rs: one entry per row of the warp.
struct StagedChunk {
  const float4* in32 = nullptr;
  const uint32_t* kw = nullptr;
  const float* mu = nullptr;
  const float* rs = nullptr;
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk_fast(const GemmEpilogue& e, const uint32_t (&v)[32], float* stage, int lane,
                                                    int row_base, int col0, int N, const StagedChunk sc_in = StagedChunk()) {
  using T = EpiTraits<EPI>;
  constexpr bool aux_in = T::act == ACT_DGELU_MUL;
  constexpr bool colsum_ok = aux_in;   // the fused bias-gradient column sum exists only on the GELU' epilogue (db_1); callers check
  const int g = lane & 7, r0 = lane >> 3;
  const int col = col0 + g * 4;
  const size_t row = (size_t)(row_base + r0);
  // ---- phase 0: every global load of the chunk in flight before anything depends on it ----
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (T::bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  uint2 in16[8];
  float4 in32[8];
  uint32_t kw[8];
  float ln_mu[8], ln_rs[8];
  float4 ln_g4 = make_float4(1.f, 1.f, 1.f, 1.f), ln_b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool ln_resid = T::resid == RESID_LN_F32 && e.ln_mean != nullptr;
  if (aux_in) {
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(e.aux) + row * (size_t)e.ld_aux + col;
    const size_t step = (size_t)e.ld_aux * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i, src += step) in16[i] = __ldg(reinterpret_cast<const uint2*>(src));
  }
  const bool staged = sc_in.in32 != nullptr;
  if (T::resid == RESID_LN_F32) {
    if (staged) {
#pragma unroll
      for (int i = 0; i < 8; ++i) in32[i] = sc_in.in32[i * 32];
    } else {
      const float* src = reinterpret_cast<const float*>(e.resid) + row * (size_t)e.ldr + col;
      const size_t step = (size_t)e.ldr * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i, src += step) in32[i] = __ldg(reinterpret_cast<const float4*>(src));
    }
    if (ln_resid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ln_mu[i] = staged ? sc_in.mu[r0 + i * 4] : __ldg(e.ln_mean + row + i * 4);
        ln_rs[i] = staged ? sc_in.rs[r0 + i * 4] : __ldg(e.ln_rstd + row + i * 4);
      }
      ln_g4 = __ldg(reinterpret_cast<const float4*>(e.ln_gamma + col));
      ln_b4 = __ldg(reinterpret_cast<const float4*>(e.ln_beta + col));
    }
  }
  if (T::drop) {   // the word holding this chunk's 32 keep flags of each row; this lane's four columns are bits 4g .. 4g+3
    if (staged) {
#pragma unroll
      for (int i = 0; i < 8; ++i) kw[i] = sc_in.kw[r0 + i * 4];
    } else {
      const size_t wpr = (size_t)((N + 31) >> 5);
      const uint32_t* src = e.drop.bits + row * wpr + (size_t)(col0 >> 5);
#pragma unroll
      for (int i = 0; i < 8; ++i, src += 4 * wpr) kw[i] = __ldg(src);
    }
  }
  // ---- phase 1: transpose the accumulator chunk through shared memory ----
  {
    float4* dst = reinterpret_cast<float4*>(stage + lane * 32);
#pragma unroll
    for (int gg = 0; gg < 8; ++gg)
      dst[gg ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * gg]), __uint_as_float(v[4 * gg + 1]),
                                         __uint_as_float(v[4 * gg + 2]), __uint_as_float(v[4 * gg + 3]));
  }
  __syncwarp();
  // ---- phase 2: math + row-contiguous stores ----
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  const float ds = T::drop ? e.drop.scale : 1.0f;
  const float2 sc = f2(ds, ds);
  const float2 b01 = f2(bias4.x * ds, bias4.y * ds), b23 = f2(bias4.z * ds, bias4.w * ds);   // bias pre-scaled by 1 / (1 - p)
  char* outp = reinterpret_cast<char*>(e.out) + (row * (size_t)e.ldo + col) * (T::out == OUT_BF16 ? 2 : 4);
  const size_t out_step = (size_t)e.ldo * 4 * (T::out == OUT_BF16 ? 2 : 4);
  __nv_bfloat16* auxp = (T::act == ACT_GELU && e.aux != nullptr) ? reinterpret_cast<__nv_bfloat16*>(e.aux) + row * (size_t)e.ld_aux + col : nullptr;
  const size_t aux_step = (size_t)e.ld_aux * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + r0;
    const float4 a = reinterpret_cast<const float4*>(stage + r * 32)[g ^ (r & 7)];
    float2 x01, x23;
    if (T::bias || T::drop) {
      x01 = __ffma2_rn(f2(a.x, a.y), sc, b01);
      x23 = __ffma2_rn(f2(a.z, a.w), sc, b23);
    } else {
      x01 = f2(a.x, a.y);
      x23 = f2(a.z, a.w);
    }
    if (T::act == ACT_GELU) {
      float2 cdf, pdf;
      gelu_parts2(x01, cdf, pdf);
      const float2 g01 = __ffma2_rn(x01, pdf, cdf);
      x01 = __fmul2_rn(x01, cdf);
      gelu_parts2(x23, cdf, pdf);
      const float2 g23 = __ffma2_rn(x23, pdf, cdf);
      x23 = __fmul2_rn(x23, cdf);
      if (auxp != nullptr) {
        uint2 z;
        z.x = pack_bf16x2(g01.x, g01.y);
        z.y = pack_bf16x2(g23.x, g23.y);
        *reinterpret_cast<uint2*>(auxp) = z;
        auxp += aux_step;
      }
    } else if (aux_in) {
      x01 = __fmul2_rn(x01, f2(bf16lo(in16[i].x), bf16hi(in16[i].x)));
      x23 = __fmul2_rn(x23, f2(bf16lo(in16[i].y), bf16hi(in16[i].y)));
    }
    if (T::drop) {
      const uint32_t k4 = kw[i] >> (g * 4);
      x01.x = (k4 & 1u) ? x01.x : 0.0f;
      x01.y = (k4 & 2u) ? x01.y : 0.0f;
      x23.x = (k4 & 4u) ? x23.x : 0.0f;
      x23.y = (k4 & 8u) ? x23.y : 0.0f;
    }
    if (T::resid == RESID_LN_F32) {
      if (ln_resid) {   // residual = LayerNorm(resid row) in fp32, same formula as layernorm_fwd_kernel
        const float2 nm = f2(-ln_mu[i], -ln_mu[i]), rs2 = f2(ln_rs[i], ln_rs[i]);
        const float2 h01 = __fmul2_rn(__fadd2_rn(f2(in32[i].x, in32[i].y), nm), rs2);
        const float2 h23 = __fmul2_rn(__fadd2_rn(f2(in32[i].z, in32[i].w), nm), rs2);
        x01 = __fadd2_rn(x01, __ffma2_rn(f2(ln_g4.x, ln_g4.y), h01, f2(ln_b4.x, ln_b4.y)));
        x23 = __fadd2_rn(x23, __ffma2_rn(f2(ln_g4.z, ln_g4.w), h23, f2(ln_b4.z, ln_b4.w)));
      } else {
        x01 = __fadd2_rn(x01, f2(in32[i].x, in32[i].y));
        x23 = __fadd2_rn(x23, f2(in32[i].z, in32[i].w));
      }
    }
    if (T::out == OUT_BF16) {
      uint2 o;
      o.x = pack_bf16x2(x01.x, x01.y);
      o.y = pack_bf16x2(x23.x, x23.y);
      *reinterpret_cast<uint2*>(outp) = o;
      if (colsum_ok && e.colsum != nullptr) {  // sum what was actually stored (bf16-rounded), like a separate column-sum pass would
        csum[0] += bf16lo(o.x); csum[1] += bf16hi(o.x); csum[2] += bf16lo(o.y); csum[3] += bf16hi(o.y);
      }
    } else {
      *reinterpret_cast<float4*>(outp) = make_float4(x01.x, x01.y, x23.x, x23.y);
    }
    outp += out_step;
  }
  if (colsum_ok && e.colsum != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], 8);
      csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], 16);
    }
    if (lane < 8)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(e.colsum + col), "f"(csum[0]), "f"(csum[1]), "f"(csum[2]), "f"(csum[3]) : "memory");
  }
  __syncwarp();
}
// whether the fast path may take this launch's chunks (run-time part of the decision)
template <int EPI>
__device__ __forceinline__ bool epi_fast_ok(const GemmEpilogue& e) {
  return EpiFast<EPI>::value && e.alpha == 1.0f && (e.colsum == nullptr || EpiTraits<EPI>::act == ACT_DGELU_MUL);
}

// EW = number of epilogue warps (a multiple of 4: EW / 4 warps share each TMEM lane quarter and take every (EW / 4)-th
// 32-column chunk).  8 everywhere except the GELU launch, whose epilogue -- not its mainloop -- sets the period of a tile with
// two warps per scheduler (6.4-7.7 us against 5.4 us on a 128 x 256 tile): there 16 warps work on 128 x 192 tiles.
template <int BN, bool A_MN, bool B_MN, int EPI, int CM, bool GROUPED, int EW = EPI_WARPS>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tma_a, const CUtensorMap& tma_b, const GemmParams& p,
                                          const GroupTable& gt, const CUtensorMap& tma_b_tail) {
  using C = Cfg<BN, CM, EW>;
  static_assert(EW % 4 == 0 && (EW == EPI_WARPS || !VLB_ENABLE_STREAMK), "epilogue warps come in groups of four");
  constexpr bool CG2 = CM == 1;   // pair MMA
  constexpr bool MC2 = CM == 2;   // independent MMAs, B tile multicast
  constexpr bool CL = CM != 0;    // any 2-CTA cluster mode
  const uint32_t rank = CL ? cluster_ctarank() : 0u;        // CTA rank inside the cluster
  const bool leader = MC2 || rank == 0;                     // CG2: the leader issues the MMAs for both CTAs
  const int worker = CL ? (blockIdx.x >> 1) : blockIdx.x;   // persistent work-loop index (a cluster shares it)
  const int nworkers = CL ? (gridDim.x >> 1) : gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * C::STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  volatile uint32_t* sk_flag = tmem_slot + 2;   // stream-K: "this CTA delivered the last K-chunk of the tile"
  uint64_t* sched_full = bars + 2 * C::STAGES + 4 + 2;    // [SCHED_DEPTH] producer -> consumers: item index published
  uint64_t* sched_empty = sched_full + SCHED_DEPTH;       // [SCHED_DEPTH] consumers (MMA thread + epilogue warps) -> producer
  volatile int* sched_item = reinterpret_cast<volatile int*>(sched_empty + SCHED_DEPTH);
  const bool dyn = !CL && p.sched != nullptr;

  // Role = logical warp index: 0 = TMA producer, 1 = MMA issuer, 2 .. EW + 1 = epilogue.  The warp schedulers favour the
  // highest warp id among the eligible warps, so with the natural order the epilogue warps out-prioritise the two single
  // threads every tile waits for: under a busy epilogue the mainloop of the NEXT tile slowed from 5.4 to 6.4 us.  With
  // roles_high the producer and the issuer are hardware warps EW and EW + 1 and the epilogue warps are 0 .. EW - 1.
  const int hw_warp = threadIdx.x >> 5;
  const int warp = p.roles_high ? (hw_warp >= EW ? hw_warp - EW : hw_warp + 2) : hw_warp;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    if (!GROUPED) {
      tma_prefetch_desc(&tma_a);
      tma_prefetch_desc(&tma_b);
    }
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), CG2 ? 2 : 1);   // pair mode: one arrive(+expect_tx) per CTA, on the leader's barrier
      mbar_init(smem_u32(&empty_bar[s]), MC2 ? 2 : 1);  // MC2: the slot is written by both CTAs' multicasts -> both MMAs must release it
    }
    mbar_init(smem_u32(&tfull_bar[0]), 1);
    mbar_init(smem_u32(&tfull_bar[1]), 1);
    mbar_init(smem_u32(&tempty_bar[0]), CG2 ? 2 * EW : EW);
    mbar_init(smem_u32(&tempty_bar[1]), CG2 ? 2 * EW : EW);
#pragma unroll
    for (int i = 0; i < SCHED_DEPTH; ++i) {
      mbar_init(smem_u32(&sched_full[i]), 1);
      mbar_init(smem_u32(&sched_empty[i]), 1 + EW);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG2) { tmem_alloc_cg2(smem_u32(tmem_slot), C::TMEM_COLS); tmem_relinquish_cg2(); }
    else     { tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CL) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int s_idx = 0;
      uint32_t s_phase = 0;
      for (int item = worker;;) {
        if (dyn) {   // hand the item index (or the end marker) to the MMA thread and the epilogue warps
          mbar_wait(smem_u32(&sched_empty[s_idx]), s_phase ^ 1u);
          sched_item[s_idx] = item < p.num_items ? item : -1;
          mbar_arrive(smem_u32(&sched_full[s_idx]));
          if (++s_idx == SCHED_DEPTH) { s_idx = 0; s_phase ^= 1u; }
        }
        if (item >= p.num_items) break;
        // the next item is requested BEFORE this item's k loop: the atomic's round trip hides under the loads
        const int next_item = dyn ? nworkers + atomicAdd(p.sched, 1) : item + nworkers;
        const ItemCoord ic = decode_item<GROUPED, BN>(item, p, gt);
        const CUtensorMap* pta = GROUPED ? &gt.ta[ic.g] : &tma_a;
        const CUtensorMap* ptb = GROUPED ? &gt.tb[ic.g] : &tma_b;
        const int m0 = (ic.m_blk * (CL ? 2 : 1) + (int)rank) * BM;           // this CTA's 128 rows of the (256-row) super-tile
        const int n0 = ic.n0 + (int)rank * (CG2 ? BN / 2 : 0);               // CG2: this CTA's half of the B tile
        const bool narrow = !CL && ic.bn != BN;                              // N-split unit of the last round (64 columns)
        const uint32_t stage_tx = narrow ? (uint32_t)(C::A_BYTES + ic.bn * BK * 2) : (uint32_t)C::STAGE_BYTES;
        const int kb_begin = ic.kb_begin, kb_end = ic.kb_end;
        // implicit-convolution operand: all divisions happen once per item; the k loop only increments
        PixelCoord cv_px{0, 0, 0};
        int cv_c0 = 0, cv_r = 0, cv_s = 0, cv_wo = 0, cv_ho = 0;
        int cvb_c0[C::B_ROWS / 64];
        uint16_t cvb_ow[C::B_ROWS / 64], cvb_oh[C::B_ROWS / 64];
        if (!GROUPED && p.cv_side == 1) {
          cv_px = conv_base_pixel(p, m0);
          const int k0 = kb_begin * BK;
          const int tap = k0 / p.cv_C;
          cv_c0 = k0 - tap * p.cv_C;
          cv_r = tap / p.cv_kw;
          cv_s = tap - cv_r * p.cv_kw;
        } else if (!GROUPED && p.cv_side == 2) {
          const int pix = kb_begin * BK;
          const int hw = p.cv_Ho * p.cv_Wo;
          cv_px.n = pix / hw;
          const int rem = pix - cv_px.n * hw;
          cv_ho = rem / p.cv_Wo;
          cv_wo = rem - cv_ho * p.cv_Wo;
          cv_px.w = cv_wo * p.cv_stride + p.cv_lower;
          cv_px.h = cv_ho * p.cv_stride + p.cv_lower;
#pragma unroll
          for (int c = 0; c < C::B_ROWS / 64; ++c) {
            const int col = n0 + c * 64;
            cvb_c0[c] = -1; cvb_ow[c] = 0; cvb_oh[c] = 0;
            if (col < p.N) {
              const int tap = col / p.cv_C;
              const int r = tap / p.cv_kw;
              cvb_c0[c] = col - tap * p.cv_C;
              cvb_ow[c] = (uint16_t)((tap - r * p.cv_kw) * p.cv_dil);
              cvb_oh[c] = (uint16_t)(r * p.cv_dil);
            }
          }
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          // pair mode: both CTAs signal the LEADER's full barrier (the leader's MMA thread consumes both halves)
          const uint32_t fb = CG2 ? mapa_cluster(smem_u32(&full_bar[stage]), 0) : smem_u32(&full_bar[stage]);
          // (default .release.cta semantics as in CUTLASS: a cluster-scope release here costs ~1000 cycles per k-block)
          if (CG2 && !leader) mbar_arrive_expect_tx_cluster(fb, C::STAGE_BYTES);
          else mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), stage_tx);
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const int k0 = kb * BK;
          auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1) {
            if (CG2) tma_load_2d_cg2(dst, tm, fb, c0, c1); else tma_load_2d(dst, tm, fb, c0, c1);
          };
          if (!A_MN && !GROUPED && p.cv_side == 1) {
            // implicit convolution: 128 consecutive output pixels x 64 channels of filter tap (r, s), gathered by the TMA unit
            tma_load_im2col_4d(sa, pta, fb, cv_c0, cv_px.w, cv_px.h, cv_px.n, (uint16_t)(cv_s * p.cv_dil), (uint16_t)(cv_r * p.cv_dil));
            cv_c0 += BK;
            if (cv_c0 == p.cv_C) { cv_c0 = 0; if (++cv_s == p.cv_kw) { cv_s = 0; ++cv_r; } }
          } else if (!A_MN) {
            load(sa, pta, k0, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(sa + c * 8192, pta, m0 + c * 64, k0);
          }
          if (MC2) {
            // this CTA fetches half of the B tile and multicasts it to both CTAs (same smem offset, same barrier offset)
            if (!B_MN) {
              tma_load_2d_multicast(sb + rank * (BN / 2) * 128, ptb, fb, k0, n0 + (int)rank * (BN / 2), 3);  // box {64 k, BN/2 rows}
            } else {
#pragma unroll
              for (int c = 0; c < BN / 128; ++c) {
                const int cc = (int)rank * (BN / 128) + c;
                tma_load_2d_multicast(sb + cc * 8192, ptb, fb, n0 + cc * 64, k0, 3);
              }
            }
          } else if (!B_MN) {
            if (narrow) tma_load_2d(sb, &tma_b_tail, fb, k0, n0);   // box {64 k, 64 rows}
            else load(sb, ptb, k0, n0);  // box {64 k, B_ROWS rows}
          } else if (!GROUPED && p.cv_side == 2) {
            // implicit convolution (weight gradient): 64 output pixels (reduction rows) x 64 channels per 64-column chunk
#pragma unroll
            for (int c = 0; c < C::B_ROWS / 64; ++c) {
              if (cvb_c0[c] >= 0) {
                tma_load_im2col_4d(sb + c * 8192, ptb, fb, cvb_c0[c], cv_px.w, cv_px.h, cv_px.n, cvb_ow[c], cvb_oh[c]);
              } else {
                tma_load_im2col_4d(sb + c * 8192, ptb, fb, 0, 0, 0, 0x3fffff, 0, 0);   // past the last column: image index out of bounds -> zero fill, full tx count
              }
            }
            // advance the base pixel by the 64 reduction rows of this k-block (w fastest, then h, then n)
            {
              int wo = cv_wo + BK;
              while (wo >= p.cv_Wo) { wo -= p.cv_Wo; if (++cv_ho == p.cv_Ho) { cv_ho = 0; ++cv_px.n; } }
              cv_wo = wo;
              cv_px.w = wo * p.cv_stride + p.cv_lower;
              cv_px.h = cv_ho * p.cv_stride + p.cv_lower;
            }
          } else {
#pragma unroll
            for (int c = 0; c < C::B_ROWS / 64; ++c)
              if (!narrow || c * 64 < ic.bn) load(sb + c * 8192, ptb, n0 + c * 64, k0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        item = next_item;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc_full = make_idesc_bf16(CG2 ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int s_idx = 0;
      uint32_t s_phase = 0;
      for (int item = worker;; ++it) {
        if (dyn) {
          mbar_wait(smem_u32(&sched_full[s_idx]), s_phase);
          item = sched_item[s_idx];
          mbar_arrive(smem_u32(&sched_empty[s_idx]));
          if (++s_idx == SCHED_DEPTH) { s_idx = 0; s_phase ^= 1u; }
          if (item < 0) break;
        } else {
          item = worker + it * nworkers;
          if (item >= p.num_items) break;
        }
        const ItemCoord ic = decode_item<GROUPED, BN>(item, p, gt);
        const int kb_begin = ic.kb_begin, kb_end = ic.kb_end;
        const uint32_t idesc = ic.bn == BN ? idesc_full : make_idesc_bf16(BM, ic.bn, A_MN ? 1 : 0, B_MN ? 1 : 0);
        const int buf = it & 1;
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        mbar_wait(smem_u32(&tempty_bar[buf]), (use & 1u) ^ 1u);
        tc_fence_after();
        if (p.trace != nullptr && it < 8) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); p.trace[((size_t)blockIdx.x * 8 + it) * 4 + 0] = tt; }
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(sa + k * p.a_kadv, p.a_lbo, p.a_sbo);
            const uint64_t bdesc = make_smem_desc_sw128(sb + k * p.b_kadv, p.b_lbo, p.b_sbo);
            if (CG2) umma_bf16_ss_cg2(d_tmem, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            else umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if (CG2) umma_commit_cg2_mc(smem_u32(&empty_bar[stage]), 3);
          else if (MC2) umma_commit_mc(smem_u32(&empty_bar[stage]), 3);
          else umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if (CG2) umma_commit_cg2_mc(smem_u32(&tfull_bar[buf]), 3); else umma_commit(smem_u32(&tfull_bar[buf]));
        if (p.trace != nullptr && it < 8) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); p.trace[((size_t)blockIdx.x * 8 + it) * 4 + 1] = tt; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = hw_warp & 3;         // TMEM lane quarter this (hardware) warp may access
    const int half = (warp - 2) >> 2;  // which of the EW / 4 warps sharing that quarter
    float* stage = epi_stage + (warp - 2) * STAGE_F32_PER_WARP;
    const DropState dstate = drop_state(p.e.drop);
    int it = 0;
    int s_idx = 0;
    uint32_t s_phase = 0;
    for (int item = worker;; ++it) {
      if (dyn) {
        mbar_wait(smem_u32(&sched_full[s_idx]), s_phase);
        item = sched_item[s_idx];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sched_empty[s_idx]));
        if (++s_idx == SCHED_DEPTH) { s_idx = 0; s_phase ^= 1u; }
        if (item < 0) break;
      } else {
        item = worker + it * nworkers;
        if (item >= p.num_items) break;
      }
      const ItemCoord ic = decode_item<GROUPED, BN>(item, p, gt);
      const int m_blk = ic.m_blk, n_blk = ic.n_blk;
      GemmEpilogue eg = p.e;
      int Mg = p.M, Ng = p.N;
      if (GROUPED) { eg.out = gt.out[ic.g]; eg.ldo = gt.ldo[ic.g]; Mg = gt.M[ic.g]; Ng = gt.N[ic.g]; }
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      if (p.epi_prefetch) {
        // The residual / saved-activation tile this epilogue will read is fetched into L2 while the tile's MMAs run: the
        // per-chunk loads below then cost an L2 hit instead of a DRAM round trip per chunk (the epilogue was latency-bound).
        using T = EpiTraits<EPI>;
        const int act = T::kStatic ? T::act : eg.act;
        const int rk = T::kStatic ? T::resid : eg.resid_kind;
        const bool aux_in = (act == ACT_DGELU_MUL || act == ACT_DRELU_MUL);
        if (aux_in || rk != RESID_NONE) {
          const char* src = reinterpret_cast<const char*>(aux_in ? eg.aux : eg.resid);
          const int esz = (!aux_in && (rk == RESID_F32 || rk == RESID_LN_F32)) ? 4 : 2;
          const size_t ld_bytes = (size_t)(aux_in ? eg.ld_aux : eg.ldr) * esz;
          const int row0 = (m_blk * (CL ? 2 : 1) + (int)rank) * BM;
          const int width = min(ic.bn, Ng - ic.n0) * esz;           // bytes per row of this unit
          const int lpr = (width + 127) >> 7;                        // 128-byte lines per row
          const int et = (warp - 2) * 32 + lane;                     // epilogue thread index 0 .. EW * 32 - 1
          for (int l = et; l < BM * lpr; l += EW * 32) {
            const int r = l / lpr, sgm = l - r * lpr;
            if (row0 + r < Mg)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(row0 + r) * ld_bytes + (size_t)ic.n0 * esz + (size_t)sgm * 128));
          }
        }
      }
      mbar_wait(smem_u32(&tfull_bar[buf]), use & 1u);
      tc_fence_after();
      if (p.trace != nullptr && it < 8 && warp == 2 && lane == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); p.trace[((size_t)blockIdx.x * 8 + it) * 4 + 2] = tt; }
      const int row_base = (m_blk * (CL ? 2 : 1) + (int)rank) * BM + q * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      if (VLB_ENABLE_STREAMK && !GROUPED && CM == 0 && ic.tail >= 0) {
        // ---- stream-K work unit: add this K-chunk's partial tile into the scratch tile; the last chunk to arrive finishes it ----
        float* sk_tile = p.sk_scratch + (size_t)ic.tail * (BM * BN);
        GemmEpilogue es;
        es.out = sk_tile; es.ldo = BN; es.out_kind = OUT_F32_ATOMIC;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if (!(p.sk_debug & 1)) epilogue_chunk<EPI_ATOMIC_F32>(es, v, stage, lane, q * 32, c * 32, BM, BN);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));   // the accumulator buffer is free again
        __threadfence();                                           // this thread's reds are visible device-wide ...
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // ... for every epilogue thread of the CTA
        if (warp == 2 && lane == 0) {
          const int old = atomicAdd(p.sk_counters + ic.tail, 1);
          const bool last = old == p.sk_chunks - 1;
          if (last) p.sk_counters[ic.tail] = 0;                    // self-cleaning: ready for the next launch
          *sk_flag = last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        if (*sk_flag != 0u && !(p.sk_debug & 2)) {
          __threadfence();
#pragma unroll 1
          for (int c = half; c < BN / 32; c += 2) {
            uint32_t v[32];   // unused
            epilogue_chunk<EPI, true>(eg, v, stage, lane, row_base, n_blk * BN + c * 32, Mg, Ng, sk_tile + (size_t)(q * 32) * BN + c * 32, BN);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // sk_flag may be rewritten by the next unit
        continue;
      }
#pragma unroll 1
      for (int c = half; c < ic.bn / 32; c += EW / 4) {
        uint32_t v[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
        const int col0 = ic.n0 + c * 32;
        if (!GROUPED && epi_fast_ok<EPI>(eg) && row_base + 32 <= Mg && col0 + 32 <= Ng)
          epilogue_chunk_fast<EPI>(eg, v, stage, lane, row_base, col0, Ng);
        else if (row_base < Mg && col0 < Ng) epilogue_chunk<EPI>(eg, v, stage, lane, row_base, col0, Mg, Ng, nullptr, 0, dstate);
      }
      tc_fence_before();
      __syncwarp();
      if (p.trace != nullptr && it < 8 && warp == 2 && lane == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); p.trace[((size_t)blockIdx.x * 8 + it) * 4 + 3] = tt; }
      if (lane == 0) {
        if (CG2 && !leader) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty_bar[buf]), 0));  // the leader's MMA thread owns both TMEMs
        else mbar_arrive(smem_u32(&tempty_bar[buf]));
      }
    }
  }

  tc_fence_before();
  if (CL) cluster_sync_all(); else __syncthreads();   // the peer's shared memory / TMEM must outlive everything that targets it
  if (warp == 1) {
    tc_fence_after();
    if (CG2) tmem_dealloc_cg2(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  if (dyn && warp == 0 && lane == 0) {
    // this thread (the producer) has made its last request; the last CTA to get here returns the counters to zero
    __threadfence();
    if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CM, int EW = EPI_WARPS>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_b_tail, const GemmParams p) {
  gemm_body<BN, A_MN, B_MN, EPI, CM, false, EW>(tma_a, tma_b, p, *reinterpret_cast<const GroupTable*>(&tma_a), tma_b_tail);  // table unused
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_grouped_tn_kernel(const __grid_constant__ GroupTable gt, const GemmParams p) {
  gemm_body<BN, true, true, EPI, 0, true>(gt.ta[0], gt.tb[0], p, gt, gt.tb[0]);
}

#include "gemm_pair192.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeIm2col get_encode_im2col_fn() {
  static PFN_encodeIm2col fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeIm2col>(p);
    }
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_cols, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.ld); mix(k.box_cols); mix(k.box_rows);
    return h;
  }
};

template <int BN, bool A_MN, bool B_MN, int EPI, int CM, int EW = EPI_WARPS>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbt, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN, CM, EW>;
  static_assert(EW == EPI_WARPS || CM == 0, "wide epilogues exist for the single-CTA kernel only");
  static bool attr_set = false;
  if (!attr_set) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, A_MN, B_MN, EPI, CM, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::SMEM_BYTES));
    attr_set = true;
  }
  if (CM == 0) {
    const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
    VLB_CHECK_CUDA(launch_pdl(gemm_kernel<BN, A_MN, B_MN, EPI, CM, EW>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, stream, ta, tb, tbt, p));
    return VLB_OK;
  }
  const int pairs = num_sms() / 2;
  const int nclusters = p.num_items < pairs ? p.num_items : pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nclusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, A_MN, B_MN, EPI, CM>, ta, tb, tbt, p));
  return VLB_OK;
}

template <int BN, bool B_MN, int EPI>
int launch_pair192(const CUtensorMap& ta, const CUtensorMap& tb, const PairParams& p, cudaStream_t stream) {
  using C = PairCfg<BN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair192_kernel<BN, B_MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int pairs = num_sms() / 2;
  const int nclusters = p.num_items < pairs ? p.num_items : pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nclusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair192_kernel<BN, B_MN, EPI>, ta, tb, p));
  return VLB_OK;
}

// epilogue that only converts the accumulator to bf16 (the data-gradient GEMMs dh / dx / dctx)
bool epilogue_is_plain_bf16(const GemmEpilogue& e) {
  return e.bias == nullptr && e.act == ACT_NONE && e.resid_kind == RESID_NONE && e.out_kind == OUT_BF16 && e.colsum == nullptr &&
         e.aux == nullptr && e.alpha == 1.0f && e.colscale == nullptr && e.drop.thresh == 0u;
}

// which compile-time epilogue matches this runtime description (EPI_GENERIC if none)
int classify_epilogue(int mode, const GemmEpilogue& e) {
  const bool bias = e.bias != nullptr;
  if (e.colsum == nullptr && e.aux == nullptr && e.alpha == 1.0f && mode == GEMM_NT && e.out_kind == OUT_BF16) {
    if (e.colscale != nullptr && bias && e.act == ACT_RELU && e.resid_kind == RESID_NONE) return EPI_CONV_RELU_BF16;
    if (e.colscale != nullptr && bias && e.act == ACT_RELU_POST && e.resid_kind == RESID_BF16) return EPI_CONV_RESID_RELU_BF16;
    if (e.colscale == nullptr && !bias && e.act == ACT_NONE && e.resid_kind == RESID_NONE) return EPI_PLAIN_BF16;
  }
  if (e.colscale != nullptr) return EPI_GENERIC;  // other per-column-scale combinations: generic epilogue
  if (e.drop.thresh != 0u) {
    if (mode == GEMM_NT && bias && e.act == ACT_NONE && e.out_kind == OUT_F32 && e.colsum == nullptr && e.alpha == 1.0f) {
      if (e.resid_kind == RESID_BF16) return EPI_BIAS_DROP_RESID16_F32;
      if (e.resid_kind == RESID_LN_F32) return EPI_BIAS_DROP_RESIDLN_F32;
    }
    return EPI_GENERIC;
  }
  if (mode == GEMM_NT && bias && e.act == ACT_NONE && e.resid_kind == RESID_LN_F32 && e.out_kind == OUT_F32 && e.colsum == nullptr &&
      e.alpha == 1.0f && e.aux == nullptr)
    return EPI_BIAS_RESIDLN_F32;
  if (mode == GEMM_NT) {
    if (bias && e.act == ACT_NONE && e.resid_kind == RESID_NONE && e.out_kind == OUT_BF16) return EPI_BIAS_BF16;
    if (bias && e.act == ACT_NONE && e.resid_kind == RESID_BF16 && e.out_kind == OUT_F32) return EPI_BIAS_RESID16_F32;
    if (bias && e.act == ACT_GELU && e.resid_kind == RESID_NONE && e.out_kind == OUT_BF16) return EPI_BIAS_GELU_AUX_BF16;
  } else if (mode == GEMM_NN) {
    if (!bias && e.act == ACT_DGELU_MUL && e.resid_kind == RESID_NONE && e.out_kind == OUT_BF16) return EPI_DGELU_BF16;
    if (!bias && e.act == ACT_NONE && e.resid_kind == RESID_BF16 && e.out_kind == OUT_BF16) return EPI_RESID16_BF16;
  } else {
    if (!bias && e.act == ACT_NONE && e.resid_kind == RESID_NONE && e.out_kind == OUT_F32_ATOMIC) return EPI_ATOMIC_F32;
  }
  return EPI_GENERIC;
}

// debug override of the MN-major descriptor geometry (bring-up aid, see tools/gemm_probe.py)
uint32_t g_dbg_mn_lbo = 0, g_dbg_mn_sbo = 0, g_dbg_mn_kadv = 0;
// timing aid: device buffer that receives 8 globaltimer stamps per CTA of the pair kernel (tools/gemm_trace.py)
unsigned long long* g_dbg_trace = nullptr;

}  // namespace

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, box_cols, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return VLB_OK;
    }
  }
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point not available");
    return VLB_ERR_CUDA;
  }
  VLB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  VLB_REQUIRE((ld * 2) % 16 == 0, "TMA leading dimension must be a multiple of 8 elements (ld=%llu)",
              (unsigned long long)ld);
  VLB_REQUIRE(box_cols * 2 <= 128 && box_rows <= 256, "TMA box too large");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                   (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    return VLB_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return VLB_OK;
}

// TMA im2col-mode map of an NHWC bf16 tensor: a load gathers `pixels` consecutive OUTPUT pixels (w fastest, then h, then n; the
// traversal stride is the convolution stride) x 64 channels, displaced by the filter-tap offset given in the instruction;
// taps that fall outside the image read zeros (the padding).  Bounding box as in the PTX ISA's im2col description:
// lower corner = -pad, upper corner = pad - dil * (k - 1)  (relative to the last pixel).
int make_tmap_im2col(CUtensorMap* out, const void* ptr, const ConvGeom& g, uint32_t pixels) {
  PFN_encodeIm2col fn = get_encode_im2col_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeIm2col entry point not available");
    return VLB_ERR_CUDA;
  }
  VLB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  VLB_REQUIRE(g.C % 64 == 0, "implicit convolution needs a multiple of 64 input channels (C=%d)", g.C);
  cuuint64_t gdim[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
  cuuint64_t gstride[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
  int lower[2] = {-g.pad, -g.pad};
  int upper[2] = {g.pad - g.dil * (g.kw - 1), g.pad - g.dil * (g.kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, lower, upper, 64, pixels, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeIm2col failed with CUresult %d (N=%d H=%d W=%d C=%d k=%dx%d stride=%d pad=%d dil=%d)", (int)r,
                   g.N, g.H, g.W, g.C, g.kh, g.kw, g.stride, g.pad, g.dil);
    return VLB_ERR_CUDA;
  }
  // drivers up to 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (same correction as CUTLASS's make_im2col_tma_copy_desc)
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (size_t)g.N * g.H * g.W * g.C * 2 < 131072)
    reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  *out = m;
  return VLB_OK;
}

bool gemm_streamk_compiled() { return VLB_ENABLE_STREAMK != 0; }

void gemm_debug_trace(unsigned long long* buf) { g_dbg_trace = buf; }

void gemm_debug_override(uint32_t mn_lbo, uint32_t mn_sbo, uint32_t mn_kadv) {
  g_dbg_mn_lbo = mn_lbo;
  g_dbg_mn_sbo = mn_sbo;
  g_dbg_mn_kadv = mn_kadv;
}

namespace {
// {next item, finished CTAs} counter pairs of the dynamic tile scheduler: a pool handed out round robin, one pair per launch
// (consecutive launches never share a pair; a pair is back to zero when its kernel ends).  Allocated on first use outside of
// stream capture; a launch that finds no pool (first call ever made inside a capture) falls back to static scheduling.
constexpr int SCHED_POOL = 2048;
int* sched_counters(cudaStream_t stream) {
  static int* pool = nullptr;
  static std::atomic<unsigned> next{0};
  static std::mutex mu;
  static const int on = [] { const char* v = getenv("VLB_DYN_SCHED"); return v ? atoi(v) : 1; }();
  if (!on) return nullptr;
  if (pool == nullptr) {
    std::lock_guard<std::mutex> g(mu);
    if (pool == nullptr) {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
      int* ptr = nullptr;
      if (cudaMalloc(&ptr, SCHED_POOL * 2 * sizeof(int)) != cudaSuccess) return nullptr;
      cudaMemset(ptr, 0, SCHED_POOL * 2 * sizeof(int));
      cudaDeviceSynchronize();
      pool = ptr;
    }
  }
  return pool + 2 * (next.fetch_add(1, std::memory_order_relaxed) % SCHED_POOL);
}

// N-split of the last partial round (see GemmParams::tail_split): relative cost of one 128 x 64 unit against a 128 x 256 tile
// (a quarter of the MMA work, but the same A tile per k-block: the unit is operand-feed bound, not MMA bound).
constexpr double kTailUnitCost = 0.45;
int tail_split_for(long tiles, int sms, int bn, int split_k, int mode, bool conv) {
  static const int on = [] { const char* v = getenv("VLB_TAIL_SPLIT"); return v ? atoi(v) : 1; }();
  if (!on || conv || split_k > 1 || mode == GEMM_TN || bn < 128) return 0;
  const long rem = tiles % sms;
  if (rem == 0 || tiles < sms) return 0;          // a single (partial) round gains nothing
  const int split = bn / 64;
  if (rem * split > sms) return 0;                // the units must fit one extra (short) round
  return split;
}

constexpr int SK_MAX_TAIL_TILES = 80;          // tail tiles are at most half a round (148 / 2)
constexpr int SK_MAX_CHUNKS = 16;

// how many K-chunks each tile of the last partial round is split into (0 / 1: leave the round as it is)
int streamk_chunks(long tiles, int sms, int num_k_blocks) {
  static const int on = [] { const char* v = getenv("VLB_STREAMK"); return v ? atoi(v) : 1; }();
  if (!VLB_ENABLE_STREAMK || !on) return 0;
  const int rem = (int)(tiles % sms);
  static const int min_kb = [] { const char* v = getenv("VLB_SK_MIN_KB"); return v ? atoi(v) : 24; }();
  // measured: partial reds + arrival + fix-up cost ~5 us per launch -- only a long reduction (K >= 1536) repays it
  if (rem == 0 || rem * 2 > sms || rem > SK_MAX_TAIL_TILES || num_k_blocks < min_kb) return 0;
  int chunks = sms / rem;
  if (chunks > num_k_blocks / 2) chunks = num_k_blocks / 2;   // at least two k-blocks per unit
  if (chunks > SK_MAX_CHUNKS) chunks = SK_MAX_CHUNKS;
  static const int cap = [] { const char* v = getenv("VLB_SK_MAXCHUNKS"); return v ? atoi(v) : SK_MAX_CHUNKS; }();
  if (chunks > cap) chunks = cap;
  return chunks;
}

// fp32 scratch tiles + arrival counters of the stream-K tail: allocated once (never while a stream is capturing), zero on
// entry of every launch and zeroed again by the unit that consumes them.  One GEMM stream per process.
bool streamk_workspace(float** scratch, int** counters, cudaStream_t stream) {
  static float* s_scratch = nullptr;
  static int* s_counters = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (s_scratch == nullptr) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return false;
    const size_t bytes = (size_t)SK_MAX_TAIL_TILES * BM * 256 * sizeof(float);
    if (cudaMalloc(&s_scratch, bytes) != cudaSuccess) { s_scratch = nullptr; return false; }
    if (cudaMalloc(&s_counters, SK_MAX_TAIL_TILES * sizeof(int)) != cudaSuccess) { cudaFree(s_scratch); s_scratch = nullptr; return false; }
    cudaMemset(s_scratch, 0, bytes);
    cudaMemset(s_counters, 0, SK_MAX_TAIL_TILES * sizeof(int));
    cudaDeviceSynchronize();
  }
  *scratch = s_scratch;
  *counters = s_counters;
  return true;
}
}  // namespace

int gemm_bf16(int mode, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
              const GemmEpilogue& epi_in, int split_k, int force_bn, cudaStream_t stream, const ConvGeom* conv, int conv_side) {
  const GemmEpilogue& epi = epi_in;
  VLB_REQUIRE(mode >= GEMM_NT && mode <= GEMM_TN, "gemm: bad mode %d", mode);
  VLB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VLB_REQUIRE(A && B && epi.out, "gemm: null pointer");
  VLB_REQUIRE(N % 8 == 0, "gemm: N (%d) must be a multiple of 8", N);
  VLB_REQUIRE(epi.ldo % 4 == 0 && (epi.out_kind != OUT_BF16 || epi.ldo % 8 == 0), "gemm: bad ldo %d", epi.ldo);
  VLB_REQUIRE(epi.resid_kind == RESID_NONE || (epi.resid != nullptr && epi.ldr % 8 == 0), "gemm: bad residual");
  VLB_REQUIRE(epi.resid_kind != RESID_LN_F32 || epi.ln_mean == nullptr || (epi.ln_rstd && epi.ln_gamma && epi.ln_beta),
              "gemm: LayerNorm-recomputed residual needs mean, rstd, gamma and beta");
  VLB_REQUIRE(epi.drop.thresh == 0u || epi.drop.bits != nullptr, "gemm: the dropout epilogue needs precomputed keep bits (vlb_dropout_bits)");
  VLB_REQUIRE((epi.act != ACT_DGELU_MUL && epi.act != ACT_DRELU_MUL) || (epi.aux != nullptr && epi.ld_aux % 8 == 0),
              "gemm: activation-gradient epilogue needs aux");
  const bool a_mn = (mode == GEMM_TN);
  const bool b_mn = (mode != GEMM_NT);

  // Tile-N choice: fewest "rounds" of the persistent grid weighted by tile cost.
  // force_bn: 0 = heuristic; 64/128/192/256 = single-CTA tile width; 1128/1256 = CTA-pair (cta_group::2) 256 x {128,256} tiles;
  // 2128/2256 = two-CTA clusters with the B tile TMA-multicast (independent 128 x {128,256} MMAs).
  static const int env_bn = [] { const char* v = getenv("VLB_FORCE_BN"); return v ? atoi(v) : 0; }();  // tuning aids
  static const int env_cg2 = [] { const char* v = getenv("VLB_CG2"); return v ? atoi(v) : 0; }();  // 0 off (default: measured ~5% slower in situ), 1 force, -1 auto
  static const int env_mc2 = [] { const char* v = getenv("VLB_MC2"); return v ? atoi(v) : 0; }();  // 1: B-multicast clusters wherever the tile is 128/256 wide
  if (force_bn == 0 && env_bn != 0 && (N >= (env_bn % 1000) || env_bn == 64)) force_bn = env_bn;
  // CTA-pair kernel with 192-row CTA tiles (gemm_pair192.cuh): force_bn 3192 / 3256, or chosen automatically for problems that
  // fill between half a wave and one wave of 384 x BN pair tiles (the encoder's N = 768 GEMMs at M = 6464: 68 of 74 pairs).
  // VLB_PAIR192: 0 = never, 1 = single-wave problems (default), 2 = wherever the shape allows it.
  {
    static const int env_pair = [] { const char* v = getenv("VLB_PAIR192"); return v ? atoi(v) : 1; }();
    static const int env_pair_nn_bn = [] { const char* v = getenv("VLB_PAIR192_NN_BN"); return v ? atoi(v) : 192; }();
    int pbn = 0;
    if (force_bn == 3192 || force_bn == 3256) {
      pbn = force_bn - 3000;
      VLB_REQUIRE(mode != GEMM_TN && conv == nullptr && split_k <= 1 && N % pbn == 0, "gemm: the pair-192 kernel needs NT / NN, no split-K, N %% %d == 0", pbn);
    } else if (force_bn == 0 && env_pair != 0 && mode != GEMM_TN && conv == nullptr && split_k <= 1) {
      int cand = mode == GEMM_NN ? env_pair_nn_bn : 192;
      if (N % cand != 0) cand = (N % 256 == 0) ? 256 : ((N % 192 == 0) ? 192 : 0);
      if (cand != 0) {
        const long items = (long)((M + 383) / 384) * (N / cand);
        const int pairs = num_sms() / 2;
        // one wave of pair tiles (the epilogue is exposed but there is only one); or -- measured at configs 3 / 4 -- the data
        // gradients with a plain bf16 store (dh / dctx / dx), whose short epilogue the double-buffered accumulator hides
        const bool one_wave = items <= pairs && 2 * items >= pairs;
        const bool light_dgrad = mode == GEMM_NN && epilogue_is_plain_bf16(epi) && M >= 384 && items > pairs;
        if (env_pair == 2 ? M >= 384 : (one_wave || light_dgrad)) pbn = cand;
      }
    }
    if (pbn != 0) {
      PairParams q;
      q.M = M; q.N = N; q.K = K;
      q.num_m_pairs = (M + 383) / 384;
      q.num_n_blocks = N / pbn;
      q.num_k_blocks = (K + BK - 1) / BK;
      q.num_items = q.num_m_pairs * q.num_n_blocks;
      static const int env_pair_prefetch = [] { const char* v = getenv("VLB_EPI_PREFETCH"); return v ? atoi(v) : 1; }();
      q.epi_prefetch = env_pair_prefetch;
      static const int env_pair_stage = [] { const char* v = getenv("VLB_EPI_STAGE"); return v ? atoi(v) : 1; }();
      q.epi_stage = env_pair_stage;
      static const int env_pair_roles_high = [] { const char* v = getenv("VLB_ROLES_HIGH"); return v ? atoi(v) : 1; }();
      q.roles_high = env_pair_roles_high;
      q.trace = g_dbg_trace;
      q.e = epi;
      CUtensorMap ta, tb;
      int rc = make_tmap_bf16_2d(&ta, A, M, K, lda, 64, 192);
      if (rc != VLB_OK) return rc;
      rc = b_mn ? make_tmap_bf16_2d(&tb, B, K, N, ldb, 64, 64) : make_tmap_bf16_2d(&tb, B, N, K, ldb, 64, pbn / 2);
      if (rc != VLB_OK) return rc;
      ProfScope prof(mode == GEMM_NT ? PROF_GEMM_NT : PROF_GEMM_NN, 2.0 * M * N * K, stream);
      int epi_id = classify_epilogue(mode, epi);
      if (epi_id == EPI_GENERIC && epilogue_is_plain_bf16(epi)) epi_id = EPI_PLAIN_BF16;
#define VLB_PAIR_DISPATCH(BN_)                                                                                                   \
      if (!b_mn) {                                                                                                               \
        if (epi_id == EPI_BIAS_BF16) return launch_pair192<BN_, false, EPI_BIAS_BF16>(ta, tb, q, stream);                        \
        if (epi_id == EPI_BIAS_RESIDLN_F32) return launch_pair192<BN_, false, EPI_BIAS_RESIDLN_F32>(ta, tb, q, stream);          \
        if (epi_id == EPI_BIAS_DROP_RESIDLN_F32) return launch_pair192<BN_, false, EPI_BIAS_DROP_RESIDLN_F32>(ta, tb, q, stream); \
        if (epi_id == EPI_BIAS_GELU_AUX_BF16) return launch_pair192<BN_, false, EPI_BIAS_GELU_AUX_BF16>(ta, tb, q, stream);      \
        if (epi_id == EPI_PLAIN_BF16) return launch_pair192<BN_, false, EPI_PLAIN_BF16>(ta, tb, q, stream);                      \
        return launch_pair192<BN_, false, EPI_GENERIC>(ta, tb, q, stream);                                                       \
      }                                                                                                                          \
      if (epi_id == EPI_DGELU_BF16) return launch_pair192<BN_, true, EPI_DGELU_BF16>(ta, tb, q, stream);                         \
      if (epi_id == EPI_PLAIN_BF16) return launch_pair192<BN_, true, EPI_PLAIN_BF16>(ta, tb, q, stream);                         \
      return launch_pair192<BN_, true, EPI_GENERIC>(ta, tb, q, stream);
      if (pbn == 256) { VLB_PAIR_DISPATCH(256) }
      VLB_PAIR_DISPATCH(192)
#undef VLB_PAIR_DISPATCH
    }
  }
  int bn = 128;
  bool cg2 = false;
  int cm = 0;
  const int sms = num_sms();
  if (force_bn == 4192) {
    bn = 192;
  } else if (force_bn >= 2000) {
    cm = 2;
    bn = force_bn - 2000;
    VLB_REQUIRE(bn == 128 || bn == 256, "gemm: cluster modes support BN 128 / 256");
  } else if (force_bn >= 1000) {
    cg2 = true;
    bn = force_bn - 1000;
    VLB_REQUIRE(bn == 128 || bn == 256, "gemm: pair mode supports BN 128 / 256");
  } else if (force_bn == 128 || force_bn == 256 || force_bn == 64 || force_bn == 192) {
    bn = force_bn;
  } else {
    // cost = rounds of the persistent grid x measured relative time of one tile; with the stream-K tail the last partial
    // round costs 1/chunks of a round plus the fix-up
    const int sk = split_k > 1 ? split_k : 1;
    auto cost1 = [&](int b) {
      const long items = (long)((M + BM - 1) / BM) * ((N + b - 1) / b) * sk;
      const double w = b == 64 ? 0.60 : (b == 128 ? 0.68 : (b == 192 ? 0.84 : 1.0));
      const int chunks = (sk == 1 && mode != GEMM_TN) ? streamk_chunks(items, sms, (K + BK - 1) / BK) : 0;
      if (chunks > 1) return ((double)(items / sms) + 1.0 / chunks + 0.2) * w;
      const long rounds = (items + sms - 1) / sms;
      if (tail_split_for(items, sms, b, sk, mode, conv != nullptr && conv_side != 0) > 1) return (double)(items / sms) * w + kTailUnitCost;
      return rounds * w;
    };
    auto cost2 = [&](int b) {  // pair tiles: half the B traffic per CTA, same MMA time per round as the 128 x b tile
      const long items = (long)((M + 2 * BM - 1) / (2 * BM)) * ((N + b - 1) / b) * sk;
      const long rounds = (items + sms / 2 - 1) / (sms / 2);
      const double w = b == 128 ? 0.55 : 0.80;
      return rounds * w;
    };
    bn = 128;
    double best = cost1(128);
    if (N >= 256 && cost1(256) < best) { best = cost1(256); bn = 256; }
    if (N >= 192 && cost1(192) < best) { best = cost1(192); bn = 192; }
    if (cost1(64) < best) { best = cost1(64); bn = 64; }
    if (env_cg2 != 0) {
      const bool force = env_cg2 == 1;
      if (N >= 256 && (cost2(256) < best || force)) { best = cost2(256); bn = 256; cg2 = true; }
      if (N >= 128 && (cost2(128) < best || (force && !cg2))) { best = cost2(128); bn = 128; cg2 = true; }
    }
  }

  if (cg2) cm = 1;
  if (cm == 0 && env_mc2 == 1 && force_bn == 0 && (bn == 128 || bn == 256)) cm = 2;
  // GELU launch (BertIntermediate): 128 x 192 tiles with 16 epilogue warps (see gemm_body); force_bn 4192 or VLB_GELU_EW16=1
  static const int env_ew16 = [] { const char* v = getenv("VLB_GELU_EW16"); return v ? atoi(v) : 0; }();
  bool ew16 = false;
  if (mode == GEMM_NT && epi.act == ACT_GELU && conv == nullptr && split_k <= 1 && N >= 192 &&
      (force_bn == 4192 || (force_bn == 0 && env_ew16 && cm == 0 && classify_epilogue(mode, epi) == EPI_BIAS_GELU_AUX_BF16))) {
    VLB_REQUIRE(classify_epilogue(mode, epi) == EPI_BIAS_GELU_AUX_BF16, "gemm: the 16-warp epilogue exists for bias + GELU (+ saved GELU') -> bf16 only");
    ew16 = true;
    bn = 192;
    cm = 0;
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = cm != 0 ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
  p.num_n_blocks = (N + bn - 1) / bn;
  p.num_k_blocks = (K + BK - 1) / BK;
  int sk = split_k < 1 ? 1 : split_k;
  if (sk > p.num_k_blocks) sk = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + sk - 1) / sk;
  sk = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.split_k = sk;
  VLB_REQUIRE(sk == 1 || epi.out_kind == OUT_F32_ATOMIC, "gemm: split-K needs the atomic fp32 epilogue");
  p.num_items = p.num_m_blocks * p.num_n_blocks * sk;
  p.e = epi;
  p.sk_full_items = p.num_items; p.sk_chunks = 0; p.sk_kb_per_chunk = 0; p.sk_scratch = nullptr; p.sk_counters = nullptr;
  static const int sk_debug = [] { const char* v = getenv("VLB_SK_DEBUG"); return v ? atoi(v) : 0; }();
  p.sk_debug = sk_debug;
  if (cm == 0 && sk == 1 && mode != GEMM_TN) {
    const int chunks = streamk_chunks(p.num_items, sms, p.num_k_blocks);
    if (chunks > 1 && streamk_workspace(&p.sk_scratch, &p.sk_counters, stream)) {
      const int tiles = p.num_items;
      const int rem = tiles % sms;
      p.sk_full_items = tiles - rem;
      p.sk_kb_per_chunk = (p.num_k_blocks + chunks - 1) / chunks;
      p.sk_chunks = (p.num_k_blocks + p.sk_kb_per_chunk - 1) / p.sk_kb_per_chunk;   // no empty chunks
      p.num_items = p.sk_full_items + rem * p.sk_chunks;
    }
  }
  // K-major operand: rows x 128B, swizzle atoms of 8 rows -> SBO 1024B, 32B per UMMA_K step.
  // MN-major operand: boxes of 64(mn) x 64(k): k-groups of 8 rows at 1024B (SBO), next 64-wide mn
  // chunk at 8192B (LBO), 16 k-rows = 2048B per UMMA_K step.
  const uint32_t mn_lbo = g_dbg_mn_lbo ? g_dbg_mn_lbo : 8192u;
  const uint32_t mn_sbo = g_dbg_mn_sbo ? g_dbg_mn_sbo : 1024u;
  const uint32_t mn_kadv = g_dbg_mn_kadv ? g_dbg_mn_kadv : 2048u;
  p.a_lbo = a_mn ? mn_lbo : 16u;  p.a_sbo = a_mn ? mn_sbo : 1024u;  p.a_kadv = a_mn ? mn_kadv : 32u;
  p.b_lbo = b_mn ? mn_lbo : 16u;  p.b_sbo = b_mn ? mn_sbo : 1024u;  p.b_kadv = b_mn ? mn_kadv : 32u;

  // N-split units for the last partial round
  p.tail_first = p.num_items; p.tail_split = 0;
  if (cm == 0 && p.sk_chunks == 0) {
    const int ts = tail_split_for(p.num_items, sms, bn, sk, mode, conv != nullptr && conv_side != 0);
    if (ts > 1) {
      const int rem = p.num_items % sms;
      p.tail_first = p.num_items - rem;
      p.tail_split = ts;
      p.num_items = p.tail_first + rem * ts;
    }
  }
  static const int env_prefetch = [] { const char* v = getenv("VLB_EPI_PREFETCH"); return v ? atoi(v) : 1; }();
  p.epi_prefetch = env_prefetch;
  p.sched = (cm == 0 && p.sk_chunks == 0 && p.num_items > sms) ? sched_counters(stream) : nullptr;
  p.trace = g_dbg_trace;
  static const int env_roles_high = [] { const char* v = getenv("VLB_ROLES_HIGH"); return v ? atoi(v) : 1; }();
  p.roles_high = env_roles_high;

  p.cv_side = 0;
  if (conv != nullptr && conv_side != 0) {
    VLB_REQUIRE((conv_side == 1 && mode == GEMM_NT) || (conv_side == 2 && mode == GEMM_TN), "gemm: implicit-conv operand / mode mismatch");
    VLB_REQUIRE(cm == 0, "gemm: implicit convolution runs in single-CTA mode");
    const long pixels = (long)conv->N * conv->Ho * conv->Wo;
    const int taps_k = conv->kh * conv->kw * conv->C;
    VLB_REQUIRE(conv_side == 1 ? (M == pixels && K == taps_k) : (K == pixels && N == taps_k), "gemm: implicit-conv shape mismatch");
    p.cv_side = conv_side;
    p.cv_C = conv->C; p.cv_Ho = conv->Ho; p.cv_Wo = conv->Wo; p.cv_stride = conv->stride; p.cv_lower = -conv->pad;
    p.cv_dil = conv->dil; p.cv_kw = conv->kw;
  }

  CUtensorMap ta, tb;
  int rc;
  if (p.cv_side == 1) rc = make_tmap_im2col(&ta, A, *conv, BM);        // A = im2col(x): 128 output pixels x 64 channels per load
  else if (!a_mn) rc = make_tmap_bf16_2d(&ta, A, M, K, lda, 64, BM);       // A [M, K]
  else       rc = make_tmap_bf16_2d(&ta, A, K, M, lda, 64, 64);       // A stored [K, M]
  if (rc != VLB_OK) return rc;
  if (p.cv_side == 2) rc = make_tmap_im2col(&tb, B, *conv, 64);        // B = im2col(x): 64 output pixels x 64 channels per load
  else if (!b_mn) rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, 64, cm != 0 ? bn / 2 : bn);  // B [N, K]; cluster modes fetch it in halves
  else       rc = make_tmap_bf16_2d(&tb, B, K, N, ldb, 64, 64);       // B stored [K, N]
  if (rc != VLB_OK) return rc;
  CUtensorMap tbt = tb;                                                 // 64-row box of a K-major B for the N-split units
  if (p.tail_split > 1 && !b_mn) {
    rc = make_tmap_bf16_2d(&tbt, B, N, K, ldb, 64, 64);
    if (rc != VLB_OK) return rc;
  }

  ProfScope prof(mode == GEMM_NT ? PROF_GEMM_NT : (mode == GEMM_NN ? PROF_GEMM_NN : PROF_GEMM_TN), 2.0 * M * N * K, stream);
#define VLB_GEMM_DISPATCH(BN_, CG_)                                                                          \
  if (!a_mn && !b_mn) {                                                                                      \
    if (epi_id == EPI_BIAS_BF16) return launch<BN_, false, false, EPI_BIAS_BF16, CG_>(ta, tb, tbt, p, stream);     \
    if (epi_id == EPI_BIAS_RESID16_F32) return launch<BN_, false, false, EPI_BIAS_RESID16_F32, CG_>(ta, tb, tbt, p, stream); \
    if (CG_ == 0 && epi_id == EPI_BIAS_DROP_RESID16_F32) return launch<BN_, false, false, EPI_BIAS_DROP_RESID16_F32, 0>(ta, tb, tbt, p, stream); \
    if (CG_ == 0 && epi_id == EPI_BIAS_RESIDLN_F32) return launch<BN_, false, false, EPI_BIAS_RESIDLN_F32, 0>(ta, tb, tbt, p, stream); \
    if (CG_ == 0 && epi_id == EPI_BIAS_DROP_RESIDLN_F32) return launch<BN_, false, false, EPI_BIAS_DROP_RESIDLN_F32, 0>(ta, tb, tbt, p, stream); \
    if (epi_id == EPI_BIAS_GELU_AUX_BF16) return launch<BN_, false, false, EPI_BIAS_GELU_AUX_BF16, CG_>(ta, tb, tbt, p, stream); \
    if (CG_ == 0 && epi_id == EPI_CONV_RELU_BF16) return launch<BN_, false, false, EPI_CONV_RELU_BF16, 0>(ta, tb, tbt, p, stream); \
    if (CG_ == 0 && epi_id == EPI_CONV_RESID_RELU_BF16) return launch<BN_, false, false, EPI_CONV_RESID_RELU_BF16, 0>(ta, tb, tbt, p, stream); \
    if (CG_ == 0 && epi_id == EPI_PLAIN_BF16) return launch<BN_, false, false, EPI_PLAIN_BF16, 0>(ta, tb, tbt, p, stream); \
    return launch<BN_, false, false, EPI_GENERIC, CG_>(ta, tb, tbt, p, stream);                                    \
  }                                                                                                          \
  if (!a_mn && b_mn) {                                                                                       \
    if (epi_id == EPI_DGELU_BF16) return launch<BN_, false, true, EPI_DGELU_BF16, CG_>(ta, tb, tbt, p, stream);    \
    if (epi_id == EPI_RESID16_BF16) return launch<BN_, false, true, EPI_RESID16_BF16, CG_>(ta, tb, tbt, p, stream); \
    return launch<BN_, false, true, EPI_GENERIC, CG_>(ta, tb, tbt, p, stream);                                     \
  }                                                                                                          \
  if (epi_id == EPI_ATOMIC_F32) return launch<BN_, true, true, EPI_ATOMIC_F32, CG_>(ta, tb, tbt, p, stream);       \
  return launch<BN_, true, true, EPI_GENERIC, CG_>(ta, tb, tbt, p, stream);
  const int epi_id = classify_epilogue(mode, epi_in);
  if (ew16) return launch<192, false, false, EPI_BIAS_GELU_AUX_BF16, 0, 16>(ta, tb, tbt, p, stream);
  if (cm == 1) {
    if (bn == 256) { VLB_GEMM_DISPATCH(256, 1) }
    VLB_GEMM_DISPATCH(128, 1)
  }
  if (cm == 2) {
    if (bn == 256) { VLB_GEMM_DISPATCH(256, 2) }
    VLB_GEMM_DISPATCH(128, 2)
  }
  if (bn == 256) { VLB_GEMM_DISPATCH(256, 0) }
  if (bn == 192) { VLB_GEMM_DISPATCH(192, 0) }
  if (bn == 64) { VLB_GEMM_DISPATCH(64, 0) }
  VLB_GEMM_DISPATCH(128, 0)
#undef VLB_GEMM_DISPATCH
}

namespace {
template <int BN, int EPI>
int launch_grouped(const GroupTable& gt, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN, 0>;
  static bool attr_set = false;
  if (!attr_set) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(gemm_grouped_tn_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  VLB_CHECK_CUDA(launch_pdl(gemm_grouped_tn_kernel<BN, EPI>, dim3(grid), dim3(GEMM_THREADS), C::SMEM_BYTES, stream, gt, p));
  return VLB_OK;
}
}  // namespace

int gemm_grouped_tn(int count, const GroupedProblem* probs, int K, int split_k, bool accumulate, int bn, cudaStream_t stream) {
  VLB_REQUIRE(count >= 1 && count <= MAX_GROUP && probs != nullptr, "gemm_grouped_tn: 1..%d problems", MAX_GROUP);
  VLB_REQUIRE(bn == 128 || bn == 256, "gemm_grouped_tn: bn must be 128 or 256");
  VLB_REQUIRE(K > 0, "gemm_grouped_tn: K");
  GemmParams p;
  p.M = p.N = 0; p.K = K;
  p.num_m_blocks = p.num_n_blocks = 0;
  p.num_k_blocks = (K + BK - 1) / BK;
  int sk = split_k < 1 ? 1 : split_k;
  if (sk > p.num_k_blocks) sk = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + sk - 1) / sk;
  sk = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.split_k = sk;
  VLB_REQUIRE(sk == 1 || accumulate, "gemm_grouped_tn: split-K needs accumulate");
  p.e = GemmEpilogue();
  p.e.out_kind = accumulate ? OUT_F32_ATOMIC : OUT_F32;
  p.tail_first = 0; p.tail_split = 0; p.epi_prefetch = 0; p.sched = nullptr; p.trace = nullptr; p.roles_high = 1;
  p.sk_full_items = 0; p.sk_chunks = 0; p.sk_kb_per_chunk = 0; p.sk_scratch = nullptr; p.sk_counters = nullptr; p.sk_debug = 0;
  p.cv_side = 0;
  p.a_lbo = p.b_lbo = g_dbg_mn_lbo ? g_dbg_mn_lbo : 8192u;
  p.a_sbo = p.b_sbo = g_dbg_mn_sbo ? g_dbg_mn_sbo : 1024u;
  p.a_kadv = p.b_kadv = g_dbg_mn_kadv ? g_dbg_mn_kadv : 2048u;
  GroupTable gt;
  gt.count = count;
  int items = 0;
  double flops = 0;
  for (int i = 0; i < count; ++i) {
    const GroupedProblem& q = probs[i];
    VLB_REQUIRE(q.M > 0 && q.N > 0 && q.N % 8 == 0 && q.M % 8 == 0 && q.A && q.B && q.out && q.ldo % 4 == 0, "gemm_grouped_tn: bad problem %d", i);
    int rc = make_tmap_bf16_2d(&gt.ta[i], q.A, K, q.M, q.lda, 64, 64);
    if (rc != VLB_OK) return rc;
    rc = make_tmap_bf16_2d(&gt.tb[i], q.B, K, q.N, q.ldb, 64, 64);
    if (rc != VLB_OK) return rc;
    gt.m_blocks[i] = (q.M + BM - 1) / BM;
    gt.n_blocks[i] = (q.N + bn - 1) / bn;
    gt.M[i] = q.M; gt.N[i] = q.N;
    gt.out[i] = q.out; gt.ldo[i] = q.ldo;
    gt.item_begin[i] = items;
    items += gt.m_blocks[i] * gt.n_blocks[i] * sk;
    flops += 2.0 * q.M * q.N * K;
  }
  for (int i = count; i <= MAX_GROUP; ++i) gt.item_begin[i] = items;
  for (int i = count; i < MAX_GROUP; ++i) { gt.m_blocks[i] = gt.n_blocks[i] = 1; gt.M[i] = gt.N[i] = 0; gt.out[i] = nullptr; gt.ldo[i] = 0; gt.ta[i] = gt.ta[0]; gt.tb[i] = gt.tb[0]; }
  p.num_items = items;
  p.sched = items > num_sms() ? sched_counters(stream) : nullptr;
  ProfScope prof(PROF_GEMM_TN, flops, stream);
  if (bn == 256) return accumulate ? launch_grouped<256, EPI_ATOMIC_F32>(gt, p, stream) : launch_grouped<256, EPI_GENERIC>(gt, p, stream);
  return accumulate ? launch_grouped<128, EPI_ATOMIC_F32>(gt, p, stream) : launch_grouped<128, EPI_GENERIC>(gt, p, stream);
}

}  // namespace vlb
