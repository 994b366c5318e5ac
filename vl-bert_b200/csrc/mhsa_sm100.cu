// Fused multi-head self-attention for the VL-BERT encoder on sm_100a (tcgen05 + TMEM + TMA).
//
// Reference semantics: BertSelfAttention.forward, external/pytorch_pretrained_bert/modeling.py:290-315
//   scores = Q K^T / sqrt(d_h) + additive_mask[b, key]   (mask = 0 / -10000, common/visual_linguistic_bert.py:119-127)
//   probs  = softmax(scores) ; ctx = probs V ; ctx written back in [B*S, H] layout (the permute+contiguous
//   of :313-315 is folded into the store).  Dropout (modeling.py:310) is p = 0 here.
//
// The whole [text ; region ; END] sequence of a sample fits one tile (S <= 128), so there is one CTA per
// (batch, head): Q/K/V head slices arrive by TMA straight from the fused QKV activation [B*S, 3H],
// S = QK^T and O = PV run on the tensor cores with accumulators in TMEM, the S x S score / probability
// matrices never touch HBM.  Thread t of the CTA owns query row t (TMEM lane t).
//
// Backward recomputes P from Q, K and the saved log-sum-exp, then dV = P^T dO, dP = dO V^T,
// dS = P o (dP - rowsum(dO o O)) / sqrt(d_h), dQ = dS K, dK = dS^T Q -- five tensor-core products per
// (batch, head) with P and dS staged once in shared memory and consumed in both orientations
// (K-major for dQ, MN-major for dK / dV).
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "philox.cuh"

namespace vlb {

namespace {

constexpr int D_HEAD = 64;
constexpr int TQ = 128;  // query rows per CTA (= TMEM lanes)
constexpr int NK = 128;  // keys per CTA

struct MhsaParams {
  int B, S, H, heads;
  float scale;            // 1 / sqrt(d_h)
  const float* add_mask;  // [B, S] additive (0 / -10000), may be null
  __nv_bfloat16* ctx;     // [B*S, H]
  float* lse;             // [B, heads, S]
  // backward
  const __nv_bfloat16* dctx;  // [B*S, H]
  __nv_bfloat16* dqkv;        // [B*S, 3H]
  float* dqkv_f32;            // [B*S, 3H] fp32 accumulation buffer, only for S > 128 (several (q-tile, k-tile) blocks)
  int ktiles;                 // number of 128-key tiles (blockIdx.z = q_tile * ktiles + k_tile)
  DropCfg drop;               // dropout on the probabilities (modeling.py:310); mask rows = (b, head, query), columns = keys
  float* dbias;               // optional [3H] fp32: += column sums of dqkv (the bias gradients of the query / key / value Linear)
};

// Dropout on the probabilities reads precomputed keep flags (DropCfg::bits, written by dropout_bits_kernel for the [B*heads*S, S]
// mask): word c of mask row `mrow` holds keys 32c .. 32c+31.  Evaluating Philox inside these kernels made them
// instruction-bound (ten rounds per four keys, twice per step: forward and the recomputation in backward).
__device__ __forceinline__ uint32_t keep_word(const DropCfg& d, uint64_t mrow, int wpr, int c) {
  return c < wpr ? __ldg(d.bits + mrow * (uint64_t)wpr + (uint64_t)c) : 0u;
}

// Shared memory maps (bytes from the 1024-aligned dynamic shared memory base).  Regions are re-used once their first
// consumer is done so that more CTAs fit an SM (the kernels are latency-bound chains, occupancy is what hides them):
//   forward : Q | K | V | mask | barriers ; P (32 KB) overwrites Q|K after S = QK^T has completed   -> 4 CTAs / SM
//   backward: Q | K | dO | V | X | dS | barriers ; P (32 KB) = V|X, written after dP = dO V^T completed -> 2 CTAs / SM
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + TQ * 128;

constexpr int SM_DO = SM_K + NK * 128;
constexpr int BWD_SM_V = SM_DO + TQ * 128;
constexpr int BWD_SM_P = BWD_SM_V;                   // aliases V | X
constexpr int SM_DS = BWD_SM_V + 2 * NK * 128;
constexpr int BWD_SM_BAR = SM_DS + TQ * NK * 2;
constexpr int BWD_SM_MASK = BWD_SM_BAR + 128;        // additive mask of this key tile (fp32, -inf past the sequence)
constexpr int BWD_SMEM = BWD_SM_MASK + NK * 4;
static_assert(BWD_SMEM <= 113 * 1024, "two backward CTAs must fit one SM");

__device__ __forceinline__ uint8_t* aligned_smem(uint8_t* raw) {
  if ((smem_u32(raw) & 1023u) != 0) __trap();  // 128B-swizzled tiles need a 1024-byte aligned base
  return raw;
}

// Write 32 consecutive keys (c0 .. c0+31) of row r into a K-major 128B-swizzled [128 x NK] bf16 tile.
__device__ __forceinline__ void store_tile_row32(uint8_t* tile, int r, int c0, const float (&x)[32]) {
  uint8_t* atom = tile + (c0 >> 6) * (TQ * 128) + r * 128;
  const int j0 = (c0 & 63) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 o;
    o.x = pack_bf16x2(x[g * 8 + 0], x[g * 8 + 1]);
    o.y = pack_bf16x2(x[g * 8 + 2], x[g * 8 + 3]);
    o.z = pack_bf16x2(x[g * 8 + 4], x[g * 8 + 5]);
    o.w = pack_bf16x2(x[g * 8 + 6], x[g * 8 + 7]);
    *reinterpret_cast<uint4*>(atom + (((j0 + g) ^ (r & 7)) << 4)) = o;
  }
}

// Eight consecutive keys (c0 .. c0+7, c0 % 8 == 0) of row r: one 16-byte piece of the swizzled tile.
__device__ __forceinline__ void store_tile_row8(uint8_t* tile, int r, int c0, const float (&x)[8]) {
  uint8_t* atom = tile + (c0 >> 6) * (TQ * 128) + r * 128;
  uint4 o;
  o.x = pack_bf16x2(x[0], x[1]);
  o.y = pack_bf16x2(x[2], x[3]);
  o.z = pack_bf16x2(x[4], x[5]);
  o.w = pack_bf16x2(x[6], x[7]);
  *reinterpret_cast<uint4*>(atom + ((((c0 & 63) >> 3) ^ (r & 7)) << 4)) = o;
}

// exp2 on the MUFU unit
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------
// forward.  NKT = number of 128-key tiles (1: S <= 128, four CTAs per SM; 2: S <= 256, two CTAs per SM).
// grid = (heads, batch, ceil(S / 128) query tiles).
// shared memory: Q (16 KB) | K (NKT x 16 KB) | pad (P needs NKT x 32 KB and overwrites Q | K | pad) | V | mask | barriers
// ------------------------------------------------------------------------------------------------
template <int NKT>
struct FwdCfg {
  static constexpr int NKEYS = 128 * NKT;
  static constexpr int P_BYTES = TQ * NKEYS * 2;
  static constexpr int QK_BYTES = TQ * 128 + NKEYS * 128;
  static constexpr int SM_V = P_BYTES > QK_BYTES ? P_BYTES : QK_BYTES;
  static constexpr int SM_MASK = SM_V + NKEYS * 128;
  static constexpr int SM_BAR = SM_MASK + NKEYS * 4;
  static constexpr int SMEM = SM_BAR + 64;
  static constexpr uint32_t TMEM_COLS = NKEYS;  // S: [0, NKEYS) ; O re-uses [0,64) once every thread has consumed S
};

template <int NKT>
__global__ void __launch_bounds__(128, NKT == 1 ? 4 : 2)
mhsa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const MhsaParams p) {
  using C = FwdCfg<NKT>;
  constexpr int NKEYS = C::NKEYS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = aligned_smem(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);  // [0] load, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* smask = reinterpret_cast<float*>(smem + C::SM_MASK);

  const int t = threadIdx.x, warp = t >> 5;
  const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * TQ;
  const int row0 = b * p.S;
  pdl_trigger();

  if (t == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();
  for (int c = t; c < NKEYS; c += 128) {
    float m = -INFINITY;  // keys beyond this sample's sequence (neighbouring sample / out of bounds) never contribute
    if (c < p.S) m = p.add_mask ? p.add_mask[(size_t)b * p.S + c] : 0.0f;
    smask[c] = m;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (t == 0) {
    const uint32_t bl = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(bl, (TQ + 2 * NKEYS) * 128);
    tma_load_2d(smem_u32(smem + SM_Q), &tm_q, bl, h * D_HEAD, row0 + q0);
    tma_load_2d(smem_u32(smem + SM_K), &tm_kv, bl, p.H + h * D_HEAD, row0);
    tma_load_2d(smem_u32(smem + C::SM_V), &tm_kv, bl, 2 * p.H + h * D_HEAD, row0);
    mbar_wait(bl, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = make_idesc_bf16(TQ, NKEYS, 0, 0);
    const uint32_t sq = smem_u32(smem + SM_Q), sk = smem_u32(smem + SM_K);
#pragma unroll
    for (int k = 0; k < D_HEAD / 16; ++k)
      umma_bf16_ss(tmem, make_smem_desc_sw128(sq + k * 32, 16, 1024), make_smem_desc_sw128(sk + k * 32, 16, 1024),
                   idesc_s, k > 0);
    umma_commit(smem_u32(&bars[1]));
  }

  // Only the key columns that exist are processed: ncols = S rounded up to 8 (the tail of the last 32-column chunk is read
  // with 8-column TMEM loads); P is written -- with zeros -- up to the next multiple of 16, which is as far as the second
  // product's k loop runs (S = 101: 104 of 128 columns of softmax work, 7 of 8 k-steps).  The arithmetic runs on the packed
  // fp32 pipe: t = s * scale + mask (FFMA2), p = 2^((t - m) log2 e) (FFMA2 + MUFU), row sum (FADD2).
  const bool dropping = p.drop.thresh != 0u && q0 + t < p.S;
  const uint64_t mrow = ((uint64_t)b * p.heads + h) * (uint64_t)p.S + (uint64_t)(q0 + t);
  const int wpr = (p.S + 31) >> 5;
  const int ncols = min(NKEYS, (p.S + 7) & ~7);
  const int nfull = ncols >> 5, ntail = (ncols & 31) >> 3;     // 32-column chunks, then 8-column groups
  const int kcols = (ncols + 15) & ~15;                        // extent of the P V product's reduction
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float2 sc2 = make_float2(p.scale, p.scale);
  // pass 1: row maximum of scale * s + mask
  float m = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < nfull; ++c) {
    uint32_t v[32];
    tmem_ld32(t_row + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 tt = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, *reinterpret_cast<const float2*>(smask + c * 32 + j));
      m = fmaxf(m, fmaxf(tt.x, tt.y));
    }
  }
#pragma unroll 1
  for (int gq = 0; gq < ntail; ++gq) {
    uint32_t v[8];
    const int c0 = nfull * 32 + gq * 8;
    tmem_ld8(t_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float2 tt = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, *reinterpret_cast<const float2*>(smask + c0 + j));
      m = fmaxf(m, fmaxf(tt.x, tt.y));
    }
  }
  // pass 2: p = exp(t - m), row sum, stage P (bf16) for the second product.  P overwrites the Q|K tiles (the S MMAs
  // that read them completed before bars[1] fired).  Dropout zeroes entries here; its 1 / (1 - p) is folded into the final
  // normalisation (the row sum and the saved log-sum-exp stay those of the full softmax).
  constexpr float LOG2E = 1.4426950408889634f;
  const float2 l2e = make_float2(LOG2E, LOG2E);
  const float2 nm2 = make_float2(-m * LOG2E, -m * LOG2E);
  float2 l2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
  for (int c = 0; c < nfull; ++c) {
    uint32_t v[32];
    float x[32];
    const uint32_t kb = dropping ? keep_word(p.drop, mrow, wpr, c) : 0xFFFFFFFFu;   // (in flight under the TMEM load)
    tmem_ld32(t_row + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 tt = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, *reinterpret_cast<const float2*>(smask + c * 32 + j));
      const float2 ee = __ffma2_rn(tt, l2e, nm2);
      const float2 pp = make_float2(ex2f(ee.x), ex2f(ee.y));
      l2 = __fadd2_rn(l2, pp);
      x[j] = ((kb >> j) & 1u) ? pp.x : 0.0f;
      x[j + 1] = ((kb >> (j + 1)) & 1u) ? pp.y : 0.0f;
    }
    store_tile_row32(smem, t, c * 32, x);  // P tile starts at offset 0 (aliases Q | K)
  }
  {
    const uint32_t kbt = (dropping && ntail > 0) ? keep_word(p.drop, mrow, wpr, nfull) : 0xFFFFFFFFu;
#pragma unroll 1
    for (int gq = 0; gq < ntail; ++gq) {
      uint32_t v[8];
      float x[8];
      const int c0 = nfull * 32 + gq * 8;
      tmem_ld8(t_row + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float2 tt = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, *reinterpret_cast<const float2*>(smask + c0 + j));
        const float2 ee = __ffma2_rn(tt, l2e, nm2);
        const float2 pp = make_float2(ex2f(ee.x), ex2f(ee.y));
        l2 = __fadd2_rn(l2, pp);
        x[j] = ((kbt >> (gq * 8 + j)) & 1u) ? pp.x : 0.0f;
        x[j + 1] = ((kbt >> (gq * 8 + j + 1)) & 1u) ? pp.y : 0.0f;
      }
      store_tile_row8(smem, t, c0, x);
    }
    if (kcols > ncols) {   // the k loop of P V runs in steps of 16 keys: the eight columns past ncols must be zero
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store_tile_row8(smem, t, ncols, z);
    }
  }
  const float l = l2.x + l2.y;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (t == 0) {
    tc_fence_after();
    constexpr uint32_t idesc_o = make_idesc_bf16(TQ, D_HEAD, 0, 1);
    const uint32_t sp = smem_u32(smem), sv = smem_u32(smem + C::SM_V);
#pragma unroll 1
    for (int j = 0; j < kcols / 16; ++j)
      umma_bf16_ss(tmem, make_smem_desc_sw128(sp + (j >> 2) * (TQ * 128) + (j & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(sv + j * 2048, 8192, 1024), idesc_o, j > 0);
    umma_commit(smem_u32(&bars[2]));
  }
  mbar_wait(smem_u32(&bars[2]), 0);
  tc_fence_after();
  {
    const float inv = (dropping ? p.drop.scale : 1.0f) / l;
    uint32_t v0[32], v1[32];
    tmem_ld32(t_row, v0);
    tmem_ld32(t_row + 32, v1);
    tmem_ld_wait();
    if (q0 + t < p.S) {
      __nv_bfloat16* dst = p.ctx + (size_t)(row0 + q0 + t) * p.H + h * D_HEAD;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v0[g * 8 + 0]) * inv, __uint_as_float(v0[g * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v0[g * 8 + 2]) * inv, __uint_as_float(v0[g * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v0[g * 8 + 4]) * inv, __uint_as_float(v0[g * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v0[g * 8 + 6]) * inv, __uint_as_float(v0[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = o;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v1[g * 8 + 0]) * inv, __uint_as_float(v1[g * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v1[g * 8 + 2]) * inv, __uint_as_float(v1[g * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v1[g * 8 + 4]) * inv, __uint_as_float(v1[g * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v1[g * 8 + 6]) * inv, __uint_as_float(v1[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + 32 + g * 8) = o;
      }
      if (p.lse) p.lse[((size_t)b * p.heads + h) * p.S + q0 + t] = m + __logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Column sums over the warp's 32 rows of a 32-column accumulator chunk (thread = row): a transposing butterfly -- in step
// `off` every lane keeps the half of its values whose column bit equals its own lane bit and adds the partner's -- leaves
// lane L with the sum of column L after 16+8+4+2+1 = 31 shuffles; one atomic per lane.  (bias gradient of the QKV Linear)
__device__ __forceinline__ void colsum32_atomic(float* dst, const uint32_t (&v)[32], bool row_valid, int lane) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = row_valid ? __uint_as_float(v[i]) : 0.0f;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float mine = up ? a[i + off] : a[i];
      const float send = up ? a[i] : a[i + off];
      a[i] = mine + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  atomicAdd(dst + lane, a[0]);
}

// 32 accumulator columns of one row -> 32 bf16 values in global memory
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, uint32_t tcol_addr, float* colsum = nullptr, int lane = 0) {
  uint32_t v0[32];
  tmem_ld32(tcol_addr, v0);
  tmem_ld_wait();
  if (colsum != nullptr) colsum32_atomic(colsum, v0, dst != nullptr, lane);
  if (dst != nullptr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v0[g * 8 + 0]), __uint_as_float(v0[g * 8 + 1]));
      o.y = pack_bf16x2(__uint_as_float(v0[g * 8 + 2]), __uint_as_float(v0[g * 8 + 3]));
      o.z = pack_bf16x2(__uint_as_float(v0[g * 8 + 4]), __uint_as_float(v0[g * 8 + 5]));
      o.w = pack_bf16x2(__uint_as_float(v0[g * 8 + 6]), __uint_as_float(v0[g * 8 + 7]));
      *reinterpret_cast<uint4*>(dst + g * 8) = o;
    }
  }
}

// 32 accumulator columns of one row -> fp32 atomic accumulation (multi-block backward, S > 128)
__device__ __forceinline__ void add_row32(float* dst, uint32_t tcol_addr, float* colsum = nullptr, int lane = 0) {
  uint32_t v0[32];
  tmem_ld32(tcol_addr, v0);
  tmem_ld_wait();
  if (colsum != nullptr) colsum32_atomic(colsum, v0, dst != nullptr, lane);
  if (dst != nullptr) {
#pragma unroll
    for (int g = 0; g < 8; ++g)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(__uint_as_float(v0[g * 4])),
                   "f"(__uint_as_float(v0[g * 4 + 1])), "f"(__uint_as_float(v0[g * 4 + 2])), "f"(__uint_as_float(v0[g * 4 + 3]))
                   : "memory");
  }
}

// 256 threads: warps w and w+4 share TMEM lane quarter (w % 4); thread pair (t, t+128) owns query/key row t % 128 and
// splits the 128 key columns (softmax / dS phase) resp. the 64 head-dim columns (store phase) in halves.
constexpr int BWD_THREADS = 256;

__global__ void __launch_bounds__(BWD_THREADS, 2) mhsa_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                  const __grid_constant__ CUtensorMap tm_dctx, const MhsaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = aligned_smem(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BWD_SM_BAR);  // [0] load, [1] S & dP ready, [2] grads ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int t = tid & 127;        // row owned by this thread
  const int half = tid >> 7;      // which half of the columns
  const int h = blockIdx.x, b = blockIdx.y;
  const int q0 = (blockIdx.z / p.ktiles) * TQ, k0 = (blockIdx.z % p.ktiles) * NK;  // this CTA's (query tile, key tile) block
  const int row0 = b * p.S;
  pdl_trigger();
  // S [0,128) dP [128,256); once every thread has consumed them: dQ [0,64) dK [64,128) dV [128,192)
  constexpr uint32_t TMEM_COLS = 256;
  constexpr uint32_t T_DQ = 0, T_DK = 64, T_DV = 128;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_dctx);
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();
  float* smask = reinterpret_cast<float*>(smem + BWD_SM_MASK);
  if (tid < NK) {
    const int col = k0 + tid;
    smask[tid] = col < p.S ? (p.add_mask ? p.add_mask[(size_t)b * p.S + col] : 0.0f) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const uint32_t sq = smem_u32(smem + SM_Q), sk = smem_u32(smem + SM_K), sv = smem_u32(smem + BWD_SM_V);
  const uint32_t sp = smem_u32(smem + BWD_SM_P), sdo = smem_u32(smem + SM_DO), sds = smem_u32(smem + SM_DS);

  if (tid == 0) {
    const uint32_t bl = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(bl, (2 * TQ + 2 * NK) * 128);
    tma_load_2d(sq, &tm_qkv, bl, h * D_HEAD, row0 + q0);
    tma_load_2d(sk, &tm_qkv, bl, p.H + h * D_HEAD, row0 + k0);
    tma_load_2d(sv, &tm_qkv, bl, 2 * p.H + h * D_HEAD, row0 + k0);
    tma_load_2d(sdo, &tm_dctx, bl, h * D_HEAD, row0 + q0);
    mbar_wait(bl, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(TQ, NK, 0, 0);
#pragma unroll
    for (int k = 0; k < D_HEAD / 16; ++k)  // S = Q K^T
      umma_bf16_ss(tmem, make_smem_desc_sw128(sq + k * 32, 16, 1024), make_smem_desc_sw128(sk + k * 32, 16, 1024),
                   idesc, k > 0);
#pragma unroll
    for (int k = 0; k < D_HEAD / 16; ++k)  // dP = dO V^T
      umma_bf16_ss(tmem + 128, make_smem_desc_sw128(sdo + k * 32, 16, 1024),
                   make_smem_desc_sw128(sv + k * 32, 16, 1024), idesc, k > 0);
    umma_commit(smem_u32(&bars[1]));
  }

  // D = rowsum(dO o O) and the saved log-sum-exp for this query row (overlaps the TMA + MMAs above).
  // The warp reads its 32 rows cooperatively -- lane l takes the 16-byte piece l % 8 of rows l / 8 + 4 j, so every load
  // instruction covers four whole 128-byte rows instead of 32 scattered sectors (the one-row-per-thread version stalled on
  // the load queue: 16 uncoalesced 128-bit loads per thread) -- and the eight partial dot products of a row meet by shuffles.
  float Dsum = 0.0f, lse = 0.0f;
  const bool valid = q0 + t < p.S;     // this thread's QUERY row exists
  const bool kvalid = k0 + t < p.S;    // this thread's KEY row exists (dK / dV rows)
  {
    const int lane = tid & 31;
    const int wrow0 = q0 + (t & ~31);                       // first query row of this warp's lane quarter
    const int piece = lane & 7, rgrp = lane >> 3;
    uint4 a[8], d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = wrow0 + rgrp + 4 * j;
      if (r < p.S) {
        a[j] = __ldg(reinterpret_cast<const uint4*>(p.ctx + (size_t)(row0 + r) * p.H + h * D_HEAD) + piece);
        d[j] = __ldg(reinterpret_cast<const uint4*>(p.dctx + (size_t)(row0 + r) * p.H + h * D_HEAD) + piece);
      } else {
        a[j] = d[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    if (valid) lse = p.lse[((size_t)b * p.heads + h) * p.S + q0 + t];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t aa[4] = {a[j].x, a[j].y, a[j].z, a[j].w}, dd[4] = {d[j].x, d[j].y, d[j].z, d[j].w};
      float part = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) part += bf16lo(aa[i]) * bf16lo(dd[i]) + bf16hi(aa[i]) * bf16hi(dd[i]);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      // row rgrp + 4 j is now complete in lanes 8 rgrp .. 8 rgrp + 7; lane L owns row L = 4 (L / 4) + L % 4
      const float mine = __shfl_sync(0xffffffffu, part, (lane & 3) * 8);
      if ((lane >> 2) == j) Dsum = mine;
    }
  }

  // Softmax recomputation and dS on the key columns that exist: ncols = this tile's keys rounded up to 8, in groups of eight
  // columns split evenly between the two threads of a row (S = 101: 13 groups, 7 + 6; the 128-column version did 8 + 8).
  // Packed fp32 arithmetic: t = s * scale + mask, p = 2^((t - lse) log2 e), P' = p o M', dS = p * scale * (dP o M' - D).
  // Query rows past the sequence get lse = +inf -> p = 0: they contribute nothing to dK / dV.
  const bool dropping = p.drop.thresh != 0u && valid;
  const uint64_t mrow = ((uint64_t)b * p.heads + h) * (uint64_t)p.S + (uint64_t)(q0 + t);
  const int wpr = (p.S + 31) >> 5;
  const int keys_here = max(0, min(NK, p.S - k0));
  const int ncols = (keys_here + 7) & ~7;
  const int kcols = (ncols + 15) & ~15;                  // extent of dQ = dS K's reduction (zero-filled past ncols)
  const int ngroups = ncols >> 3;
  const int g_split = (ngroups + 1) >> 1;
  const int g_begin = half ? g_split : 0, g_end = half ? ngroups : g_split;
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  {
    constexpr float LOG2E = 1.4426950408889634f;
    const float2 sc2 = make_float2(p.scale, p.scale), l2e = make_float2(LOG2E, LOG2E);
    const float nl = valid ? -lse * LOG2E : -INFINITY;
    const float2 nl2 = make_float2(nl, nl), nD2 = make_float2(-Dsum, -Dsum);
    const float keep_mul = dropping ? p.drop.scale : 1.0f;
#pragma unroll 1
    for (int gq = g_begin; gq < g_end; ++gq) {
      const int c0 = gq * 8;
      uint32_t vs[8], vd[8];
      float pr[8], ds[8];
      const uint32_t kbits = dropping ? (keep_word(p.drop, mrow, wpr, (k0 + c0) >> 5) >> ((k0 + c0) & 31)) : 0xFFu;
      tmem_ld8(t_row + c0, vs);
      tmem_ld8(t_row + 128 + c0, vd);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float2 tt = __ffma2_rn(make_float2(__uint_as_float(vs[j]), __uint_as_float(vs[j + 1])), sc2, *reinterpret_cast<const float2*>(smask + c0 + j));
        const float2 ee = __ffma2_rn(tt, l2e, nl2);
        const float2 pp = make_float2(ex2f(ee.x), ex2f(ee.y));
        const float2 k2 = make_float2(((kbits >> j) & 1u) ? keep_mul : 0.0f, ((kbits >> (j + 1)) & 1u) ? keep_mul : 0.0f);
        const float2 pm = __fmul2_rn(pp, k2);
        const float2 u = __ffma2_rn(make_float2(__uint_as_float(vd[j]), __uint_as_float(vd[j + 1])), k2, nD2);
        const float2 dsv = __fmul2_rn(__fmul2_rn(pp, sc2), u);
        pr[j] = pm.x; pr[j + 1] = pm.y;
        ds[j] = dsv.x; ds[j + 1] = dsv.y;
      }
      store_tile_row8(smem + BWD_SM_P, t, c0, pr);   // P overwrites V (+ spare): dP = dO V^T completed before bars[1]
      store_tile_row8(smem + SM_DS, t, c0, ds);
    }
    if (half && kcols > ncols) {   // dQ's k loop runs in steps of 16 keys
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store_tile_row8(smem + SM_DS, t, ncols, z);
      store_tile_row8(smem + BWD_SM_P, t, ncols, z);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc_q = make_idesc_bf16(TQ, D_HEAD, 0, 1);   // dQ = dS K      (A K-major, B MN-major)
    constexpr uint32_t idesc_kv = make_idesc_bf16(NK, D_HEAD, 1, 1);  // dK = dS^T Q, dV = P^T dO (both MN-major)
#pragma unroll 1
    for (int j = 0; j < kcols / 16; ++j)
      umma_bf16_ss(tmem + T_DQ, make_smem_desc_sw128(sds + (j >> 2) * (TQ * 128) + (j & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(sk + j * 2048, 8192, 1024), idesc_q, j > 0);
#pragma unroll
    for (int j = 0; j < TQ / 16; ++j)
      umma_bf16_ss(tmem + T_DK, make_smem_desc_sw128(sds + j * 2048, TQ * 128, 1024),
                   make_smem_desc_sw128(sq + j * 2048, 8192, 1024), idesc_kv, j > 0);
#pragma unroll
    for (int j = 0; j < TQ / 16; ++j)
      umma_bf16_ss(tmem + T_DV, make_smem_desc_sw128(sp + j * 2048, TQ * 128, 1024),
                   make_smem_desc_sw128(sdo + j * 2048, 8192, 1024), idesc_kv, j > 0);
    umma_commit(smem_u32(&bars[2]));
  }
  mbar_wait(smem_u32(&bars[2]), 0);
  tc_fence_after();
  {
    // each thread of the pair handles 32 of the 64 head-dim columns: dQ of query row q0+t, dK / dV of key row k0+t
    const size_t col = (size_t)h * D_HEAD + half * 32;
    const int lane = tid & 31;
    float* cq = p.dbias ? p.dbias + col : nullptr;               // fused bias gradients: db_q / db_k / db_v += column sums
    float* ck = p.dbias ? p.dbias + p.H + col : nullptr;
    float* cv = p.dbias ? p.dbias + 2 * p.H + col : nullptr;
    if (p.dqkv_f32 == nullptr) {  // single block: final values, bf16
      store_row32(valid ? p.dqkv + (size_t)(row0 + q0 + t) * (3 * p.H) + col : nullptr, t_row + T_DQ + half * 32, cq, lane);
      store_row32(kvalid ? p.dqkv + (size_t)(row0 + k0 + t) * (3 * p.H) + p.H + col : nullptr, t_row + T_DK + half * 32, ck, lane);
      store_row32(kvalid ? p.dqkv + (size_t)(row0 + k0 + t) * (3 * p.H) + 2 * p.H + col : nullptr, t_row + T_DV + half * 32, cv, lane);
    } else {                      // partial sums over the other tile dimension: fp32 atomics, converted afterwards
      add_row32(valid ? p.dqkv_f32 + (size_t)(row0 + q0 + t) * (3 * p.H) + col : nullptr, t_row + T_DQ + half * 32, cq, lane);
      add_row32(kvalid ? p.dqkv_f32 + (size_t)(row0 + k0 + t) * (3 * p.H) + p.H + col : nullptr, t_row + T_DK + half * 32, ck, lane);
      add_row32(kvalid ? p.dqkv_f32 + (size_t)(row0 + k0 + t) * (3 * p.H) + 2 * p.H + col : nullptr, t_row + T_DV + half * 32, cv, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

}  // namespace

template <int NKT>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tkv, const MhsaParams& p, int B, int heads, int qtiles, cudaStream_t stream) {
  using C = FwdCfg<NKT>;
  static bool attr = false;
  if (!attr) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(mhsa_fwd_kernel<NKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  VLB_CHECK_CUDA(launch_pdl(mhsa_fwd_kernel<NKT>, dim3(heads, B, qtiles), dim3(128), C::SMEM, stream, tq, tkv, p));
  return VLB_OK;
}

int mhsa_forward(const void* qkv, const float* add_mask, void* ctx, float* lse, int B, int S, int H, int heads,
                 cudaStream_t stream, const VlbDropout* drop) {
  VLB_REQUIRE(qkv && ctx, "mhsa_forward: null pointer");
  VLB_REQUIRE(drop_valid(drop), "mhsa_forward: bad dropout configuration");
  VLB_REQUIRE(drop == nullptr || drop->p == 0.0f || drop->keep_bits != nullptr, "mhsa_forward: dropout needs precomputed keep bits (vlb_dropout_bits over [B*heads*S, S])");
  VLB_REQUIRE(H == heads * D_HEAD, "mhsa: head size must be 64 (H=%d heads=%d)", H, heads);
  VLB_REQUIRE(S >= 1 && S <= 256, "mhsa: sequence length %d not supported (1..256)", S);
  MhsaParams p{};
  p.B = B; p.S = S; p.H = H; p.heads = heads;
  p.scale = 0.125f;
  p.add_mask = add_mask;
  p.ctx = static_cast<__nv_bfloat16*>(ctx);
  p.lse = lse;
  p.drop = make_drop(drop);
  const int nkt = S <= 128 ? 1 : 2;
  CUtensorMap tq, tkv;
  int rc = make_tmap_bf16_2d(&tq, qkv, (uint64_t)B * S, 3 * H, 3 * H, 64, 128);
  if (rc != VLB_OK) return rc;
  rc = make_tmap_bf16_2d(&tkv, qkv, (uint64_t)B * S, 3 * H, 3 * H, 64, 128 * nkt);
  if (rc != VLB_OK) return rc;
  ProfScope prof(PROF_MHSA_FWD, 4.0 * B * heads * (double)S * S * D_HEAD, stream);
  const int qtiles = (S + TQ - 1) / TQ;
  return nkt == 1 ? launch_fwd<1>(tq, tkv, p, B, heads, qtiles, stream) : launch_fwd<2>(tq, tkv, p, B, heads, qtiles, stream);
}

int cast_f32_to_bf16(const float* in, void* out, size_t n, cudaStream_t stream);

int mhsa_backward(const void* qkv, const float* add_mask, const void* ctx, const float* lse, const void* dctx, void* dqkv,
                  float* scratch_f32, int B, int S, int H, int heads, cudaStream_t stream, const VlbDropout* drop, float* dbias_qkv) {
  VLB_REQUIRE(qkv && ctx && lse && dctx && dqkv, "mhsa_backward: null pointer");
  VLB_REQUIRE(drop_valid(drop), "mhsa_backward: bad dropout configuration");
  VLB_REQUIRE(drop == nullptr || drop->p == 0.0f || drop->keep_bits != nullptr, "mhsa_backward: dropout needs the keep bits the forward used");
  VLB_REQUIRE(H == heads * D_HEAD, "mhsa: head size must be 64 (H=%d heads=%d)", H, heads);
  VLB_REQUIRE(S >= 1 && S <= 256, "mhsa: sequence length %d not supported (1..256)", S);
  const int tiles = (S + TQ - 1) / TQ;
  VLB_REQUIRE(tiles == 1 || scratch_f32 != nullptr, "mhsa_backward: S > 128 needs the fp32 scratch buffer [B*S, 3H]");
  MhsaParams p{};
  p.B = B; p.S = S; p.H = H; p.heads = heads;
  p.scale = 0.125f;
  p.add_mask = add_mask;
  p.ctx = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(ctx));
  p.lse = const_cast<float*>(lse);
  p.dctx = static_cast<const __nv_bfloat16*>(dctx);
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.dqkv_f32 = tiles == 1 ? nullptr : scratch_f32;
  p.ktiles = tiles;
  p.drop = make_drop(drop);
  p.dbias = dbias_qkv;
  CUtensorMap tm, tmd;
  int rc = make_tmap_bf16_2d(&tm, qkv, (uint64_t)B * S, 3 * H, 3 * H, 64, 128);
  if (rc != VLB_OK) return rc;
  rc = make_tmap_bf16_2d(&tmd, dctx, (uint64_t)B * S, H, H, 64, 128);
  if (rc != VLB_OK) return rc;
  static bool attr = false;
  if (!attr) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(mhsa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr = true;
  }
  const size_t n = (size_t)B * S * 3 * H;
  if (tiles > 1) VLB_CHECK_CUDA(cudaMemsetAsync(scratch_f32, 0, n * sizeof(float), stream));
  {
    ProfScope prof(PROF_MHSA_BWD, 8.0 * B * heads * (double)S * S * D_HEAD, stream);
    VLB_CHECK_CUDA(launch_pdl(mhsa_bwd_kernel, dim3(heads, B, tiles * tiles), dim3(BWD_THREADS), BWD_SMEM, stream, tm, tmd, p));
  }
  if (tiles > 1) return cast_f32_to_bf16(scratch_f32, dqkv, n, stream);
  return VLB_OK;
}

}  // namespace vlb
