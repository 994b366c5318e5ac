// tcgen05 GEMM for sm_100a: persistent, warp-specialised (TMA producer / single-thread MMA issuer /
// 4 epilogue warps), 128 x BN x 64 tiles, bf16 operands staged by TMA into 128B-swizzled shared
// memory, fp32 accumulators double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Fused epilogues cover everything the VL-BERT encoder layer needs around its GEMMs
// (reference: external/pytorch_pretrained_bert/modeling.py:291-293 QKV, :330-333 output dense +
// residual, :362-363 intermediate dense + erf-GELU, :375-378 output dense + residual) and their
// backward passes (dgrad with fused GELU' / residual-gradient add, wgrad with split-K fp32
// reduction).
#include <mutex>
#include <unordered_map>

#include "gemm_sm100.cuh"

namespace vlb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;  // warp0 = TMA, warp1 = MMA + TMEM alloc, warps 2..5 = epilogue

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int split_k, kb_per_split;
  int num_items;
  // descriptor geometry (bytes); see make_smem_desc_sw128
  uint32_t a_lbo, a_sbo, a_kadv;
  uint32_t b_lbo, b_sbo, b_kadv;
  GemmEpilogue e;
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// One 32-column slice of a row of the accumulator -> global memory, with the fused epilogue.
__device__ __forceinline__ void epilogue_store32(const GemmEpilogue& e, const uint32_t (&v)[32], int row,
                                                 int col0, int N) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col >= N) break;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]) * e.alpha;
    if (e.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + col + 4));
      x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
      x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
    }
    if (e.act == ACT_GELU) {
      if (e.aux != nullptr) {
        uint4 z;
        z.x = pack_bf16x2(x[0], x[1]); z.y = pack_bf16x2(x[2], x[3]);
        z.z = pack_bf16x2(x[4], x[5]); z.w = pack_bf16x2(x[6], x[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.aux) + (size_t)row * e.ld_aux + col) = z;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
    } else if (e.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.0f);
    } else if (e.act == ACT_DGELU_MUL || e.act == ACT_DRELU_MUL) {
      const uint4 z = *reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(e.aux) + (size_t)row * e.ld_aux + col);
      const uint32_t zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z0 = bf16lo(zz[j]), z1 = bf16hi(zz[j]);
        if (e.act == ACT_DGELU_MUL) {
          x[2 * j] *= gelu_erf_grad(z0);
          x[2 * j + 1] *= gelu_erf_grad(z1);
        } else {
          x[2 * j] = z0 > 0.0f ? x[2 * j] : 0.0f;
          x[2 * j + 1] = z1 > 0.0f ? x[2 * j + 1] : 0.0f;
        }
      }
    }
    if (e.resid_kind == RESID_BF16) {
      const uint4 r = *reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(e.resid) + (size_t)row * e.ldr + col);
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[2 * j] += bf16lo(rr[j]);
        x[2 * j + 1] += bf16hi(rr[j]);
      }
    } else if (e.resid_kind == RESID_F32) {
      const float* rp = reinterpret_cast<const float*>(e.resid) + (size_t)row * e.ldr + col;
      const float4 r0 = *reinterpret_cast<const float4*>(rp);
      const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
      x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
      x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
    }
    if (e.out_kind == OUT_BF16) {
      uint4 o;
      o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
      o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)row * e.ldo + col) = o;
    } else {
      float* op = reinterpret_cast<float*>(e.out) + (size_t)row * e.ldo + col;
      if (e.out_kind == OUT_F32) {
        *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
      } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op), "f"(x[0]), "f"(x[1]),
                     "f"(x[2]), "f"(x[3])
                     : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + 4), "f"(x[4]), "f"(x[5]),
                     "f"(x[6]), "f"(x[7])
                     : "memory");
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * C::STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tfull_bar[0]), 1);
    mbar_init(smem_u32(&tfull_bar[1]), 1);
    mbar_init(smem_u32(&tempty_bar[0]), 4);
    mbar_init(smem_u32(&tempty_bar[1]), 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int items_per_split = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int split = item / items_per_split;
        const int rem = item - split * items_per_split;
        const int m_blk = rem / p.num_n_blocks;
        const int n_blk = rem - m_blk * p.num_n_blocks;
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d(sa, &tma_a, fb, k0, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tma_a, fb, m0 + c * 64, k0);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tma_b, fb, k0, n0);  // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tma_b, fb, n0 + c * 64, k0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int split = item / items_per_split;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
        const int buf = it & 1;
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        mbar_wait(smem_u32(&tempty_bar[buf]), (use & 1u) ^ 1u);
        tc_fence_after();
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
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&tfull_bar[buf]));  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int rem = item % items_per_split;
      const int m_blk = rem / p.num_n_blocks;
      const int n_blk = rem - m_blk * p.num_n_blocks;
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(smem_u32(&tfull_bar[buf]), use & 1u);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row < p.M && col0 < p.N) epilogue_store32(p.e, v, row, col0, p.N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

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

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  gemm_kernel<BN, A_MN, B_MN><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  VLB_CHECK_LAUNCH();
  return VLB_OK;
}

// debug override of the MN-major descriptor geometry (bring-up aid, see tools/gemm_probe.py)
uint32_t g_dbg_mn_lbo = 0, g_dbg_mn_sbo = 0, g_dbg_mn_kadv = 0;

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

void gemm_debug_override(uint32_t mn_lbo, uint32_t mn_sbo, uint32_t mn_kadv) {
  g_dbg_mn_lbo = mn_lbo;
  g_dbg_mn_sbo = mn_sbo;
  g_dbg_mn_kadv = mn_kadv;
}

int gemm_bf16(int mode, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
              const GemmEpilogue& epi, int split_k, int force_bn, cudaStream_t stream) {
  VLB_REQUIRE(mode >= GEMM_NT && mode <= GEMM_TN, "gemm: bad mode %d", mode);
  VLB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VLB_REQUIRE(A && B && epi.out, "gemm: null pointer");
  VLB_REQUIRE(N % 8 == 0, "gemm: N (%d) must be a multiple of 8", N);
  VLB_REQUIRE(epi.ldo % 4 == 0 && (epi.out_kind != OUT_BF16 || epi.ldo % 8 == 0), "gemm: bad ldo %d", epi.ldo);
  VLB_REQUIRE(epi.resid_kind == RESID_NONE || (epi.resid != nullptr && epi.ldr % 8 == 0), "gemm: bad residual");
  VLB_REQUIRE((epi.act != ACT_DGELU_MUL && epi.act != ACT_DRELU_MUL) || (epi.aux != nullptr && epi.ld_aux % 8 == 0),
              "gemm: activation-gradient epilogue needs aux");
  const bool a_mn = (mode == GEMM_TN);
  const bool b_mn = (mode != GEMM_NT);

  // Tile-N choice: fewest "rounds" of the persistent grid weighted by tile cost.
  int bn = 128;
  if (force_bn == 128 || force_bn == 256 || force_bn == 64) {
    bn = force_bn;
  } else {
    const int sms = num_sms();
    const int mb = (M + BM - 1) / BM;
    auto cost = [&](int b) {
      const long items = (long)mb * ((N + b - 1) / b) * (split_k > 1 ? split_k : 1);
      const long rounds = (items + sms - 1) / sms;
      return rounds * b;  // time ~ rounds * tile width
    };
    bn = 128;
    long best = cost(128);
    if (N >= 256 && cost(256) < best) { best = cost(256); bn = 256; }
    if (cost(64) < best) { best = cost(64); bn = 64; }
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + BM - 1) / BM;
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
  // K-major operand: rows x 128B, swizzle atoms of 8 rows -> SBO 1024B, 32B per UMMA_K step.
  // MN-major operand: boxes of 64(mn) x 64(k): k-groups of 8 rows at 1024B (SBO), next 64-wide mn
  // chunk at 8192B (LBO), 16 k-rows = 2048B per UMMA_K step.
  const uint32_t mn_lbo = g_dbg_mn_lbo ? g_dbg_mn_lbo : 8192u;
  const uint32_t mn_sbo = g_dbg_mn_sbo ? g_dbg_mn_sbo : 1024u;
  const uint32_t mn_kadv = g_dbg_mn_kadv ? g_dbg_mn_kadv : 2048u;
  p.a_lbo = a_mn ? mn_lbo : 16u;  p.a_sbo = a_mn ? mn_sbo : 1024u;  p.a_kadv = a_mn ? mn_kadv : 32u;
  p.b_lbo = b_mn ? mn_lbo : 16u;  p.b_sbo = b_mn ? mn_sbo : 1024u;  p.b_kadv = b_mn ? mn_kadv : 32u;

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_bf16_2d(&ta, A, M, K, lda, 64, BM);       // A [M, K]
  else       rc = make_tmap_bf16_2d(&ta, A, K, M, lda, 64, 64);       // A stored [K, M]
  if (rc != VLB_OK) return rc;
  if (!b_mn) rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, 64, bn);       // B [N, K]
  else       rc = make_tmap_bf16_2d(&tb, B, K, N, ldb, 64, 64);       // B stored [K, N]
  if (rc != VLB_OK) return rc;

  ProfScope prof(mode == GEMM_NT ? PROF_GEMM_NT : (mode == GEMM_NN ? PROF_GEMM_NN : PROF_GEMM_TN), 2.0 * M * N * K, stream);
#define VLB_GEMM_DISPATCH(BN_)                                                         \
  if (!a_mn && !b_mn) return launch<BN_, false, false>(ta, tb, p, stream);             \
  if (!a_mn && b_mn) return launch<BN_, false, true>(ta, tb, p, stream);               \
  return launch<BN_, true, true>(ta, tb, p, stream);
  if (bn == 256) { VLB_GEMM_DISPATCH(256) }
  if (bn == 64) { VLB_GEMM_DISPATCH(64) }
  VLB_GEMM_DISPATCH(128)
#undef VLB_GEMM_DISPATCH
}

}  // namespace vlb
