// CTA-pair GEMM with 192-row CTA tiles (included by gemm_sm100.cu after the epilogue definitions; same anonymous namespace).
//
// Why this shape.  The encoder's token matrix has M = B*S = 6464 rows at BASELINE config 2 and its N = 768 GEMMs
// (attention output / FFN-down forward, modeling.py:330,375, and the dctx / dh / dx data gradients) are 51 x 3 tiles of
// 128 x 256: 1.03 rounds of a 148-CTA persistent grid, and every 128 x 256 x 64 k-block pulls 48 KB through the L2->SM
// fabric for 4.2 MFLOP (11.4 B/kFLOP), which is what bounds the mainloop (~6300 B/clk chip-wide).  Here a cluster of two
// CTAs owns a 384 x BN tile: each CTA stages ITS 192 rows of A (24 KB per k-block) and ONE HALF of the B tile, and one
// elected thread of the leader CTA issues, per 16-wide k-step, a cta_group::2 MMA with M = 256 (rows 0..127 of both CTAs)
// plus one with M = 128 (rows 128..191 of both CTAs: 64 rows per CTA run at full rate only as half of a pair MMA).
// 6464 x 768 is then 17 x 4 = 68 pair tiles = 136 CTAs: ONE wave on 92 % of the SMs, at 7.6 B/kFLOP.
//
// TMEM per CTA (512 columns allocated):
//   acc1[b] = columns [b * BN, (b + 1) * BN):   lane l   <-> local row l (0..127), column c
//   acc2    = columns [A2, A2 + BN/2):   lane l (l < 64)   <-> local row 128 + l,        column c            (c < BN/2)
//                                        lane l (l >= 64)  <-> local row 128 + (l - 64), column BN/2 + c
//   (the "2x2" accumulator layout of a cta_group::2 MMA with 64 rows per CTA).
//   BN = 192: acc1 is double-buffered (2 x 192 columns) and A2 = 384; the epilogue drains acc2 first and releases it, so the
//   MMAs of the next tile start after a third of the epilogue and overlap the rest of it.  BN = 256: single buffers, A2 = 256.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM allocation, warps 2..9 = epilogue; the barrier
// protocol is the pair protocol of gemm_body<CM = 1>: full barriers live in the leader and count one arrive per CTA, MMA
// commits are multicast to both CTAs, the peer's epilogue warps release the accumulator on the leader's barrier.
//
// Operands: A K-major [M, K] (activations / their gradients); B K-major [N, K] (forward) or MN-major [K, N] (data gradient:
// the weight matrix as stored).  N % BN == 0.

template <int BN, bool B_MN>
struct PairCfg {
  static constexpr int ROWS = 192;                                    // rows of A per CTA
  static constexpr int A_BYTES = ROWS * BK * 2;                       // 24 KB
  static constexpr int B_HALF = BN / 2;                               // columns of the B tile staged by this CTA
  static constexpr int B_CHUNKS = (B_HALF + 63) / 64;                 // MN-major: 64-column boxes (the last one may be half used)
  static constexpr int B_BYTES = B_MN ? B_CHUNKS * 8192 : B_HALF * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * STAGE_F32_PER_WARP * 4;
  static constexpr int STAGES = (227 * 1024 - EPI_STAGE_BYTES - 1024 - 256) / STAGE_BYTES >= 5 ? 5 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr bool DB = 2 * BN + BN / 2 <= 512;                  // acc1 double-buffered
  static constexpr uint32_t ACC2_COL = DB ? 2 * BN : BN;
  static_assert(BN % 32 == 0 && BN + BN / 2 <= 512, "accumulators must fit the tensor memory");
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzled tiles");
};

struct PairParams {
  int M, N, K;
  int num_m_pairs, num_n_blocks, num_k_blocks;
  int num_items;
  unsigned long long* trace;   // timing aid (vlb_debug_gemm_trace): 8 globaltimer stamps per CTA, nullptr = off
  int roles_high;     // 1: producer / MMA issuer are the two highest hardware warps (scheduler priority, see gemm_body)
  int epi_stage;      // 1: last tile of a CTA: residual rows / keep flags / LayerNorm statistics staged in the idle operand ring
  int epi_prefetch;   // 1: the epilogue warps pull the tile's residual / saved-activation / keep-flag lines into L2 while its MMAs run
  GemmEpilogue e;
};

__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int slot) {
  if (trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * 8 + slot] = t;
  }
}

template <int BN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_pair192_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const PairParams p) {
  using C = PairCfg<BN, B_MN>;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int worker = blockIdx.x >> 1;
  const int nworkers = gridDim.x >> 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  (used in the leader)
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * C::STAGES;   // [2] accumulators of a tile complete (multicast commit), indexed by acc1 buffer
  uint64_t* tempty1_bar = tfull_bar + 2;        // [2] acc1[b] drained (leader's barriers, both CTAs' epilogue warps arrive)
  uint64_t* tempty2_bar = tempty1_bar + 2;      // acc2 drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty2_bar + 1);

  const int hw_warp = threadIdx.x >> 5;
  const int warp = p.roles_high ? (hw_warp >= EPI_WARPS ? hw_warp - EPI_WARPS : hw_warp + 2) : hw_warp;   // logical role, see gemm_body
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) trace_stamp(p.trace, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull_bar[b]), 1);
      mbar_init(smem_u32(&tempty1_bar[b]), 2 * EPI_WARPS);
    }
    mbar_init(smem_u32(tempty2_bar), 2 * EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(smem_u32(tmem_slot), C::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (threadIdx.x == 0) trace_stamp(p.trace, 1);

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < p.num_items; item += nworkers) {
        const int m_pair = item / p.num_n_blocks;
        const int n_blk = item - m_pair * p.num_n_blocks;
        const int m0 = (m_pair * 2 + (int)rank) * C::ROWS;
        const int n0 = n_blk * BN + (int)rank * C::B_HALF;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = mapa_cluster(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), C::STAGE_BYTES);
          else mbar_arrive_expect_tx_cluster(fb, C::STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const int k0 = kb * BK;
          tma_load_2d_cg2(sa, &tma_a, fb, k0, m0);                 // box {64 k, 192 rows}
          if (!B_MN) {
            tma_load_2d_cg2(sb, &tma_b, fb, k0, n0);               // box {64 k, BN/2 rows}
          } else {
#pragma unroll
            for (int c = 0; c < C::B_CHUNKS; ++c) tma_load_2d_cg2(sb + c * 8192, &tma_b, fb, n0 + c * 64, k0);   // box {64 n, 64 k}
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc256 = make_idesc_bf16(256, BN, 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc128 = make_idesc_bf16(128, BN, 0, B_MN ? 1 : 0);
      constexpr uint32_t b_lbo = B_MN ? 8192u : 16u, b_sbo = 1024u, b_kadv = B_MN ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int item = worker; item < p.num_items; item += nworkers, ++it) {
        const uint32_t buf = C::DB ? (it & 1u) : 0u;
        const uint32_t use = C::DB ? (it >> 1) : it;                     // how often this acc1 buffer has been used before
        mbar_wait(smem_u32(&tempty1_bar[buf]), (use & 1u) ^ 1u);
        mbar_wait(smem_u32(tempty2_bar), (it & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d1 = tmem_base + buf * BN, d2 = tmem_base + C::ACC2_COL;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          if (kb == 0 && it == 0) trace_stamp(p.trace, 2);
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t bdesc = make_smem_desc_sw128(sb + k * b_kadv, b_lbo, b_sbo);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_bf16_ss_cg2(d1, make_smem_desc_sw128(sa + k * 32, 16, 1024), bdesc, idesc256, acc);
            umma_bf16_ss_cg2(d2, make_smem_desc_sw128(sa + 128 * 128 + k * 32, 16, 1024), bdesc, idesc128, acc);
          }
          umma_commit_cg2_mc(smem_u32(&empty_bar[stage]), 3);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_cg2_mc(smem_u32(&tfull_bar[buf]), 3);
        if (it == 0) trace_stamp(p.trace, 3);
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = hw_warp & 3;         // TMEM lane quarter this (hardware) warp may access
    const int half = (warp - 2) >> 2;  // which of the two warps sharing that quarter
    float* stage = epi_stage + (warp - 2) * STAGE_F32_PER_WARP;
    const DropState dstate = drop_state(p.e.drop);
    uint32_t it = 0;
    for (int item = worker; item < p.num_items; item += nworkers, ++it) {
      const int m_pair = item / p.num_n_blocks;
      const int n_blk = item - m_pair * p.num_n_blocks;
      const int row_cta = (m_pair * 2 + (int)rank) * C::ROWS;
      const int col_tile = n_blk * BN;
      const uint32_t buf = C::DB ? (it & 1u) : 0u;
      const uint32_t use = C::DB ? (it >> 1) : it;
      if (p.epi_prefetch) {
        // A single wave leaves this epilogue fully exposed, and each of its chunks starts with a DRAM round trip for the
        // residual (fp32 LayerNorm input), the saved GELU' or the keep flags: fetch those lines into L2 while the tile's MMAs
        // run, so that the chunk loads below cost an L2 hit.
        using T = EpiTraits<EPI>;
        const int act = T::kStatic ? T::act : p.e.act;
        const int rk = T::kStatic ? T::resid : p.e.resid_kind;
        const bool aux_in = (act == ACT_DGELU_MUL || act == ACT_DRELU_MUL);
        const int et = (warp - 2) * 32 + lane;                       // epilogue thread index 0 .. 255
        const int rows = min(C::ROWS, p.M - row_cta);
        if (rows > 0 && (aux_in || rk != RESID_NONE)) {
          const char* src = reinterpret_cast<const char*>(aux_in ? p.e.aux : p.e.resid);
          const int esz = (!aux_in && (rk == RESID_F32 || rk == RESID_LN_F32)) ? 4 : 2;
          const size_t ld_bytes = (size_t)(aux_in ? p.e.ld_aux : p.e.ldr) * esz;
          const int lpr = (BN * esz + 127) >> 7;                     // 128-byte lines per row of the tile
          for (int l = et; l < rows * lpr; l += EPI_WARPS * 32) {
            const int r = l / lpr, sgm = l - r * lpr;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(row_cta + r) * ld_bytes + (size_t)col_tile * esz + (size_t)sgm * 128));
          }
        }
        if (rows > 0 && (T::kStatic ? T::drop : (p.e.drop.thresh != 0u))) {
          const int wpr = (p.N + 31) >> 5;
          for (int r = et; r < rows; r += EPI_WARPS * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.e.drop.bits + (size_t)(row_cta + r) * wpr + (col_tile >> 5)));
        }
        if (rows > 0 && rk == RESID_LN_F32 && p.e.ln_mean != nullptr && et < 16) {
          const float* st = (et < 8 ? p.e.ln_mean : p.e.ln_rstd) + row_cta + (et & 7) * 32;
          if (row_cta + (et & 7) * 32 < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(st));
        }
      }
      mbar_wait(smem_u32(&tfull_bar[buf]), use & 1u);
      tc_fence_after();
      if (it == 0 && warp == 2 && lane == 0) trace_stamp(p.trace, 4);
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      // ---- operand staging for the exposed epilogue of the CTA's last tile (see StagedChunk) ----
      // All MMAs have completed and the producer has nothing left to load, so the operand ring is idle: every residual vector,
      // keep-flag word and LayerNorm statistic this warp will need goes into its slice of it with cp.async, all in flight at once.
      using TE = EpiTraits<EPI>;
      constexpr int kSlice = 22528;   // 5 chunks x 4 KB residual + 5 x 32 keep words + 2 x 2 x 32 statistics, rounded up
      constexpr bool kCanStage = EpiFast<EPI>::value && TE::resid == RESID_LN_F32 && EPI_WARPS * kSlice <= C::STAGES * C::STAGE_BYTES &&
                                 (C::B_HALF / 32 + 1) / 2 + (BN / 32 + 1) / 2 <= 5;
      const bool staged = kCanStage && p.epi_stage && item + nworkers >= p.num_items && epi_fast_ok<EPI>(p.e);
      uint8_t* slice = smem + (warp - 2) * kSlice;
      float4* s_in = reinterpret_cast<float4*>(slice);
      uint32_t* s_kw = reinterpret_cast<uint32_t*>(slice + 5 * 4096);
      float* s_mu = reinterpret_cast<float*>(slice + 5 * 4096 + 5 * 128);
      float* s_rs = s_mu + 64;
      const int rows2 = row_cta + 128 + (q & 1) * 32;            // first row of this warp's part of acc2 / acc1
      const int rows1 = row_cta + q * 32;
      const int col2 = col_tile + (q >> 1) * C::B_HALF;
      if (staged) {
        const int g = lane & 7, r0 = lane >> 3;
        const float* rsrc = reinterpret_cast<const float*>(p.e.resid);
        const size_t wpr = (size_t)((p.N + 31) >> 5);
        int k = 0;
        auto issue = [&](int rb, int c0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = rb + r0 + 4 * i;
            const bool ok = row < p.M;
            cp_async_16(smem_u32(s_in + (k * 8 + i) * 32 + lane), rsrc + (size_t)(ok ? row : 0) * p.e.ldr + c0 + g * 4, ok ? 16u : 0u);
          }
          if (TE::drop) {
            const bool ok = rb + lane < p.M;
            cp_async_4(smem_u32(s_kw + k * 32 + lane), p.e.drop.bits + (size_t)(ok ? rb + lane : 0) * wpr + (size_t)(c0 >> 5), ok ? 4u : 0u);
          }
          ++k;
        };
        for (int c = half; c < C::B_HALF / 32; c += 2) issue(rows2, col2 + c * 32);
        for (int c = half; c < BN / 32; c += 2) issue(rows1, col_tile + c * 32);
        if (p.e.ln_mean != nullptr) {
          const bool ok2 = rows2 + lane < p.M, ok1 = rows1 + lane < p.M;
          cp_async_4(smem_u32(s_mu + lane), p.e.ln_mean + (ok2 ? rows2 + lane : 0), ok2 ? 4u : 0u);
          cp_async_4(smem_u32(s_rs + lane), p.e.ln_rstd + (ok2 ? rows2 + lane : 0), ok2 ? 4u : 0u);
          cp_async_4(smem_u32(s_mu + 32 + lane), p.e.ln_mean + (ok1 ? rows1 + lane : 0), ok1 ? 4u : 0u);
          cp_async_4(smem_u32(s_rs + 32 + lane), p.e.ln_rstd + (ok1 ? rows1 + lane : 0), ok1 ? 4u : 0u);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
      }
      int kst = 0;   // staged chunk counter (same order as the issue loops)
      auto release = [&](uint64_t* bar) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(smem_u32(bar));
          else mbar_arrive_cluster(mapa_cluster(smem_u32(bar), 0));
        }
      };
      // rows 128..191 first (the single acc2 buffer is what the next tile's MMAs wait for): lane quarters 0/1 hold the left
      // half of the columns, quarters 2/3 the right half
      {
        const int row_base = rows2;
        const int col_half = col2;
#pragma unroll 1
        for (int c = half; c < C::B_HALF / 32; c += 2, ++kst) {
          uint32_t v[32];
          tmem_ld32(t_lane + C::ACC2_COL + c * 32, v);
          tmem_ld_wait();
          StagedChunk sc;
          if (staged) { sc.in32 = s_in + kst * 256 + lane; sc.kw = s_kw + kst * 32; sc.mu = s_mu; sc.rs = s_rs; }
          if (epi_fast_ok<EPI>(p.e) && row_base + 32 <= p.M) epilogue_chunk_fast<EPI>(p.e, v, stage, lane, row_base, col_half + c * 32, p.N, sc);
          else if (row_base < p.M) epilogue_chunk<EPI>(p.e, v, stage, lane, row_base, col_half + c * 32, p.M, p.N, nullptr, 0, dstate);
        }
        release(tempty2_bar);
        if (it == 0 && warp == 2 && lane == 0) trace_stamp(p.trace, 5);
      }
      // rows 0..127 of this CTA
      {
        const int row_base = rows1;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2, ++kst) {
          uint32_t v[32];
          tmem_ld32(t_lane + buf * BN + c * 32, v);
          tmem_ld_wait();
          StagedChunk sc;
          if (staged) { sc.in32 = s_in + kst * 256 + lane; sc.kw = s_kw + kst * 32; sc.mu = s_mu + 32; sc.rs = s_rs + 32; }
          if (epi_fast_ok<EPI>(p.e) && row_base + 32 <= p.M) epilogue_chunk_fast<EPI>(p.e, v, stage, lane, row_base, col_tile + c * 32, p.N, sc);
          else if (row_base < p.M) epilogue_chunk<EPI>(p.e, v, stage, lane, row_base, col_tile + c * 32, p.M, p.N, nullptr, 0, dstate);
        }
        release(&tempty1_bar[buf]);
        if (it == 0 && warp == 2 && lane == 0) trace_stamp(p.trace, 6);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's shared memory / tensor memory must outlive everything that targets it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
  }
  if (threadIdx.x == 0) trace_stamp(p.trace, 7);
}
