// Host-side interface of the tcgen05 GEMM (see gemm_sm100.cu).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace vlb {

enum GemmOutKind : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ATOMIC = 2 };
// RESID_LN_F32: the residual is the fp32 OUTPUT of a LayerNorm that is not stored: it is recomputed in the epilogue from the
// LayerNorm's stored fp32 input `resid` [M, ldr], its row statistics ln_mean / ln_rstd [M] and ln_gamma / ln_beta [N]
// (ln_mean == nullptr: `resid` is used as it is, i.e. RESID_F32).  This keeps the encoder's residual stream in fp32 --
// like the reference under autocast -- without writing a second copy of every LayerNorm output.
enum GemmResidKind : int { RESID_NONE = 0, RESID_BF16 = 1, RESID_F32 = 2, RESID_LN_F32 = 3 };
enum GemmAct : int {
  ACT_NONE = 0,
  ACT_GELU = 1,       // x = gelu_erf(x); if aux != null gelu_erf'(pre-activation) is stored there (bf16) for backward
  ACT_RELU = 2,
  ACT_DGELU_MUL = 3,  // x = x * aux[row, col]                 (aux: the bf16 gelu' saved by ACT_GELU)
  ACT_DRELU_MUL = 4,  // x = aux[row, col] > 0 ? x : 0         (aux: bf16 post-activation)
  ACT_RELU_POST = 5,  // ReLU applied AFTER the residual add   (ResNet bottleneck output: relu(bn3(conv3) + identity))
};
// Operand layout modes.  "K-major" = reduction dimension contiguous in memory.
//   GEMM_NT : C[M,N] = A[M,K] * B[N,K]^T      (forward  y = x W^T ; A, B K-major)
//   GEMM_NN : C[M,N] = A[M,K] * B[K,N]        (dgrad    dx = dy W ; B is N-contiguous = MN-major)
//   GEMM_TN : C[M,N] = A[K,M]^T * B[K,N]      (wgrad    dW = dy^T x ; both MN-major)
enum GemmMode : int { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2 };

struct GemmEpilogue {
  void* out = nullptr;
  int ldo = 0;
  int out_kind = OUT_BF16;
  const float* bias = nullptr;   // [N] fp32, optional
  const void* resid = nullptr;   // [M, ldr], optional
  int ldr = 0;
  int resid_kind = RESID_NONE;
  int act = ACT_NONE;
  void* aux = nullptr;           // bf16 [M, ld_aux]
  int ld_aux = 0;
  float alpha = 1.0f;            // scale applied to the accumulator before everything else
  float* colsum = nullptr;       // optional [N] fp32: += column sums of the stored values (bias gradient of the producer)
  const float* colscale = nullptr;  // optional [N] fp32: x = acc * colscale[col] before the bias (frozen BatchNorm scale)
  DropCfg drop = DropCfg{0u, 1.0f, 0u, nullptr, nullptr};  // dropout on the [M,N] value after bias/activation, before the residual add (needs drop.bits)
  const float* ln_mean = nullptr;   // RESID_LN_F32: see GemmResidKind
  const float* ln_rstd = nullptr;
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
};

// timing aid: 8 globaltimer stamps per CTA of the pair kernel are written to buf (device memory, >= grid * 8 entries); nullptr = off
void gemm_debug_trace(unsigned long long* buf);

// whether the experimental stream-K tail was compiled in (-DVLB_ENABLE_STREAMK=1)
bool gemm_streamk_compiled();

// geometry of one convolution on an NHWC tensor (shared by the lowering kernels in conv.cu and the implicit-GEMM operand)
struct ConvGeom {
  int N, H, W, C;      // input NHWC
  int Ho, Wo;          // output spatial
  int kh, kw, stride, pad, dil;
  int Kp;              // padded row length of an explicit col matrix (>= kh*kw*C, multiple of 8); unused by the implicit path
};

// All matrices are bf16 row-major with leading dimensions in elements (multiples of 8).
// conv / conv_side: implicit convolution operand gathered by TMA im2col-mode loads from the NHWC tensor passed as that operand's
// pointer (its leading dimension is ignored): conv_side 1 = A of an NT GEMM (M = N*Ho*Wo output pixels, K = kh*kw*C),
// 2 = B of a TN GEMM (K = output pixels, N = kh*kw*C).  Needs C % 64 == 0.
int gemm_bf16(int mode, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
              const GemmEpilogue& epi, int split_k, int force_bn, cudaStream_t stream, const ConvGeom* conv = nullptr,
              int conv_side = 0);

// Grouped weight-gradient launch: out_i[M_i, N_i] (+)= A_i[K, M_i]^T B_i[K, N_i] for up to 4 problems sharing K, in ONE
// persistent grid (equal-cost tiles of all problems are interleaved -> full rounds).  accumulate: fp32 atomic "+="
// (required when split_k > 1), else plain fp32 stores.
struct GroupedProblem {
  int M, N;
  const void* A; int lda;
  const void* B; int ldb;
  float* out; int ldo;
};
int gemm_grouped_tn(int count, const GroupedProblem* probs, int K, int split_k, bool accumulate, int bn, cudaStream_t stream);

// TMA descriptor helper shared with the attention kernels: 2D bf16 row-major [rows, cols] with
// 128B swizzle; box = {box_cols (<=64), box_rows (<=256)}.
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows);

}  // namespace vlb
