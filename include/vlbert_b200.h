/*
 * vlbert_b200 -- C ABI of the B200-native VL-BERT hot path (libvlbert_b200.so).
 *
 * Drop-in boundary for the path BASELINE.json's north_star names: the visual-linguistic transformer
 * encoder (reference: common/visual_linguistic_bert.py, external/pytorch_pretrained_bert/modeling.py)
 * and the region-feature front end (reference: common/fast_rcnn.py, common/lib/roi_pooling).
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers, sizes and leading dimensions in ELEMENTS; no torch types.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Work is only
 *     ENQUEUED on that stream; nothing synchronises and nothing is allocated behind the caller's
 *     back -- workspaces are passed in (the reference's native op enqueues on the current stream,
 *     common/lib/roi_pooling/cuda/ROIAlign_cuda.cu:273).
 *   - return value: 0 on success, negative on error (VLB_ERR_*); never throws.  The message of the
 *     last error on the calling thread is available from vlb_last_error_string() (the reference
 *     raises RuntimeError through AT_ASSERTM / THCudaCheck, ROIAlign_cuda.cu:262-264,297; the
 *     Python binding turns a negative return into the same RuntimeError).
 *   - "bf16" = __nv_bfloat16 storage, row-major; "f32" = float.
 *   - there is NO CPU fallback: every entry point needs an sm_100a device.
 */
#ifndef VLBERT_B200_H_
#define VLBERT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLB_OK 0
#define VLB_ERR_INVALID (-1)
#define VLB_ERR_CUDA (-2)
#define VLB_ERR_UNSUPPORTED (-3)

/* ---- library ------------------------------------------------------------------------------- */
int vlb_abi_version(void);
const char* vlb_last_error_string(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t vlb_launch_count(void);

/* ---- GEMM (tcgen05) --------------------------------------------------------------------------
 * Replaces the torch.nn.Linear / F.linear calls of the encoder layer and their autograd backward
 * (modeling.py:291-293, :330, :362, :375; common/fast_rcnn.py:105-109 obj_downsample).
 *   mode 0 (NT): C[M,N] = A[M,K] B[N,K]^T   forward   y = x W^T
 *   mode 1 (NN): C[M,N] = A[M,K] B[K,N]     dgrad     dx = dy W
 *   mode 2 (TN): C[M,N] = A[K,M]^T B[K,N]   wgrad     dW = dy^T x
 * Epilogue: x = alpha*acc; x += bias[col]; activation; x += resid[row,col]; store.
 *   out_kind   0 bf16, 1 f32, 2 f32 atomic accumulate (required when split_k > 1)
 *   resid_kind 0 none, 1 bf16, 2 f32
 *   act        0 none, 1 erf-GELU (pre-activation stored to aux if non-null), 2 ReLU,
 *              3 multiply by GELU'(aux), 4 multiply by [aux > 0]
 */
int vlb_gemm_bf16(int mode, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                  void* out, int ldo, int out_kind, const float* bias, const void* resid, int ldr,
                  int resid_kind, int act, void* aux, int ld_aux, float alpha, int split_k,
                  int force_bn, void* stream);

/* bring-up aid: override the MN-major shared-memory descriptor geometry (0 = default). */
void vlb_debug_gemm_desc(uint32_t mn_lbo, uint32_t mn_sbo, uint32_t mn_kadv);

#ifdef __cplusplus
}
#endif
#endif /* VLBERT_B200_H_ */
