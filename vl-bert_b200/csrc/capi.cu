// extern "C" entry points of libvlbert_b200.so (declared in include/vlbert_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/vlbert_b200.h"
#include "common.cuh"
#include "gemm_sm100.cuh"

namespace vlb {

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_last_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return VLB_ERR_CUDA;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void gemm_debug_override(uint32_t mn_lbo, uint32_t mn_sbo, uint32_t mn_kadv);

}  // namespace vlb

using namespace vlb;

extern "C" {

int vlb_abi_version(void) { return 1; }
const char* vlb_last_error_string(void) { return g_err; }
int64_t vlb_launch_count(void) { return g_launches.load(); }

int vlb_gemm_bf16(int mode, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* out,
                  int ldo, int out_kind, const float* bias, const void* resid, int ldr, int resid_kind, int act,
                  void* aux, int ld_aux, float alpha, int split_k, int force_bn, void* stream) {
  GemmEpilogue e;
  e.out = out; e.ldo = ldo; e.out_kind = out_kind;
  e.bias = bias;
  e.resid = resid; e.ldr = ldr; e.resid_kind = resid_kind;
  e.act = act; e.aux = aux; e.ld_aux = ld_aux; e.alpha = alpha;
  int rc = gemm_bf16(mode, M, N, K, A, lda, B, ldb, e, split_k, force_bn, static_cast<cudaStream_t>(stream));
  if (rc == VLB_OK) count_launch(1);
  return rc;
}

void vlb_debug_gemm_desc(uint32_t mn_lbo, uint32_t mn_sbo, uint32_t mn_kadv) {
  gemm_debug_override(mn_lbo, mn_sbo, mn_kadv);
}

}  // extern "C"
