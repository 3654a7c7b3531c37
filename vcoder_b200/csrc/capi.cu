// extern "C" surface of libvcoder_b200 (declarations + reference citations: include/vcoder_b200.h).
#include "vcoder_b200.h"

#include "common.h"

namespace vcb {
static thread_local std::string g_err;
void set_thread_error(const std::string& m) { g_err = m; }
const char* get_thread_error() { return g_err.c_str(); }
}  // namespace vcb

using namespace vcb;

static int require_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_thread_error("no CUDA device visible: libvcoder_b200 has no CPU fallback");
    return kErrNoDevice;
  }
  return kOk;
}

extern "C" {

const char* vcb_version(void) { return "vcoder_b200 0.1.0 (sm_100a)"; }
const char* vcb_last_error(void) { return get_thread_error(); }

int vcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vcb_op_gemm_f16(const vcb_half* A, int32_t lda, const vcb_half* W, int32_t ldw, int32_t M, int32_t N, int32_t K,
                    const vcb_half* bias, const vcb_half* residual, int32_t ldr, int32_t act, vcb_half* C,
                    int32_t ldc, int32_t block_n, void* stream) {
  VCB_TRY(require_device());
  GemmEpilogue ep;
  ep.C = reinterpret_cast<__half*>(C);
  ep.ldc = ldc;
  ep.bias = reinterpret_cast<const __half*>(bias);
  ep.residual = reinterpret_cast<const __half*>(residual);
  ep.ldr = ldr;
  ep.act = act;
  return gemm_f16(reinterpret_cast<const __half*>(A), lda, reinterpret_cast<const __half*>(W), ldw, M, N, K, ep,
                  block_n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
