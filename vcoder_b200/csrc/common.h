// Shared host-side declarations for libvcoder_b200: error codes, launch-check macros and the
// internal (C++) kernel launcher signatures. The public C ABI lives in include/vcoder_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace vcb {

// Error codes mirrored in include/vcoder_b200.h
enum : int {
  kOk = 0,
  kErrInvalid = -1,       // bad argument / shape
  kErrCuda = -2,          // CUDA runtime / driver failure (message has details)
  kErrState = -3,         // call order (e.g. decode before prefill)
  kErrNoMem = -4,         // KV pages / workspace exhausted
  kErrRagged = -5,        // batch whose spliced lengths differ (the reference raises NameError there)
  kErrSpecialIndex = -6,  // negative special id would reach the embedding table (reference: IndexError)
  kErrNoDevice = -7,      // no CUDA device: there is deliberately NO CPU fallback
};

void set_thread_error(const std::string& msg);
const char* get_thread_error();

#define VCB_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      ::vcb::set_thread_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return ::vcb::kErrCuda;                                                                            \
    }                                                                                                    \
  } while (0)

#define VCB_CHECK(cond, code, msg)                                                                       \
  do {                                                                                                   \
    if (!(cond)) {                                                                                       \
      ::vcb::set_thread_error(std::string(msg) + " [" #cond "]");                                        \
      return (code);                                                                                     \
    }                                                                                                    \
  } while (0)

#define VCB_TRY(expr)                                                                                    \
  do {                                                                                                   \
    int _r = (expr);                                                                                     \
    if (_r != 0) return _r;                                                                              \
  } while (0)

// -------------------------------------------------------------------------------------------------
// GEMM (tcgen05): C[M,N] = epilogue(A[M,K] * W[N,K]^T)
// -------------------------------------------------------------------------------------------------
enum GemmAct : int { kActNone = 0, kActQuickGelu = 1, kActGeluErf = 2, kActSwiGLU = 3, kActRopeKv = 4 };

struct GemmEpilogue {
  __half* C = nullptr;            // output, row-major, ldc elements per row
  int ldc = 0;
  const __half* bias = nullptr;   // [N] or null
  const __half* residual = nullptr;  // row-major, ldr per row, or null (may alias C)
  int ldr = 0;
  int act = kActNone;             // kActSwiGLU: W rows interleaved (8 gate, 8 up); output has N/2 columns
  // Output-row remap (patch-embed writes patch p of image b to row b*(group+skip)+off+p):
  //   out_row = m + (m / out_group) * out_skip + out_off        (out_group == 0: identity)
  int out_group = 0, out_skip = 0, out_off = 0;
  // Residual-row remap (position embedding is indexed by patch, not by image):
  //   res_row = res_mod ? (m % res_mod) + res_off : out_row
  int res_mod = 0, res_off = 0;
  // kActRopeKv (LLaMA QKV projection at prefill, N = 3 * rk_H): the epilogue rounds q | k | v to fp16, applies rotate-half
  // RoPE to q and k (position of row m = rk_pos0 + m % rk_T), writes q to C and k / v straight into the paged KV cache
  // [page][head][rk_page_tokens][128] of sequence m / rk_T (replaces the separate rope_kv_kernel pass: one read and one
  // write of the whole QKV buffer per layer).  No bias / residual / row remap in this mode.
  const float2* rk_cos_sin = nullptr;  // [position][64] (cos, sin)
  __half* rk_k_pool = nullptr;
  __half* rk_v_pool = nullptr;
  const int32_t* rk_block_table = nullptr;  // [sequence][rk_max_pages]
  int rk_T = 0, rk_pos0 = 0, rk_max_pages = 0, rk_heads = 0, rk_H = 0, rk_page_tokens = 0;
};

// A: [M,K] fp16 row-major (lda elements), W: [N,K] fp16 row-major (ldw elements). K % 64 == 0,
// lda/ldw % 8 == 0, N % 16 == 0. block_n: 0 = auto, else 128 or 256.
int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
             int block_n, cudaStream_t stream);

int device_sm_count();

// Launch with the programmatic-stream-serialization (PDL) attribute; the kernel MUST call griddep_wait().
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// -------------------------------------------------------------------------------------------------
// launch accounting / profiling (profile.cu)
// -------------------------------------------------------------------------------------------------
enum KernelCat : int {
  kCatGemm = 0, kCatSkinny, kCatAttnVit, kCatAttnPrefill, kCatAttnDecode, kCatNorm, kCatRope, kCatMisc, kNumCat
};
struct KernelScope {  // RAII around one launcher call: counts kernels, optionally brackets them with events
  KernelScope(int cat, cudaStream_t st, int n_kernels = 1);
  ~KernelScope();
  int cat_; cudaStream_t st_; int rec_;
};
// Device-side timeline (tools/trace_decode.py): while a trace is open, launchers that support it hand every CTA three
// 64-bit slots {enter, after griddepcontrol.wait, end} stamped with %globaltimer.  Works under graph capture (the slots are
// baked into the captured parameters; a replay overwrites them), so the production decode graph can be inspected.
enum TraceTag : int { kTraceSkinny = 0, kTraceAttnSplit = 1, kTraceAttnCombine = 2, kTraceRmsnorm = 3, kTraceOther = 4 };
unsigned long long* trace_slots(int tag, int n_ctas, int aux);  // null when no trace is open
int trace_begin(int max_ctas);
int trace_epoch();  // changes whenever a trace is opened or closed (captured graphs hold the slot pointers)
// launches_out: [max_launches][4] = {tag, n_ctas, base slot, aux}; stamps_out: [cap] u64 (3 per CTA).  Closes the trace.
int trace_read(int32_t* launches_out, int max_launches, int32_t* n_launches, unsigned long long* stamps_out, long cap);
uint64_t launch_count();
void launch_count_add(int64_t n);
void prof_enable(bool on);
bool prof_enabled();
int prof_read(float* ms_by_cat, uint64_t* n_by_cat);

}  // namespace vcb
