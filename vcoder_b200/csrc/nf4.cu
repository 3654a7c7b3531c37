// bitsandbytes NF4 (+ double quantisation) as a load-time weight transform: W -> dequantize(quantize(W)) in fp16.
// Replaces what `load_pretrained_model(load_4bit=True)` sets up in the reference (vcoder_llava/model/builder.py:28-37:
// BitsAndBytesConfig(nf4, double quant, compute dtype fp16)): every nn.Linear of the language model and the projectors
// except lm_head becomes bnb.nn.Linear4bit, whose forward de-quantises the weight to fp16 and runs an fp16 matmul.  The
// engine keeps the de-quantised fp16 weights (same arithmetic as bnb's de-quantise + matmul path; the 4-bit storage and
// the 4x smaller weight stream are NOT implemented), so `load_4bit` reproduces the reference's numerics, not its memory.
// Algorithm: bitsandbytes 0.41 functional.quantize_4bit / csrc dQuantizeNF4 / dDequantizeNF4 (published: QLoRA, 2023);
// oracle: oracle/bnb_nf4_oracle.py.  PARITY UNPINNED (bitsandbytes is not available offline): see DESIGN.md section 4.
#include <cmath>
#include <vector>

#include "common.h"
#include "vcoder_b200.h"

namespace vcb {
namespace {

__constant__ float c_nf4[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                                0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};
__constant__ float c_thr[15];    // midpoints of adjacent levels (dQuantizeNF4's decision thresholds)
__constant__ float c_code8[256]; // the 8-bit dynamic map (create_dynamic_map), sorted

// one thread per 64-weight block: absmax (fp32)
__global__ void nf4_absmax_kernel(const __half* __restrict__ w, long n_blocks, float* __restrict__ absmax) {
  const long b = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const uint4* p = reinterpret_cast<const uint4*>(w + b * 64);
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 v = p[i];
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  absmax[b] = m;
}

// deterministic mean in double: 256 CTAs write fixed-order partial sums, one thread adds them
__global__ void nf4_sum_kernel(const float* __restrict__ a, long n, double* __restrict__ parts) {
  __shared__ double s[256];
  double acc = 0.0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) acc += a[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) parts[blockIdx.x] = s[0];
}
__global__ void nf4_mean_kernel(const double* __restrict__ parts, int n_parts, long n, float* __restrict__ offset) {
  double t = 0.0;
  for (int i = 0; i < n_parts; ++i) t += parts[i];
  *offset = static_cast<float>(t / static_cast<double>(n));
}

// one thread per 256-group of (absmax - offset): nearest 8-bit code, de-quantised back in place
__global__ void nf4_double_quant_kernel(float* __restrict__ absmax, long n, const float* __restrict__ offset_p) {
  const long g = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long i0 = g * 256;
  if (i0 >= n) return;
  const long i1 = min(n, i0 + 256);
  const float offset = *offset_p;
  float m2 = 0.f;
  for (long i = i0; i < i1; ++i) m2 = fmaxf(m2, fabsf(absmax[i] - offset));
  const float inv = 1.0f / (m2 > 0.f ? m2 : 1.0f);
  for (long i = i0; i < i1; ++i) {
    const float x = (absmax[i] - offset) * inv;
    int lo = 0, hi = 255;  // first code value >= x
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (c_code8[mid] < x) lo = mid + 1; else hi = mid;
    }
    int q = lo;
    if (q > 0 && (static_cast<double>(c_code8[q]) - x) >= (static_cast<double>(x) - c_code8[q - 1])) q = q - 1;  // ties: lower index
    absmax[i] = __fadd_rn(__fmul_rn(c_code8[q], m2), offset);  // two roundings, as the reference's separate multiply and add
  }
}

// one thread per 64-weight block: nearest NF4 level of w / absmax_raw, de-quantised with the (double-quantised) absmax
__global__ void nf4_roundtrip_kernel(const __half* __restrict__ w, long n_blocks, const float* __restrict__ absmax_raw,
                                     const float* __restrict__ absmax_deq, __half* __restrict__ out, uint8_t* __restrict__ idx_out) {
  const long b = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const float a = absmax_raw[b];
  const float inv = 1.0f / (a > 0.f ? a : 1.0f);
  const float d = absmax_deq[b];
  for (int i = 0; i < 64; ++i) {
    const float x = __half2float(w[b * 64 + i]) * inv;
    int q = 0;
#pragma unroll
    for (int t = 0; t < 15; ++t) q += (x > c_thr[t]) ? 1 : 0;
    out[b * 64 + i] = __float2half_rn(c_nf4[q] * d);
    if (idx_out) idx_out[b * 64 + i] = static_cast<uint8_t>(q);
  }
}

}  // namespace
}  // namespace vcb

extern "C" int vcb_op_nf4_roundtrip_f16(const vcb_half* w_dev, int64_t n, int32_t double_quant, const float* code8_host,
                                        vcb_half* out_dev, uint8_t* idx_out_dev, void* stream) {
  using namespace vcb;
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd <= 0) {
    cudaGetLastError();
    set_thread_error("no CUDA device visible: libvcoder_b200 has no CPU fallback");
    return kErrNoDevice;
  }
  VCB_CHECK(w_dev != nullptr && out_dev != nullptr && n > 0 && n % 64 == 0, kErrInvalid, "vcb_op_nf4_roundtrip_f16: n must be a positive multiple of 64");
  VCB_CHECK((reinterpret_cast<uintptr_t>(w_dev) & 15) == 0, kErrInvalid, "vcb_op_nf4_roundtrip_f16: input must be 16-byte aligned");
  VCB_CHECK(!double_quant || code8_host != nullptr, kErrInvalid, "vcb_op_nf4_roundtrip_f16: double quantisation needs the 256-entry code table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool thr_set = false;
  if (!thr_set) {
    static const double lv[16] = {-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                                  -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                                  0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
                                  0.7229568362236023, 1.0};
    float thr[15];
    for (int i = 0; i < 15; ++i) {
      // the levels are float32 constants; their midpoint is taken in double and rounded once, as the oracle does
      thr[i] = static_cast<float>((static_cast<double>(static_cast<float>(lv[i])) + static_cast<double>(static_cast<float>(lv[i + 1]))) / 2.0);
    }
    VCB_CUDA(cudaMemcpyToSymbol(c_thr, thr, sizeof(thr)));
    thr_set = true;
  }
  if (double_quant) VCB_CUDA(cudaMemcpyToSymbolAsync(c_code8, code8_host, 256 * sizeof(float), 0, cudaMemcpyHostToDevice, st));
  const long n_blocks = n / 64;
  float* absmax = nullptr;
  const size_t bytes = static_cast<size_t>(n_blocks) * 4;
  VCB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&absmax), 2 * bytes + 256 * 8 + 64, st));
  float* absmax_deq = absmax + n_blocks;
  double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(absmax) + ((2 * bytes + 7) & ~static_cast<size_t>(7)));
  float* offset = reinterpret_cast<float*>(parts + 256);
  const unsigned g64 = static_cast<unsigned>((n_blocks + 127) / 128);
  nf4_absmax_kernel<<<g64, 128, 0, st>>>(reinterpret_cast<const __half*>(w_dev), n_blocks, absmax);
  cudaMemcpyAsync(absmax_deq, absmax, bytes, cudaMemcpyDeviceToDevice, st);
  if (double_quant) {
    nf4_sum_kernel<<<256, 256, 0, st>>>(absmax, n_blocks, parts);
    nf4_mean_kernel<<<1, 1, 0, st>>>(parts, 256, n_blocks, offset);
    const long groups = (n_blocks + 255) / 256;
    nf4_double_quant_kernel<<<static_cast<unsigned>((groups + 63) / 64), 64, 0, st>>>(absmax_deq, n_blocks, offset);
  }
  nf4_roundtrip_kernel<<<g64, 128, 0, st>>>(reinterpret_cast<const __half*>(w_dev), n_blocks, absmax, absmax_deq,
                                            reinterpret_cast<__half*>(out_dev), idx_out_dev);
  const cudaError_t le = cudaGetLastError();
  cudaFreeAsync(absmax, st);
  VCB_CUDA(le);
  VCB_CUDA(cudaStreamSynchronize(st));  // code8_host may be a temporary of the caller
  return kOk;
}
