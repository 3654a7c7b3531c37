// On-device sampling for `model.generate(do_sample=True, temperature=, top_p=)` — the call serve/cli.py:121-132 and
// serve/chat.py:141-151 make.  Restates HF GenerationMixin.sample():
//   scores = logits[:, -1] ; TemperatureLogitsWarper (scores / T) ; TopKLogitsWarper (k = 50, the GenerationConfig
//   default nobody on the path overrides; ties with the k-th score survive: `scores < topk[-1]` is what is removed) ;
//   TopPLogitsWarper (sort, cumulative softmax, drop while the ascending cumulative mass <= 1 - top_p, keep >= 1) ;
//   probs = softmax(scores) ; next = multinomial(probs, 1)
// (transformers/generation/logits_process.py Temperature/TopK/TopP warpers, generation/utils.py sample()).
//
// One CTA per sequence: the whole fp16 logit row (V <= 32768) is bitonic-sorted in shared memory as packed
// (orderable fp16 key << 16 | 0xFFFF - index) words, so top-k, top-p and the inverse-CDF draw all run on the SORTED
// row with deterministic block scans — no global scratch, no atomics, identical result for identical (seed, step, row).
// torch.multinomial's own RNG stream cannot be reproduced (and differs CPU vs CUDA in the reference itself); the draw is
// the inverse CDF of the filtered distribution at u = Philox4x32-10(seed; step, row), which the oracle restates bit for bit.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kSampleThreads = 1024;
constexpr int kSampleMaxV = 32768;

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
}

// Philox4x32-10 (Salmon et al. 2011), counter = (step, row, 0, 0), key = seed.  Returns the first output word.
__host__ __device__ inline uint32_t philox_first(uint64_t seed, uint32_t c0, uint32_t c1) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = 0u, x3 = 0u;
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return x0;
}

__device__ __forceinline__ uint32_t half_bits_to_key(uint32_t h) {
  if (h == 0x8000u) h = 0u;  // -0 == +0
  return (h & 0x8000u) ? (~h & 0xFFFFu) : (h | 0x8000u);
}
__device__ __forceinline__ float key_to_float(uint32_t key) {
  const uint32_t h = (key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu);
  return __half2float(__ushort_as_half(static_cast<unsigned short>(h)));
}

__device__ __forceinline__ int block_sum_int(int v, int* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  int t = (lane < kSampleThreads / 32) ? scratch[lane] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// exclusive prefix of v over the block in thread order; *total = block sum.  Fixed evaluation order (deterministic).
__device__ __forceinline__ float block_exclusive_scan(float v, float* scratch, float* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  float w = (lane < kSampleThreads / 32) ? scratch[lane] : 0.f;
  float winc = w;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  const float warp_excl = __shfl_sync(0xffffffffu, winc - w, warp);
  *total = __shfl_sync(0xffffffffu, winc, 31);
  return warp_excl + (inc - v);
}

struct SampleDeviceParams {  // lives in device memory so a captured decode graph follows vcb_set_sampling
  float inv_temperature;
  float top_p;
  int32_t top_k;
  uint32_t seed_lo, seed_hi;
};

// in-place bitonic sort of n (power of two) words, descending
__device__ __forceinline__ void bitonic_desc(uint32_t* a, int n) {
  const int tid = threadIdx.x;
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += kSampleThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i + j;
        const uint32_t x = a[i], y = a[p];
        const bool desc = (i & k) == 0;
        if (desc ? (x < y) : (x > y)) { a[i] = y; a[p] = x; }
      }
      __syncthreads();
    }
  }
}

constexpr int kCandMax = 2048;  // top-k fast path: survivors (k + ties with the k-th score) sorted on their own

__global__ void __launch_bounds__(kSampleThreads)
sample_kernel(const __half* __restrict__ logits, int V, int n_sort, const SampleDeviceParams* __restrict__ prm,
              const int32_t* __restrict__ step_dev, int step_host, int32_t* next, int32_t* trace, int trace_stride,
              int32_t* kept_out, int allow_select) {
  extern __shared__ uint32_t s_all[];  // [n_sort] packed row | [kCandMax] top-k survivors
  __shared__ float s_f[kSampleThreads / 32];
  __shared__ int s_i[kSampleThreads / 32];
  __shared__ int s_hist[256];
  __shared__ int s_sel[4];
  __shared__ float s_z2;
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  uint32_t* s_w = s_all;
  const unsigned short* row = reinterpret_cast<const unsigned short*>(logits) + static_cast<long>(b) * V;
  for (int i = tid; i < n_sort; i += kSampleThreads)
    s_w[i] = i < V ? ((half_bits_to_key(row[i]) << 16) | (0xFFFFu - static_cast<uint32_t>(i))) : 0u;
  const float inv_t = prm->inv_temperature, top_p = prm->top_p;
  const int top_k = prm->top_k;
  __syncthreads();

  int n_k = V;
  bool sorted = false;
  if (allow_select && top_k > 0 && top_k <= 1024 && top_k < V) {
    // ---- top-k by two-level radix select on the 16-bit key (integer histograms: exact and deterministic), then only the
    //      survivors are sorted.  TopKLogitsWarper keeps every score >= the k-th largest, ties included.
    for (int level = 0; level < 2; ++level) {
      for (int i = tid; i < 256; i += kSampleThreads) s_hist[i] = 0;
      __syncthreads();
      const uint32_t hb = static_cast<uint32_t>(s_sel[0]);
      for (int i = tid; i < V; i += kSampleThreads) {
        const uint32_t key = s_w[i] >> 16;
        if (level == 0) atomicAdd(&s_hist[key >> 8], 1);
        else if ((key >> 8) == hb) atomicAdd(&s_hist[key & 0xFFu], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int need = level == 0 ? top_k : s_sel[1];  // rank (1-based, from the top) still to find inside this level
        int bin = 255;
        for (; bin > 0; --bin) {
          if (s_hist[bin] >= need) break;
          need -= s_hist[bin];
        }
        if (level == 0) { s_sel[0] = bin; s_sel[1] = need; }
        else s_sel[2] = (s_sel[0] << 8) | bin;
      }
      __syncthreads();
    }
    const uint32_t kth = static_cast<uint32_t>(s_sel[2]);
    int c = 0;
    for (int i = tid; i < V; i += kSampleThreads) c += ((s_w[i] >> 16) >= kth) ? 1 : 0;
    n_k = block_sum_int(c, s_i);
    if (n_k <= kCandMax) {
      uint32_t* cand = s_all + n_sort;
      int n2 = 32;
      while (n2 < n_k) n2 <<= 1;
      if (tid == 0) s_sel[3] = 0;
      for (int i = tid; i < n2; i += kSampleThreads) cand[i] = 0u;
      __syncthreads();
      for (int i = tid; i < V; i += kSampleThreads) {
        const uint32_t wv = s_w[i];
        if ((wv >> 16) >= kth) cand[atomicAdd(&s_sel[3], 1)] = wv;  // slot order is arbitrary; the sort below fixes it
      }
      __syncthreads();
      bitonic_desc(cand, n2);
      s_w = cand;
      n_sort = n2;
      sorted = true;
    }
  }
  if (!sorted) {
    // ---- full bitonic sort, descending: score desc, index asc among equal scores
    bitonic_desc(s_w, n_sort);
    // ---- top-k with ties (TopKLogitsWarper removes scores < the k-th largest)
    n_k = V;
    if (top_k > 0 && top_k < V) {
      const uint32_t kth = s_w[top_k - 1] >> 16;
      int c = 0;
      for (int i = tid; i < n_sort; i += kSampleThreads) c += (i < V && (s_w[i] >> 16) >= kth) ? 1 : 0;
      n_k = block_sum_int(c, s_i);
    }
  }
  // ---- softmax numerators on the sorted row; every thread owns a contiguous chunk so the scan runs in rank order
  const int ept = n_sort / kSampleThreads > 0 ? n_sort / kSampleThreads : 1;
  const int lo = tid * ept;
  const float top = key_to_float(s_w[0] >> 16);
  float local = 0.f;
  for (int e = 0; e < ept; ++e) {
    const int i = lo + e;
    if (i < n_k && i < n_sort) local += expf((key_to_float(s_w[i] >> 16) - top) * inv_t);
  }
  float z1;
  const float excl = block_exclusive_scan(local, s_f, &z1);
  // ---- top-p: rank i survives iff the mass of the ranks before it is < top_p (== HF's ascending cumsum > 1 - top_p)
  const float cut = top_p * z1;
  const bool use_p = top_p < 1.f;
  int c = 0;
  {
    float run = excl;
    for (int e = 0; e < ept; ++e) {
      const int i = lo + e;
      if (i < n_k && i < n_sort) {
        if (!use_p || run < cut) ++c;
        run += expf((key_to_float(s_w[i] >> 16) - top) * inv_t);
      }
    }
  }
  int n_p = block_sum_int(c, s_i);
  if (n_p < 1) n_p = 1;  // min_tokens_to_keep
  // ---- Z2 = mass of the survivors
  {
    float run = excl;
    for (int e = 0; e < ept; ++e) {
      const int i = lo + e;
      if (i < n_k && i < n_sort) {
        run += expf((key_to_float(s_w[i] >> 16) - top) * inv_t);
        if (i == n_p - 1) s_z2 = run;
      }
    }
  }
  __syncthreads();
  const int step = step_dev ? *step_dev : step_host;
  const uint64_t seed = (static_cast<uint64_t>(prm->seed_hi) << 32) | prm->seed_lo;
  const float u = static_cast<float>(philox_first(seed, static_cast<uint32_t>(step), static_cast<uint32_t>(b)) >> 8) *
                  (1.0f / 16777216.0f);
  const float target = u * s_z2;
  // ---- inverse CDF: first rank whose inclusive mass exceeds target
  c = 0;
  {
    float run = excl;
    for (int e = 0; e < ept; ++e) {
      const int i = lo + e;
      if (i < n_p && i < n_sort) {
        run += expf((key_to_float(s_w[i] >> 16) - top) * inv_t);
        if (run <= target) ++c;
      }
    }
  }
  int pick = block_sum_int(c, s_i);
  if (pick > n_p - 1) pick = n_p - 1;
  if (tid == 0) {
    const int tok = static_cast<int>(0xFFFFu - (s_w[pick] & 0xFFFFu));
    next[b] = tok;
    if (trace) trace[static_cast<long>(b) * trace_stride + step] = tok;
    if (kept_out) kept_out[b] = n_p;
  }
}

}  // namespace

int sample_rows(const __half* logits, int B, int V, const void* params_dev, const int32_t* step_dev, int step_host,
                int32_t* next, int32_t* trace, int trace_stride, int32_t* kept_out, cudaStream_t s) {
  VCB_CHECK(V >= 2 && V <= kSampleMaxV, kErrInvalid, "sampling: vocabulary must be in [2, 32768]");
  int n_sort = kSampleThreads;  // at least one element per thread keeps the chunked scans simple
  while (n_sort < V) n_sort <<= 1;
  const size_t smem = static_cast<size_t>(n_sort + kCandMax) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    VCB_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (kSampleMaxV + kCandMax) * 4));
    attr_set = true;
  }
  static int full_env = -1;  // VCB_SAMPLE_FULLSORT=1: always sort the whole row (A/B against the top-k select path)
  if (full_env < 0) { const char* ev = getenv("VCB_SAMPLE_FULLSORT"); full_env = ev ? atoi(ev) : 0; }
  KernelScope scope(kCatMisc, s);
  VCB_CUDA(launch_pdl(sample_kernel, dim3(B), dim3(kSampleThreads), smem, s, logits, V, n_sort,
                      static_cast<const SampleDeviceParams*>(params_dev), step_dev, step_host, next, trace,
                      trace_stride, kept_out, full_env ? 0 : 1));
  return kOk;
}

// host packer for the device parameter block (layout = SampleDeviceParams)
void pack_sample_params(float temperature, int top_k, float top_p, uint64_t seed, void* out20) {
  SampleDeviceParams p;
  p.inv_temperature = 1.0f / temperature;
  p.top_p = top_p;
  p.top_k = top_k;
  p.seed_lo = static_cast<uint32_t>(seed);
  p.seed_hi = static_cast<uint32_t>(seed >> 32);
  memcpy(out20, &p, sizeof(p));
}

uint32_t philox_first_host(uint64_t seed, uint32_t step, uint32_t row) { return philox_first(seed, step, row); }

}  // namespace vcb
