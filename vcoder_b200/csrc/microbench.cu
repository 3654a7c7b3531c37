// Hardware micro-benchmarks used to size the attention kernels (evidence for DESIGN.md section 3.2; tools/microbench.py).
// Not on the product path: nothing in the engine calls these.
#include <vector>

#include "common.h"
#include "ptx.cuh"
#include "vcoder_b200.h"

namespace vcb {
namespace {

// TMEM -> register read rate: every warp of the CTA loops over tcgen05.ld.32x32b.x32 (4 KiB per instruction) on its own
// lane quadrant, ILP loads in flight before each tcgen05.wait::ld.
template <int ILP>
__global__ void tmem_read_bench_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[ILP][32];
#pragma unroll
    for (int i = 0; i < ILP; ++i) tmem_ld_32x32(addr + static_cast<uint32_t>((((it * ILP + i) & 7) * 32 + (warp >> 2) * 256) & 511), v[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= v[i][0] ^ v[i][31];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// MUFU ex2 rate next to FMA-pipe work: every thread runs `iters` rounds of 8 independent ex2.approx (+ 8 FFMA)
__global__ void ex2_bench_kernel(int iters, int with_fma, unsigned long long* cycles, float* sink) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = fast_exp2(x[i]);
      if (with_fma) x[i] = fmaf(x[i], -0.5f, -0.25f);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) sink[0] = s;
}

// packed-half MUFU: ex2.approx.ftz.f16x2 — two exponentials per instruction; is it twice the rate?
__global__ void ex2_h2_bench_kernel(int iters, unsigned long long* cycles, float* sink) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0xB000B400u + threadIdx.x + i;  // small negative halves
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  if (s == 0x12345678u) sink[0] = 1.f;
}

// The exponential / pack / store loop of the ViT attention softmax (attn_vit_pair.cu) in isolation: every thread owns one
// 64-score half row in registers and writes its 128-byte P row (swizzled 16-byte chunks) per round.  variant bit 0: no row
// sums; bit 1: no shared-memory stores; bit 2: sums / packs placed 16 exponentials behind their MUFU (software skew).
template <int kVariant>
__global__ void softmax_loop_bench_kernel(int iters, unsigned long long* cycles, float* sink) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint32_t v[2][32];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) v[c][i] = __float_as_uint(-0.01f * static_cast<float>((threadIdx.x * 7 + c * 32 + i) & 255));
  const int r = threadIdx.x & 127;
  uint8_t* prow = sm + (threadIdx.x >> 7) * 16384 + r * 128;
  float L = 0.f;
  const float sc = 0.18f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float msc = 0.001f * static_cast<float>(it & 15);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[16];
      if (kVariant & 4) {
        constexpr int kSkew = 16;
        float e[32];
#pragma unroll
        for (int i = 0; i < 32 + kSkew; i += 2) {
          if (i < 32) {
            e[i] = fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc));
            e[i + 1] = fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc));
          }
          if (i >= kSkew) {
            const int j = i - kSkew;
            if (!(kVariant & 1)) { l0 += e[j]; l1 += e[j + 1]; }
            pk[j >> 1] = pack_half2(e[j], e[j + 1]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc));
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc));
          if (!(kVariant & 1)) { l0 += p0; l1 += p1; }
          pk[i >> 1] = pack_half2(p0, p1);
        }
      }
      if (!(kVariant & 2)) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int chunk = (c * 4 + w) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + (chunk << 4)) = make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]);
        }
      } else {
#pragma unroll
        for (int w = 0; w < 16; ++w) l1 += __uint_as_float(pk[w] & 0x3f800000u);
      }
    }
    L += l0 + l1;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  if (L == 123.456f) sink[0] = L + static_cast<float>(sm[threadIdx.x]);
}

}  // namespace

// which = 0: TMEM read, returns bytes per clock per SM; which = 1: ex2 (param = with_fma), returns ex2 per clock per SM
int microbench(int which, int warps, int param, int iters, float* result) {
  VCB_CHECK(result != nullptr && warps >= 1 && warps <= 32 && iters >= 1, kErrInvalid, "microbench: bad arguments");
  const int n_sm = device_sm_count();
  unsigned long long* cyc = nullptr;
  VCB_CUDA(cudaMalloc(reinterpret_cast<void**>(&cyc), n_sm * 8 + 64));
  VCB_CUDA(cudaMemset(cyc, 0, n_sm * 8 + 64));
  void* sink = reinterpret_cast<char*>(cyc) + n_sm * 8;
  for (int rep = 0; rep < 2; ++rep) {
    if (which == 0) {
      if (param <= 1) tmem_read_bench_kernel<1><<<n_sm, warps * 32>>>(iters, cyc, static_cast<uint32_t*>(sink));
      else if (param == 2) tmem_read_bench_kernel<2><<<n_sm, warps * 32>>>(iters, cyc, static_cast<uint32_t*>(sink));
      else tmem_read_bench_kernel<4><<<n_sm, warps * 32>>>(iters, cyc, static_cast<uint32_t*>(sink));
    } else if (which == 1) {
      ex2_bench_kernel<<<n_sm, warps * 32>>>(iters, param, cyc, static_cast<float*>(sink));
    } else if (which == 3) {
      const size_t smb = static_cast<size_t>((warps + 3) / 4) * 16384;
      float* sk = static_cast<float*>(sink);
#define VCB_SLB(V)                                                                                                  \
  case V:                                                                                                           \
    cudaFuncSetAttribute(softmax_loop_bench_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);        \
    softmax_loop_bench_kernel<V><<<n_sm, warps * 32, smb>>>(iters, cyc, sk);                                        \
    break;
      switch (param & 7) { VCB_SLB(0) VCB_SLB(1) VCB_SLB(2) VCB_SLB(3) VCB_SLB(4) VCB_SLB(5) VCB_SLB(6) VCB_SLB(7) }
#undef VCB_SLB
    } else {
      ex2_h2_bench_kernel<<<n_sm, warps * 32>>>(iters, cyc, static_cast<float*>(sink));
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> h(n_sm);
  if (e == cudaSuccess) e = cudaMemcpy(h.data(), cyc, n_sm * 8, cudaMemcpyDeviceToHost);
  cudaFree(cyc);
  VCB_CUDA(e);
  double sum = 0;
  for (auto c : h) sum += static_cast<double>(c);
  const double avg = sum / n_sm;
  if (which == 0) {
    const int ilp = param <= 1 ? 1 : param == 2 ? 2 : 4;
    *result = static_cast<float>(static_cast<double>(warps) * iters * ilp * 4096.0 / avg);
  } else if (which == 1) {
    *result = static_cast<float>(static_cast<double>(warps) * 32 * iters * 8.0 / avg);
  } else if (which == 3) {
    *result = static_cast<float>(static_cast<double>(warps) * 32 * iters * 64.0 / avg);  // exponentials per clock per SM
  } else {
    *result = static_cast<float>(static_cast<double>(warps) * 32 * iters * 16.0 / avg);  // two exponentials per instruction
  }
  return kOk;
}

}  // namespace vcb

extern "C" int vcb_microbench(int32_t which, int32_t warps, int32_t param, int32_t iters, float* result) {
  return vcb::microbench(which, warps, param, iters, result);
}
