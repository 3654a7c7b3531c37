// Weight-streaming "skinny" GEMM for decode and lm_head:  Y[B,N] = X[B,K] · W[N,K]^T,  B <= 32.
// Replaces the per-token nn.Linear calls of HF LlamaModel at decode (modeling_llama.py:171-184,224-) and the
// lm_head on the last position (reference vcoder_ds_llava_llama.py:92-93).  HBM-bound: every weight byte is read
// exactly once per step with 16-byte loads; the batch rides along as the N=8 dimension of mma.m16n8k16
// (W tile = A operand, 16 output features x 16 k; X^T = B operand), so no tensor-core padding is wasted.
//
// K-permutation trick: a dot product is invariant under any permutation of k applied to both operands, so each
// thread loads 8 CONTIGUOUS halves of a W row (one LDG.128) and of an X row and feeds them to two MMAs as if they
// were the strided fragment elements the instruction expects.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {


__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// WARPS = split-K factor inside the CTA: small-N problems use more warps per 16-row tile so that enough
// 16-byte loads are in flight per SM to cover HBM latency.
template <int NB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
skinny_gemm_kernel(SkinnyParams p) {
  constexpr int kSkWarps = WARPS;
  constexpr int kSkThreads = WARPS * 32;
  __shared__ float red[kSkWarps][16][8 * NB + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const __half* w_lo = p.W + static_cast<long>(row0 + g) * p.K + t * 8;
  const __half* w_hi = w_lo + 8L * p.K;
  const __half* xrow[NB];
  bool xok[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    xok[nb] = (nb * 8 + g) < p.B;
    xrow[nb] = p.X + static_cast<long>(xok[nb] ? nb * 8 + g : 0) * p.ldx + t * 8;
  }
  float acc[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) { acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f; }

  const int n_chunks = p.K / 32;
  constexpr int U = 4;  // chunks in flight per warp iteration: 8 independent 16-byte weight loads per thread
  int c = warp;
  for (; c + (U - 1) * kSkWarps < n_chunks; c += U * kSkWarps) {
    uint4 wl[U], wh[U], xv[U][NB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k0 = (c + u * kSkWarps) * 32;
      wl[u] = ldg_stream(w_lo + k0);
      wh[u] = ldg_stream(w_hi + k0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k0 = (c + u * kSkWarps) * 32;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        xv[u][nb] = xok[nb] ? __ldg(reinterpret_cast<const uint4*>(xrow[nb] + k0)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        mma_16816(acc[nb], wl[u].x, wh[u].x, wl[u].y, wh[u].y, xv[u][nb].x, xv[u][nb].y);
        mma_16816(acc[nb], wl[u].z, wh[u].z, wl[u].w, wh[u].w, xv[u][nb].z, xv[u][nb].w);
      }
    }
  }
  for (; c < n_chunks; c += kSkWarps) {
    const int k0 = c * 32;
    const uint4 wl = ldg_stream(w_lo + k0);
    const uint4 wh = ldg_stream(w_hi + k0);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const uint4 xv = xok[nb] ? __ldg(reinterpret_cast<const uint4*>(xrow[nb] + k0)) : make_uint4(0, 0, 0, 0);
      mma_16816(acc[nb], wl.x, wh.x, wl.y, wh.y, xv.x, xv.y);
      mma_16816(acc[nb], wl.z, wh.z, wl.w, wh.w, xv.z, xv.w);
    }
  }

  // cross-warp (split-K) reduction through shared memory
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    red[warp][g][nb * 8 + 2 * t] = acc[nb][0];
    red[warp][g][nb * 8 + 2 * t + 1] = acc[nb][1];
    red[warp][g + 8][nb * 8 + 2 * t] = acc[nb][2];
    red[warp][g + 8][nb * 8 + 2 * t + 1] = acc[nb][3];
  }
  __syncthreads();
  if (p.swiglu) {
    for (int idx = tid; idx < 8 * 8 * NB; idx += kSkThreads) {
      const int e = idx & 7, n = idx >> 3;
      if (n < p.B) {
        float gate = 0.f, up = 0.f;
#pragma unroll
        for (int w = 0; w < kSkWarps; ++w) { gate += red[w][e][n]; up += red[w][e + 8][n]; }
        p.Y[static_cast<long>(n) * p.ldy + blockIdx.x * 8 + e] = __float2half_rn(silu_f(gate) * up);
      }
    }
  } else {
    for (int idx = tid; idx < 16 * 8 * NB; idx += kSkThreads) {
      const int r = idx & 15, n = idx >> 4;
      if (n < p.B) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kSkWarps; ++w) v += red[w][r][n];
        if (p.residual) v += __half2float(p.residual[static_cast<long>(n) * p.ldr + row0 + r]);
        p.Y[static_cast<long>(n) * p.ldy + row0 + r] = __float2half_rn(v);
      }
    }
  }
}

}  // namespace

int skinny_gemm(const SkinnyParams& p, cudaStream_t stream) {
  VCB_CHECK(p.B >= 1 && p.B <= 32, kErrInvalid, "skinny_gemm: batch must be in [1, 32]");
  VCB_CHECK(p.N % 16 == 0 && p.K % 32 == 0, kErrInvalid, "skinny_gemm: N % 16 and K % 32 must be 0");
  VCB_CHECK(p.ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
            kErrInvalid, "skinny_gemm: X/W must be 16-byte aligned");
  const int grid = p.N / 16;
  const int nb = (p.B + 7) / 8;
  // aim for >= ~32 resident warps per SM
  const int warps = grid <= 512 ? 16 : (grid <= 1024 ? 8 : 4);
  KernelScope scope(kCatSkinny, stream);
#define VCB_SK(NB_, W_) skinny_gemm_kernel<NB_, W_><<<grid, W_ * 32, 0, stream>>>(p)
#define VCB_SKW(NB_)                                     \
  do {                                                   \
    if (warps == 16) VCB_SK(NB_, 16);                    \
    else if (warps == 8) VCB_SK(NB_, 8);                 \
    else VCB_SK(NB_, 4);                                 \
  } while (0)
  switch (nb) {
    case 1: VCB_SKW(1); break;
    case 2: VCB_SKW(2); break;
    case 3: VCB_SKW(3); break;
    default: VCB_SKW(4); break;
  }
#undef VCB_SKW
#undef VCB_SK
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb
