// Weight-streaming "skinny" GEMM for decode and lm_head:  Y[B,N] = X[B,K] · W[N,K]^T,  B <= 32.
// Replaces the per-token nn.Linear calls of HF LlamaModel at decode (modeling_llama.py:171-184,224-) and the
// lm_head on the last position (reference vcoder_ds_llava_llama.py:92-93).  HBM-bound: every weight byte is read
// exactly once per step with 16-byte loads; the batch rides along as the N=8 dimension of mma.m16n8k16
// (W tile = A operand, 16 output features x 16 k; X^T = B operand), so no tensor-core padding is wasted.
//
// K-permutation trick: a dot product is invariant under any permutation of k applied to both operands, so each
// thread loads 8 CONTIGUOUS halves of a W row (one LDG.128) and of an X row and feeds them to two MMAs as if they
// were the strided fragment elements the instruction expects.
//
// Programmatic dependent launch: the kernel starts while its predecessor drains, issues its first weight loads
// (weights never depend on the predecessor), and only then waits on the grid dependency before touching X — so the
// HBM stream does not stop at kernel boundaries.
#include <cstdlib>
#include <string>

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
// Asynchronous bulk prefetch of `bytes` (multiple of 16) contiguous bytes into L2: keeps the HBM queues full without
// holding registers or L1 miss slots; the demand loads that follow hit L2.
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, uint32_t b) {
  __half2 r = __hmul2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

constexpr int U = 4;  // 32-wide k chunks per batch of loads: 8 independent 16-byte weight loads per thread

// WARPS = split-K factor inside the CTA: small-N problems use more warps per 16-row tile so that enough
// 16-byte loads are in flight per SM to cover HBM latency.
template <int NB, int WARPS, bool DB>
__global__ void __launch_bounds__(WARPS * 32)
skinny_gemm_kernel(SkinnyParams p) {
  constexpr int kThreads = WARPS * 32;
  __shared__ float red[WARPS][16][8 * NB + 1];
  __shared__ float s_ssq[WARPS][8 * NB];
  griddep_launch();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const __half* w_lo = p.W + static_cast<long>(row0 + g) * p.K + t * 8;
  const __half* w_hi = w_lo + 8L * p.K;

  // chunk -> warp mapping: mode 0 interleaves 64-byte chunks over the warps of the CTA; mode 1 gives every warp one
  // contiguous k-range so a warp streams each of its 16 rows sequentially (U*64 contiguous bytes per row per batch)
  const int n_chunks = p.K / 32;
  int c, c_end, c_step;
  if (p.mode == 0) {
    c = warp; c_end = n_chunks; c_step = WARPS;
  } else {
    const int per = (n_chunks + WARPS - 1) / WARPS;
    c = warp * per; c_end = min(n_chunks, c + per); c_step = 1;
  }
  auto load_w = [&](int cc, uint4 (&wl)[U], uint4 (&wh)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ch = cc + u * c_step;
      if (ch < c_end) {
        wl[u] = ldg_stream(w_lo + ch * 32);
        wh[u] = ldg_stream(w_hi + ch * 32);
      } else {
        wl[u] = make_uint4(0, 0, 0, 0);
        wh[u] = make_uint4(0, 0, 0, 0);
      }
    }
  };

  // L2 prefetch: lane r (< 16) owns row r of the tile and runs kPfDist batches ahead of the demand loads
  constexpr int kPfDist = 4;
  const __half* pf_row = p.W + static_cast<long>(row0 + (lane & 15)) * p.K;
  auto prefetch_batch = [&](int cc) {
    if (p.prefetch && lane < 16 && cc < c_end) {
      if (c_step == 1) {
        const int n = min(U, c_end - cc);
        bulk_prefetch_l2(pf_row + cc * 32, n * 64);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (cc + u * c_step < c_end) bulk_prefetch_l2(pf_row + (cc + u * c_step) * 32, 64);
      }
    }
  };
#pragma unroll
  for (int d = 1; d <= kPfDist; ++d) prefetch_batch(c + d * U * c_step);

  uint4 wl[U], wh[U];
  load_w(c, wl, wh);   // weights do not depend on the previous kernel: start streaming before the dependency wait
  griddep_wait();      // X (and the residual) are produced by the previous kernel in the stream

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

  float ssq[NB];  // fused RMSNorm: this thread's share of sum(x^2) for batch row nb*8+g
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) ssq[nb] = 0.f;
  const bool fuse_norm = p.norm_w != nullptr;

  for (; c < c_end; c += U * c_step) {
    uint4 xv[U][NB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ch = c + u * c_step;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        xv[u][nb] = (xok[nb] && ch < c_end) ? __ldg(reinterpret_cast<const uint4*>(xrow[nb] + ch * 32))
                                             : make_uint4(0, 0, 0, 0);
      }
      if (fuse_norm && ch < c_end) {
        // y = rstd * W (x o w_norm): the norm weight is folded into X here (fp16 x fp16 products are exact in fp32, so
        // one HMUL2 rounding == fp32 multiply + round), rstd is applied to the fp32 accumulator in the epilogue.
        const uint4 wn = __ldg(reinterpret_cast<const uint4*>(p.norm_w + ch * 32 + t * 8));
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const float2 a = unpack_half2(xv[u][nb].x), b2 = unpack_half2(xv[u][nb].y), c2 = unpack_half2(xv[u][nb].z),
                       d2 = unpack_half2(xv[u][nb].w);
          ssq[nb] += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y + c2.x * c2.x + c2.y * c2.y + d2.x * d2.x + d2.y * d2.y;
          xv[u][nb].x = hmul2_u32(xv[u][nb].x, wn.x);
          xv[u][nb].y = hmul2_u32(xv[u][nb].y, wn.y);
          xv[u][nb].z = hmul2_u32(xv[u][nb].z, wn.z);
          xv[u][nb].w = hmul2_u32(xv[u][nb].w, wn.w);
        }
      }
    }
    uint4 nl[U], nh[U];  // DB: next batch of weights in flight while this one is multiplied
    if (DB) load_w(c + U * c_step, nl, nh);
    prefetch_batch(c + (kPfDist + 1) * U * c_step);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        mma_16816(acc[nb], wl[u].x, wh[u].x, wl[u].y, wh[u].y, xv[u][nb].x, xv[u][nb].y);
        mma_16816(acc[nb], wl[u].z, wh[u].z, wl[u].w, wh[u].w, xv[u][nb].z, xv[u][nb].w);
      }
    }
    if (DB) {
#pragma unroll
      for (int u = 0; u < U; ++u) { wl[u] = nl[u]; wh[u] = nh[u]; }
    } else {
      load_w(c + U * c_step, wl, wh);
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
  if (fuse_norm) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float v = ssq[nb];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0) s_ssq[warp][nb * 8 + g] = v;
    }
  }
  __syncthreads();
  auto rstd_of = [&](int n) -> float {
    if (!fuse_norm) return 1.f;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) v += s_ssq[w][n];
    return rsqrtf(v / static_cast<float>(p.K) + p.norm_eps);
  };
  if (p.swiglu) {
    for (int idx = tid; idx < 8 * 8 * NB; idx += kThreads) {
      const int e = idx & 7, n = idx >> 3;
      if (n < p.B) {
        float gate = 0.f, up = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { gate += red[w][e][n]; up += red[w][e + 8][n]; }
        const float rs = rstd_of(n);
        gate *= rs;
        up *= rs;
        p.Y[static_cast<long>(n) * p.ldy + blockIdx.x * 8 + e] = __float2half_rn(silu_f(gate) * up);
      }
    }
  } else {
    for (int idx = tid; idx < 16 * 8 * NB; idx += kThreads) {
      const int r = idx & 15, n = idx >> 4;
      if (n < p.B) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += red[w][r][n];
        v *= rstd_of(n);
        if (p.residual) v += __half2float(p.residual[static_cast<long>(n) * p.ldr + row0 + r]);
        p.Y[static_cast<long>(n) * p.ldy + row0 + r] = __float2half_rn(v);
      }
    }
  }
}

template <int NB, int WARPS>
int launch_sk(const SkinnyParams& p, int grid, cudaStream_t stream) {
  if (p.double_buffer)
    VCB_CUDA(launch_pdl(skinny_gemm_kernel<NB, WARPS, true>, dim3(grid), dim3(WARPS * 32), 0, stream, p));
  else
    VCB_CUDA(launch_pdl(skinny_gemm_kernel<NB, WARPS, false>, dim3(grid), dim3(WARPS * 32), 0, stream, p));
  return kOk;
}
template <int NB>
int launch_sk_w(const SkinnyParams& p, int grid, int warps, cudaStream_t stream) {
  if (warps == 16) return launch_sk<NB, 16>(p, grid, stream);
  if (warps == 8) return launch_sk<NB, 8>(p, grid, stream);
  return launch_sk<NB, 4>(p, grid, stream);
}

}  // namespace

int skinny_gemm_tma(const SkinnyParams& p, cudaStream_t stream);
int skinny_gemm_tc5(const SkinnyParams& p, cudaStream_t stream);
bool skinny_tc5_supported(const SkinnyParams& p);

int skinny_gemm(const SkinnyParams& p, cudaStream_t stream) {
  if (p.pn_parts != nullptr || p.cn_rstd != nullptr) {  // the chained RMSNorm lives in the TMA kernel only
    VCB_CHECK(p.K % 64 == 0, kErrInvalid, "skinny_gemm: the chained RMSNorm needs K % 64 == 0");
    return skinny_gemm_tma(p, stream);
  }
  {
    // 0 = LDG kernel (this file), 1 = TMA-fed mma.sync kernel, one CTA per tile (skinny_tma.cu), 2 = tcgen05 (skinny_tc5.cu),
    // 3 = TMA-fed, balanced multi-tile CTAs (skinny_stream.cu; default; bit-identical to 1)
    static int impl = -1;
    if (impl < 0) {
      const char* e = getenv("VCB_SKINNY_IMPL");
      const std::string v = e ? e : "";
      impl = v == "ldg" ? 0 : (v == "tc5" ? 2 : (v == "tma" ? 1 : 3));
    }
    if (impl == 2 && skinny_tc5_supported(p)) return skinny_gemm_tc5(p, stream);
    if (impl == 3 && p.K % 64 == 0 && p.norm_w == nullptr && p.Wp == nullptr) return skinny_gemm_stream(p, stream);
    if (impl >= 1 && p.K % 64 == 0) return skinny_gemm_tma(p, stream);
  }
  VCB_CHECK(p.B >= 1 && p.B <= 32, kErrInvalid, "skinny_gemm: batch must be in [1, 32]");
  VCB_CHECK(p.N % 16 == 0 && p.K % 32 == 0, kErrInvalid, "skinny_gemm: N % 16 and K % 32 must be 0");
  VCB_CHECK(p.ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
            kErrInvalid, "skinny_gemm: X/W must be 16-byte aligned");
  const int grid = p.N / 16;
  const int nb = (p.B + 7) / 8;
  // aim for >= ~32 resident warps per SM
  int warps = grid <= 512 ? 16 : 8;  // measured: 8 warps + contiguous k-ranges (mode 1) is the best LDG variant
  SkinnyParams pp = p;
  {
    static int env_mode = -1;
    if (env_mode < 0) { const char* e = getenv("VCB_SKINNY_MODE"); env_mode = e ? atoi(e) : 1; }
    pp.mode = env_mode;
    static int env_pf = -1, env_db = -1, env_w = -1;
    if (env_pf < 0) { const char* e = getenv("VCB_SKINNY_PREFETCH"); env_pf = e ? atoi(e) : 0; }
    if (env_db < 0) { const char* e = getenv("VCB_SKINNY_DB"); env_db = e ? atoi(e) : 0; }
    if (env_w < 0) { const char* e = getenv("VCB_SKINNY_WARPS"); env_w = e ? atoi(e) : 0; }
    pp.prefetch = env_pf;
    pp.double_buffer = env_db;
    if (env_w == 4 || env_w == 8 || env_w == 16) warps = env_w;
  }
  KernelScope scope(kCatSkinny, stream);
  switch (nb) {
    case 1: return launch_sk_w<1>(pp, grid, warps, stream);
    case 2: return launch_sk_w<2>(pp, grid, warps, stream);
    case 3: return launch_sk_w<3>(pp, grid, warps, stream);
    default: return launch_sk_w<4>(pp, grid, warps, stream);
  }
}

}  // namespace vcb
