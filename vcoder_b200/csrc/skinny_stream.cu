// Weight-streaming skinny GEMM of the decode step, multi-tile ("persistent") form:  Y[B,N] = X[B,K] · W[N,K]^T,  B <= 32.
// Replaces the per-token nn.Linear calls of LlamaDecoderLayer / lm_head at decode time
// (transformers/models/llama/modeling_llama.py:171-184, 199-222 projections; reached from
// VCoder*LlavaLlamaForCausalLM.forward, vcoder_ds_llava_llama.py:81-93).
//
// Same arithmetic, in the same order, as skinny_tma_kernel<NB, 4, 2, 4> (a 16-row tile of W is the A operand of
// mma.m16n8k16, the batch is its n = 8; four warps split K, each with its own ring of 8 KiB TMA boxes): results are
// bit-identical.  What changes is the work distribution.  With one CTA per 16-row tile, the N = 22016 gate|up GEMM is 1376
// CTAs on 444 resident slots: three full waves that stream at ~7 TB/s and a fourth wave of 44 CTAs that takes as long as a
// full one (device timeline, profiles/r02_trace_decode_split_combine.csv: release spread 24.6 us = 3 x the 8.2 us CTA
// time, body 31.6 us); qkv (768 tiles) and lm_head (2000 tiles) lose the same way.  Here every CTA owns `tpc` CONSECUTIVE
// tiles with tpc = ceil(tiles / resident slots), so the grid is a single balanced wave (gate|up: 344 CTAs x 4 tiles), and a
// warp's TMA ring runs straight through its tiles: while the CTA reduces and stores tile i, the boxes of tile i+1 are
// already in flight — no CTA start-up bubble between tiles either.
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
int make_tmap_w3d(CUtensorMap* m, const void* ptr, uint64_t K, uint64_t N, uint32_t box_kblocks);
namespace {

constexpr int kKBlk = 64;    // k elements per 128-byte swizzled row
constexpr int kWarps = 4;
constexpr int kGroup = 4;    // k-blocks per stage: 4 x 16 rows x 128 B = 8 KiB
constexpr int kStageBytes = kGroup * 16 * kKBlk * 2;

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + fast_exp2(-1.4426950408889634f * x)); }

template <int NB, int kDepth>
__global__ void __launch_bounds__(kWarps * 32)
skinny_stream_kernel(const __grid_constant__ CUtensorMap tmW, SkinnyParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // [warps][kDepth] stages | red[2] | barriers
  typedef float RedT[kWarps][16][8 * NB + 1];
  RedT* red = reinterpret_cast<RedT*>(smem + kWarps * kDepth * kStageBytes);  // double-buffered by tile parity
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWarps * kDepth * kStageBytes + 2 * sizeof(RedT) + 8);
  griddep_launch();
  trace_mark(p.trace, blockIdx.x, 0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t ring = smem_u32(smem) + warp * (kDepth * kStageBytes);
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int n_tiles = p.N / 16;
  const int tile0 = blockIdx.x * p.tpc;
  const int my_tiles = min(p.tpc, n_tiles - tile0);
  // this warp's range of k-groups (a group = 4 k-blocks = 256 k): the same split of K as skinny_tma_kernel<NB,4,*,4>
  const int n_kblocks = p.K / kKBlk;
  const int n_groups = (n_kblocks + kGroup - 1) / kGroup;
  const int per = (n_groups + kWarps - 1) / kWarps;
  const int g_begin = min(n_groups, warp * per), g_end = min(n_groups, g_begin + per);
  const int my_groups = g_end - g_begin;
  const int total = my_tiles * my_groups;  // stages this warp streams, tile after tile

  // stage number q of this warp = k-group (q % my_groups) of tile (q / my_groups)
  auto issue = [&](int q) {
    const int s = q % kDepth;
    const uint32_t b = smem_u32(&bars[warp * kDepth + s]);
    const int tile = tile0 + q / my_groups, grp = g_begin + q % my_groups;
    mbar_expect_tx(b, kStageBytes);
    tma_load_3d(ring + s * kStageBytes, &tmW, b, 0, tile * 16, grp * kGroup);
  };
  if (lane == 0) {
    for (int s = 0; s < kDepth; ++s) mbar_init(smem_u32(&bars[warp * kDepth + s]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
    for (int q = 0; q < kDepth && q < total; ++q) issue(q);  // weights do not depend on the previous kernel
  }
  __syncwarp();
  griddep_wait();  // X (and the residual) come from the previous kernel
  trace_mark(p.trace, blockIdx.x, 1);

  const __half* xrow[NB];
  bool xok[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    xok[nb] = (nb * 8 + g) < p.B;
    xrow[nb] = p.X + static_cast<long>(xok[nb] ? nb * 8 + g : 0) * p.ldx + t * 8;
  }
  // X fragments of one k-group (8 chunks of 32 k); the K-permutation trick: thread t feeds 16-byte piece t of a chunk
  auto load_x = [&](int i, uint4 (&dst)[2 * kGroup][NB]) {
    const int k0 = (g_begin + i) * kGroup * kKBlk;
#pragma unroll
    for (int c = 0; c < 2 * kGroup; ++c) {
      const int k = k0 + c * 32;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
        dst[c][nb] = (xok[nb] && k < p.K) ? __ldg(reinterpret_cast<const uint4*>(xrow[nb] + k)) : make_uint4(0, 0, 0, 0);
    }
  };
  const uint32_t r_lo = static_cast<uint32_t>(g) * 128, r_hi = static_cast<uint32_t>(g + 8) * 128;
  constexpr int kAcc = 4;
  constexpr int kThreads = kWarps * 32;

  uint4 xv[2 * kGroup][NB], xnext[2 * kGroup][NB];
  if (my_groups > 0) load_x(0, xv);
  int q = 0;
  for (int ti = 0; ti < my_tiles; ++ti) {
    float acc[NB][kAcc][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int a = 0; a < kAcc; ++a) { acc[nb][a][0] = acc[nb][a][1] = acc[nb][a][2] = acc[nb][a][3] = 0.f; }
    for (int i = 0; i < my_groups; ++i, ++q) {
      const int s = q % kDepth;
      const uint32_t ph = static_cast<uint32_t>((q / kDepth) & 1);
      // next group's activations travel while this stage is consumed (X depends on k only: the first group of the next
      // tile is the first group again)
      if (i + 1 < my_groups) load_x(i + 1, xnext);
      else if (ti + 1 < my_tiles) load_x(0, xnext);
      mbar_wait(smem_u32(&bars[warp * kDepth + s]), ph);
      const uint32_t st = ring + s * kStageBytes;
#pragma unroll
      for (int kb = 0; kb < kGroup; ++kb) {
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          // 32-k chunk (kb, hlf): thread t takes 16-byte chunk 4*hlf + t of rows g and g+8 (swizzle: chunk ^ (row & 7))
          const uint32_t ch = static_cast<uint32_t>(4 * hlf + t);
          const uint4 wl = lds128(st + kb * 2048 + r_lo + ((ch ^ static_cast<uint32_t>(g & 7)) << 4));
          const uint4 wh = lds128(st + kb * 2048 + r_hi + ((ch ^ static_cast<uint32_t>((g + 8) & 7)) << 4));
          const int a0 = (kb * 2 + hlf) & (kAcc - 1);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const uint4 x = xv[kb * 2 + hlf][nb];
            mma_16816(acc[nb][a0], wl.x, wh.x, wl.y, wh.y, x.x, x.y);
            mma_16816(acc[nb][(a0 + 2) & (kAcc - 1)], wl.z, wh.z, wl.w, wh.w, x.z, x.w);
          }
        }
      }
      __syncwarp();  // every lane is done with stage s before it is refilled
      if (lane == 0 && q + kDepth < total) issue(q + kDepth);
#pragma unroll
      for (int c = 0; c < 2 * kGroup; ++c)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) xv[c][nb] = xnext[c][nb];
    }
    // ---- tile done: fold the accumulator sets, cross-warp (split-K) reduction through shared memory, epilogue
    RedT& rd = red[ti & 1];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float a4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a4[j] = (acc[nb][0][j] + acc[nb][1][j]) + (acc[nb][2][j] + acc[nb][3][j]);
      rd[warp][g][nb * 8 + 2 * t] = a4[0];
      rd[warp][g][nb * 8 + 2 * t + 1] = a4[1];
      rd[warp][g + 8][nb * 8 + 2 * t] = a4[2];
      rd[warp][g + 8][nb * 8 + 2 * t + 1] = a4[3];
    }
    __syncthreads();  // one barrier per tile: the buffer of tile i-1 is only rewritten by tile i+1, behind this barrier
    const int tile = tile0 + ti;
    const int row0 = tile * 16;
    if (p.swiglu) {
      for (int idx = tid; idx < 8 * 8 * NB; idx += kThreads) {
        const int e = idx & 7, n = idx >> 3;
        if (n < p.B) {
          float gate = 0.f, up = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) { gate += rd[w][e][n]; up += rd[w][e + 8][n]; }
          p.Y[static_cast<long>(n) * p.ldy + tile * 8 + e] = __float2half_rn(silu_f(gate) * up);
        }
      }
    } else {
      for (int idx = tid; idx < 16 * 8 * NB; idx += kThreads) {
        const int r = idx & 15, n = idx >> 4;
        if (n < p.B) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) v += rd[w][r][n];
          if (p.residual) v += __half2float(p.residual[static_cast<long>(n) * p.ldr + row0 + r]);
          p.Y[static_cast<long>(n) * p.ldy + row0 + r] = __float2half_rn(v);
        }
      }
    }
  }
  trace_mark(p.trace, blockIdx.x, 2);
}

template <int NB, int kDepth>
int run_stream(SkinnyParams p, int env_tpc, cudaStream_t stream) {
  constexpr int smem = kWarps * kDepth * kStageBytes + 2 * kWarps * 16 * (8 * NB + 1) * 4 + 8 + kWarps * kDepth * 8 + 16;
  static_assert(smem <= 227 * 1024, "skinny_stream: ring does not fit in shared memory");
  static int occ = 0;  // resident CTAs per SM of this instantiation (registers and shared memory)
  if (occ == 0) {
    VCB_CUDA(cudaFuncSetAttribute(skinny_stream_kernel<NB, kDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int n = 0;
    VCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, skinny_stream_kernel<NB, kDepth>, kWarps * 32, smem));
    occ = n > 0 ? n : 1;
  }
  const int tiles = p.N / 16;
  const int slots = occ * device_sm_count();
  // one balanced wave: every CTA streams the same number of consecutive tiles (the last CTA may get fewer)
  int tpc = env_tpc > 0 ? env_tpc : (tiles + slots - 1) / slots;
  if (tpc < 1) tpc = 1;
  p.tpc = tpc;
  const int grid = (tiles + tpc - 1) / tpc;
  CUtensorMap tm;
  VCB_TRY(make_tmap_w3d(&tm, p.W, p.K, p.N, kGroup));
  p.trace = trace_slots(kTraceSkinny, grid, p.N);
  KernelScope scope(kCatSkinny, stream);
  VCB_CUDA(launch_pdl(skinny_stream_kernel<NB, kDepth>, dim3(grid), dim3(kWarps * 32), smem, stream, tm, p));
  return kOk;
}

}  // namespace

// VCB_SKINNY_TPC: 0 (default) = one balanced wave, tpc = ceil(tiles / resident CTAs); n > 0 = fixed tiles per CTA.
// VCB_SKINNY_STREAM_DEPTH = 2 (default) | 3: stages of 8 KiB per warp.
int skinny_gemm_stream(const SkinnyParams& p, cudaStream_t stream) {
  VCB_CHECK(p.B >= 1 && p.B <= 32, kErrInvalid, "skinny_gemm: batch must be in [1, 32]");
  VCB_CHECK(p.N % 16 == 0 && p.K % 64 == 0, kErrInvalid, "skinny_gemm_stream: N % 16 and K % 64 must be 0");
  VCB_CHECK(p.norm_w == nullptr && p.pn_parts == nullptr && p.cn_rstd == nullptr && p.Wp == nullptr, kErrInvalid,
            "skinny_gemm_stream: plain GEMM (+ residual / SwiGLU) only");
  static int env_tpc = -1, env_depth = -1;
  if (env_tpc < 0) { const char* e = getenv("VCB_SKINNY_TPC"); env_tpc = e ? atoi(e) : 0; }
  if (env_depth < 0) { const char* e = getenv("VCB_SKINNY_STREAM_DEPTH"); env_depth = e ? atoi(e) : 2; }
  const int nb = (p.B + 7) / 8;
  if (env_depth == 3) {
    switch (nb) {
      case 1: return run_stream<1, 3>(p, env_tpc, stream);
      case 2: return run_stream<2, 3>(p, env_tpc, stream);
      case 3: return run_stream<3, 3>(p, env_tpc, stream);
      default: return run_stream<4, 3>(p, env_tpc, stream);
    }
  }
  switch (nb) {
    case 1: return run_stream<1, 2>(p, env_tpc, stream);
    case 2: return run_stream<2, 2>(p, env_tpc, stream);
    case 3: return run_stream<3, 2>(p, env_tpc, stream);
    default: return run_stream<4, 2>(p, env_tpc, stream);
  }
}

}  // namespace vcb
