// TMA-fed variant of the weight-streaming skinny GEMM (decode / lm_head):  Y[B,N] = X[B,K] · W[N,K]^T,  B <= 32.
// Same math and the same K-permutation MMA trick as skinny.cu, but the weight stream is driven by the TMA engine:
// every warp owns a K-range of the CTA's 16-row tile and keeps a ring of kDepth bulk-tensor loads in flight
// ([4 k-blocks][16 rows][64 k] = 8 KiB each, 128B-swizzled so the fragment reads are bank-conflict free).  The bytes
// in flight per SM (up to ~190 KiB) no longer depend on registers or on L1 miss slots, which is what the plain-LDG
// kernel runs out of.  Programmatic dependent launch: ring fill happens BEFORE the grid-dependency wait — weights do
// not depend on the previous kernel, so HBM keeps streaming across kernel boundaries.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kKBlk = 64;               // k elements per 128-byte swizzled row
constexpr int kPrefetchGroupsDefault = 0;  // see DESIGN.md section 3.1 for the measurement behind this default

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
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
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, uint32_t b) {
  __half2 r = __hmul2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + fast_exp2(-1.4426950408889634f * x)); }

// 1-D bulk copy global -> shared, completion on an mbarrier (no tensor map: the source is already the smem image)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// L2 prefetches (no shared-memory destination, no completion tracking): issued BEFORE the grid-dependency wait for the
// k-groups right behind the ring, so HBM keeps working through the kernel boundary and the first ring refills hit L2
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// W viewed as a 3-D tensor [K/64][N][64]: dim0 = k within a 64-block, dim1 = row, dim2 = k-block.
// PACKED: the weight stream comes from the tile-major copy made by pack_skinny_weights (p.Wp): the 16 x K slab of a CTA is
// one contiguous 32*K-byte run of 2 KiB blocks [k-block][16 rows][128 B, chunks pre-swizzled], so a stage is ONE linear
// 8 KiB bulk copy and every DRAM page the stream touches is consumed completely (the row-major source gives 512-byte runs
// at an 8 KiB stride).
template <int NB, int WARPS, int kDepth, int kGroup, bool PACKED, bool CHAIN>
__global__ void __launch_bounds__(WARPS * 32)
skinny_tma_kernel(const __grid_constant__ CUtensorMap tmW, SkinnyParams p) {
  constexpr int kStageBytes = kGroup * 16 * kKBlk * 2;  // kGroup k-blocks x 16 rows x 128 B
  extern __shared__ __align__(1024) uint8_t smem[];  // [WARPS][kDepth] stages | red | barriers (no static smem: base stays 1 KiB aligned)
  typedef float RedT[16][8 * NB + 1];
  RedT* red = reinterpret_cast<RedT*>(smem + WARPS * kDepth * kStageBytes);
  typedef uint64_t BarT[kDepth];
  BarT* bars = reinterpret_cast<BarT*>(smem + WARPS * kDepth * kStageBytes + sizeof(RedT) * WARPS + 8);
  float* s_ssq = reinterpret_cast<float*>(smem + WARPS * kDepth * kStageBytes + sizeof(RedT) * WARPS + 8 + sizeof(BarT) * WARPS);  // [WARPS][8*NB]
  griddep_launch();
  trace_mark(p.trace, blockIdx.x, 0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const uint32_t ring = smem_u32(smem) + warp * (kDepth * kStageBytes);
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // this warp's range of k-groups (a group = 4 k-blocks = 256 k)
  const int n_kblocks = p.K / kKBlk;
  const int n_groups = (n_kblocks + kGroup - 1) / kGroup;
  const int per = (n_groups + WARPS - 1) / WARPS;
  const int g_begin = min(n_groups, warp * per), g_end = min(n_groups, g_begin + per);
  const int my_groups = g_end - g_begin;

  // one stage = k-group `grp` of this CTA's 16 rows
  auto issue = [&](int s, int grp) {
    const uint32_t b = smem_u32(&bars[warp][s]);
    if constexpr (PACKED) {
      const int kb0 = grp * kGroup;
      const uint32_t bytes = static_cast<uint32_t>(min(kGroup, n_kblocks - kb0)) * 2048u;  // last group may be short
      mbar_expect_tx(b, bytes);
      bulk_load_1d(ring + s * kStageBytes,
                   reinterpret_cast<const uint8_t*>(p.Wp) + (static_cast<size_t>(blockIdx.x) * n_kblocks + kb0) * 2048u, bytes, b);
    } else {
      mbar_expect_tx(b, kStageBytes);
      tma_load_3d(ring + s * kStageBytes, &tmW, b, 0, row0, grp * kGroup);
    }
  };
  if (lane == 0) {
    for (int s = 0; s < kDepth; ++s) mbar_init(smem_u32(&bars[warp][s]), 1);
    fence_mbar_init();
    if constexpr (!PACKED) tma_prefetch_desc(&tmW);
    for (int s = 0; s < kDepth && s < my_groups; ++s) issue(s, g_begin + s);  // fill the ring before waiting on the previous kernel
    // CTAs of the first wave sit at the grid-dependency wait while the previous kernel drains: pull the next k-groups of
    // their slab into L2 meanwhile (later waves start after the wait has long been released and would only add requests)
    if (p.pf_groups > 0 && static_cast<int>(blockIdx.x) < p.pf_ctas) {
      const int g0 = g_begin + kDepth, g1 = min(g_end, g0 + p.pf_groups);
      if constexpr (PACKED) {
        if (g1 > g0) {
          const int kb0 = g0 * kGroup, kb1 = min(n_kblocks, g1 * kGroup);
          bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.Wp) + (static_cast<size_t>(blockIdx.x) * n_kblocks + kb0) * 2048u,
                           static_cast<uint32_t>(kb1 - kb0) * 2048u);
        }
      } else {
        for (int grp = g0; grp < g1; ++grp) tma_prefetch_3d(&tmW, 0, row0, grp * kGroup);
      }
    }
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
  // four independent accumulator sets per batch tile: consecutive 32-k chunks go to different sets, so the mma.sync
  // dependency chain per stage is 4 deep instead of 16 (summed once at the end)
  constexpr int kAcc = 4;
  float acc[NB][kAcc][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int a = 0; a < kAcc; ++a) { acc[nb][a][0] = acc[nb][a][1] = acc[nb][a][2] = acc[nb][a][3] = 0.f; }

  float ssq[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) ssq[nb] = 0.f;
  const bool fuse_norm = p.norm_w != nullptr;
  const uint32_t r_lo = static_cast<uint32_t>(g) * 128, r_hi = static_cast<uint32_t>(g + 8) * 128;

  // X fragments of one k-group (8 chunks of 32 k), optionally with the fused RMSNorm weight folded in
  auto load_x = [&](int i, uint4 (&dst)[2 * kGroup][NB]) {
    const int k0 = (g_begin + i) * kGroup * kKBlk;
#pragma unroll
    for (int c = 0; c < 2 * kGroup; ++c) {
      const int k = k0 + c * 32;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
        dst[c][nb] = (xok[nb] && k < p.K) ? __ldg(reinterpret_cast<const uint4*>(xrow[nb] + k)) : make_uint4(0, 0, 0, 0);
      if (fuse_norm && k < p.K) {
        const uint4 wn = __ldg(reinterpret_cast<const uint4*>(p.norm_w + k + t * 8));
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const float2 a = unpack_half2(dst[c][nb].x), b2 = unpack_half2(dst[c][nb].y), c2 = unpack_half2(dst[c][nb].z),
                       d2 = unpack_half2(dst[c][nb].w);
          ssq[nb] += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y + c2.x * c2.x + c2.y * c2.y + d2.x * d2.x + d2.y * d2.y;
          dst[c][nb].x = hmul2_u32(dst[c][nb].x, wn.x);
          dst[c][nb].y = hmul2_u32(dst[c][nb].y, wn.y);
          dst[c][nb].z = hmul2_u32(dst[c][nb].z, wn.z);
          dst[c][nb].w = hmul2_u32(dst[c][nb].w, wn.w);
        }
      }
    }
  };

  uint4 xv[2 * kGroup][NB], xnext[2 * kGroup][NB];
  if (my_groups > 0) load_x(0, xv);
  // chained RMSNorm, consumer half: X is fp16(x * w_norm); rstd[b] was finalised by the producer's last CTA and scales the
  // fp32 accumulators in the epilogue (nothing extra in the streaming loop)
  const bool chain_in = CHAIN && p.cn_rstd != nullptr;  // CHAIN = false: the instantiation every other caller gets
  for (int i = 0; i < my_groups; ++i) {
    const int s = i % kDepth;
    const uint32_t ph = static_cast<uint32_t>((i / kDepth) & 1);
    if (i + 1 < my_groups) load_x(i + 1, xnext);  // next group's activations travel while this stage is consumed
    mbar_wait(smem_u32(&bars[warp][s]), ph);
    const uint32_t st = ring + s * kStageBytes;
#pragma unroll
    for (int kb = 0; kb < kGroup; ++kb) {
      // a short last group leaves stale bytes behind the copied k-blocks (the tensor-map path zero-fills them instead)
      if (PACKED && (g_begin + i) * kGroup + kb >= n_kblocks) continue;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        // 32-k chunk (kb, hlf): thread t takes 16-byte chunk 4*hlf + t of rows g and g+8 (swizzle: chunk ^ (row & 7))
        const uint32_t ch = static_cast<uint32_t>(4 * hlf + t);
        const uint4 wl = lds128(st + kb * 2048 + r_lo + ((ch ^ static_cast<uint32_t>(g & 7)) << 4));
        const uint4 wh = lds128(st + kb * 2048 + r_hi + ((ch ^ static_cast<uint32_t>((g + 8) & 7)) << 4));
        constexpr int kA0 = 0;
        const int a0 = (kb * 2 + hlf) & (kAcc - 1);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const uint4 x = xv[kb * 2 + hlf][nb];
          mma_16816(acc[nb][a0], wl.x, wh.x, wl.y, wh.y, x.x, x.y);
          mma_16816(acc[nb][(a0 + 2) & (kAcc - 1)], wl.z, wh.z, wl.w, wh.w, x.z, x.w);
        }
        (void)kA0;
      }
    }
    __syncwarp();  // every lane is done with stage s before it is refilled
    if (lane == 0 && i + kDepth < my_groups) issue(s, g_begin + i + kDepth);
    if (i + 1 < my_groups) {
#pragma unroll
      for (int c = 0; c < 2 * kGroup; ++c)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) xv[c][nb] = xnext[c][nb];
    }
  }
  // fold the accumulator sets
  float accs[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int j = 0; j < 4; ++j) accs[nb][j] = (acc[nb][0][j] + acc[nb][1][j]) + (acc[nb][2][j] + acc[nb][3][j]);

  // cross-warp (split-K) reduction through shared memory
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    red[warp][g][nb * 8 + 2 * t] = accs[nb][0];
    red[warp][g][nb * 8 + 2 * t + 1] = accs[nb][1];
    red[warp][g + 8][nb * 8 + 2 * t] = accs[nb][2];
    red[warp][g + 8][nb * 8 + 2 * t + 1] = accs[nb][3];
  }
  if (fuse_norm) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float v = ssq[nb];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0) s_ssq[warp * 8 * NB + nb * 8 + g] = v;
    }
  }
  __syncthreads();
  auto rstd_of = [&](int n) -> float {
    if (chain_in) return __ldcg(p.cn_rstd + n);
    if (!fuse_norm) return 1.f;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) v += s_ssq[w * 8 * NB + n];
    return rsqrtf(v / static_cast<float>(p.K) + p.norm_eps);
  };
  constexpr int kThreads = WARPS * 32;
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
    const bool chain_out = CHAIN && p.pn_parts != nullptr;
    const float pn_wr = chain_out ? __half2float(p.pn_w[row0 + (tid & 15)]) : 0.f;  // idx & 15 == tid & 15: kThreads % 16 == 0
    for (int idx = tid; idx < 16 * 8 * NB; idx += kThreads) {  // every iteration is taken by all threads or by none
      const int r = idx & 15, n = idx >> 4;
      float sq = 0.f;
      if (n < p.B) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += red[w][r][n];
        v *= rstd_of(n);
        if (p.residual) v += __half2float(p.residual[static_cast<long>(n) * p.ldr + row0 + r]);
        const __half h = __float2half_rn(v);
        p.Y[static_cast<long>(n) * p.ldy + row0 + r] = h;
        if (chain_out) {
          // chained RMSNorm, producer half: the norm sees the fp16 residual stream, exactly like the separate kernel
          const float xv = __half2float(h);
          sq = xv * xv;
          p.pn_out[static_cast<long>(n) * p.pn_ld + row0 + r] = __float2half_rn(xv * pn_wr);
        }
      }
      if (chain_out) {  // sum over this CTA's 16 columns: the 16 lanes that share n
        sq += __shfl_xor_sync(0xffffffffu, sq, 8);
        sq += __shfl_xor_sync(0xffffffffu, sq, 4);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        if (r == 0 && n < 32) p.pn_parts[static_cast<long>(blockIdx.x) * 32 + n] = sq;
      }
    }
    if (chain_out) {
      // The CTA that draws the last ticket folds every CTA's partial sum(x^2) (fixed order: deterministic) into
      // rstd[b] = rsqrt(mean(x^2) + eps): the RMSNorm statistic is complete when this kernel is, no norm kernel follows.
      __threadfence();
      __syncthreads();
      int* s_tk = reinterpret_cast<int*>(s_ssq);
      if (tid == 0) *s_tk = atomicAdd(p.pn_ticket, 1);
      __syncthreads();
      if (*s_tk == static_cast<int>(gridDim.x) - 1) {
        __threadfence();
        float* sred = reinterpret_cast<float*>(red);  // the split-K buffers are dead: [WARPS][32] floats
        const int n = tid & 31, chunk = tid >> 5;
        float v = 0.f;
        if (n < p.B)
          for (int c = chunk; c < static_cast<int>(gridDim.x); c += WARPS) v += __ldcg(p.pn_parts + static_cast<long>(c) * 32 + n);
        __syncthreads();
        sred[chunk * 32 + n] = v;
        __syncthreads();
        if (tid < p.B) {
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < WARPS; ++w) tot += sred[w * 32 + tid];
          p.pn_rstd[tid] = rsqrtf(tot / static_cast<float>(p.N) + p.pn_eps);
        }
        if (tid == 0) *p.pn_ticket = 0;  // ready for the next launch (stream-ordered)
      }
    }
  }
  trace_mark(p.trace, blockIdx.x, 2);
}

template <int NB, int WARPS, int kDepth, int kGroup, bool PACKED = false, bool CHAIN = false>
int launch_tma(const CUtensorMap& tm, const SkinnyParams& p, int grid, cudaStream_t stream) {
  constexpr int kStageBytes = kGroup * 16 * kKBlk * 2;
  constexpr int smem = WARPS * kDepth * kStageBytes + WARPS * 16 * (8 * NB + 1) * 4 + 8 + WARPS * kDepth * 8 + WARPS * 8 * NB * 4 + 16;
  static_assert(smem <= 227 * 1024, "skinny_tma: ring does not fit in shared memory");
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(skinny_tma_kernel<NB, WARPS, kDepth, kGroup, PACKED, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  VCB_CUDA(launch_pdl(skinny_tma_kernel<NB, WARPS, kDepth, kGroup, PACKED, CHAIN>, dim3(grid), dim3(WARPS * 32), smem, stream, tm, p));
  return kOk;
}
// packed weight stream: the measured-best ring shape (4 warps x 2 stages x 8 KiB), or 3 stages (VCB_SKINNY_TMA_DEPTH=3)
template <int NB>
int launch_packed(const SkinnyParams& p, int grid, int depth, cudaStream_t stream) {
#ifdef VCB_EXPERIMENTAL
  CUtensorMap unused;
  memset(&unused, 0, sizeof(unused));
  if (p.pn_parts != nullptr || p.cn_rstd != nullptr) return launch_tma<NB, 4, 2, 4, true, true>(unused, p, grid, stream);
  if (depth >= 3) return launch_tma<NB, 4, 3, 4, true>(unused, p, grid, stream);
  return launch_tma<NB, 4, 2, 4, true>(unused, p, grid, stream);
#else
  (void)p; (void)grid; (void)depth; (void)stream;
  set_thread_error("the tile-major packed weight stream is an experiment (no gain measured): rebuild with VCB_EXPERIMENTAL=1");
  return kErrInvalid;
#endif
}

// row-major W[N,K] -> tile-major, pre-swizzled copy (one thread per 16-byte chunk): chunk c of row n goes to
// ((tile * K/64 + kb) * 16 + r) * 8 + (ch ^ (r & 7)) with tile = n/16, r = n%16, kb = c/8, ch = c%8 — exactly the shared
// memory image the 128B-swizzled tensor-map load of the same k-block produces
__global__ void pack_skinny_kernel(const uint4* __restrict__ W, uint4* __restrict__ Wp, int N, int K) {
  const long cpr = K / 8;  // chunks per row
  const long total = static_cast<long>(N) * cpr;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = i / cpr, c = i % cpr;
    const long tile = n >> 4, r = n & 15, kb = c >> 3, ch = c & 7;
    Wp[((tile * (K / 64) + kb) * 16 + r) * 8 + (ch ^ (r & 7))] = W[i];
  }
}
template <int NB>
int launch_tma_w(const CUtensorMap& tm, const SkinnyParams& p, int grid, int warps, int depth, int group,
                 cudaStream_t stream) {
  // the chained RMSNorm (opt-in experiment) exists for the default ring shape only
  if (p.pn_parts != nullptr || p.cn_rstd != nullptr) return launch_tma<NB, 4, 2, 4, false, true>(tm, p, grid, stream);
#ifdef VCB_EXPERIMENTAL  // ring-shape sweep (tools/sweep_decode.py; no shape beat 4 warps x 2 stages x 8 KiB: DESIGN.md 3.1)
  if (group == 2) {
    if (depth <= 2) return launch_tma<NB, 4, 2, 2>(tm, p, grid, stream);
    if (depth == 3) return launch_tma<NB, 4, 3, 2>(tm, p, grid, stream);
    return launch_tma<NB, 4, 4, 2>(tm, p, grid, stream);
  }
  if (warps >= 8) {
    if (depth <= 2) return launch_tma<NB, 8, 2, 4>(tm, p, grid, stream);
    return launch_tma<NB, 8, 3, 4>(tm, p, grid, stream);
  }
  if (warps >= 4) {
    if (depth <= 2) return launch_tma<NB, 4, 2, 4>(tm, p, grid, stream);
    if (depth == 3) return launch_tma<NB, 4, 3, 4>(tm, p, grid, stream);
    return launch_tma<NB, 4, 4, 4>(tm, p, grid, stream);
  }
  if (depth <= 2) return launch_tma<NB, 2, 2, 4>(tm, p, grid, stream);
  return launch_tma<NB, 2, 3, 4>(tm, p, grid, stream);
#else
  (void)warps; (void)depth; (void)group;
  return launch_tma<NB, 4, 2, 4>(tm, p, grid, stream);
#endif
}

}  // namespace

int make_tmap_w3d(CUtensorMap* m, const void* ptr, uint64_t K, uint64_t N, uint32_t box_kblocks);

// Ring shape per GEMM shape.  The default (4 warps x 2 stages, three CTAs per SM) suits the shapes with many CTAs; the
// N = hidden shapes (o-proj, down-proj: N/16 = 256 CTAs on 148 SMs) leave a third of the shared memory unused with it, so
// they get deeper rings.  vcb_tune_skinny / VCB_SKINNY_TUNE="N:K:warps:depth,..." override the table (tools/bench_decode.py).
namespace {
struct SkinnyTune { int N, K, warps, depth; };
std::vector<SkinnyTune>& tune_table() {
  static std::vector<SkinnyTune> t;
  static bool init = false;
  if (!init) {
    init = true;
    if (const char* e = getenv("VCB_SKINNY_TUNE")) {
      const char* q = e;
      while (*q) {
        SkinnyTune u{0, 0, 0, 0};
        if (sscanf(q, "%d:%d:%d:%d", &u.N, &u.K, &u.warps, &u.depth) == 4) t.push_back(u);
        while (*q && *q != ',') ++q;
        if (*q == ',') ++q;
      }
    }
  }
  return t;
}
}  // namespace

int tune_skinny(int N, int K, int warps, int depth) {
  VCB_CHECK((warps == 0 || warps == 2 || warps == 4 || warps == 8) && depth >= 0 && depth <= 4, kErrInvalid,
            "tune_skinny: warps in {0 (default), 2, 4, 8}, depth in 0..4");
  auto& t = tune_table();
  for (auto& u : t)
    if (u.N == N && u.K == K) { u.warps = warps; u.depth = depth; return kOk; }
  t.push_back({N, K, warps, depth});
  return kOk;
}

int pack_skinny_weights(const __half* W, __half* Wp, int N, int K, cudaStream_t stream) {
  VCB_CHECK(W != nullptr && Wp != nullptr && N % 16 == 0 && K % 64 == 0, kErrInvalid,
            "pack_skinny_weights: N % 16 and K % 64 must be 0");
  VCB_CHECK((reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wp) & 15) == 0, kErrInvalid,
            "pack_skinny_weights: pointers must be 16-byte aligned");
  KernelScope scope(kCatMisc, stream);
  pack_skinny_kernel<<<8 * device_sm_count(), 256, 0, stream>>>(reinterpret_cast<const uint4*>(W), reinterpret_cast<uint4*>(Wp), N, K);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

int skinny_gemm_tma(const SkinnyParams& p_in, cudaStream_t stream) {
  SkinnyParams p = p_in;
  VCB_CHECK(p.B >= 1 && p.B <= 32, kErrInvalid, "skinny_gemm: batch must be in [1, 32]");
  VCB_CHECK(p.N % 16 == 0 && p.K % 64 == 0, kErrInvalid, "skinny_gemm_tma: N % 16 and K % 64 must be 0");
  const int grid = p.N / 16;
  const int nb = (p.B + 7) / 8;
  {
    static int env_pf = -1, env_pc = -1;  // VCB_SKINNY_PF: k-groups per warp prefetched into L2 before the dependency wait
    if (env_pf < 0) { const char* e = getenv("VCB_SKINNY_PF"); env_pf = e ? atoi(e) : kPrefetchGroupsDefault; }
    if (env_pc < 0) { const char* e = getenv("VCB_SKINNY_PF_CTAS"); env_pc = e ? atoi(e) : 3 * device_sm_count(); }
    p.pf_groups = env_pf;
    p.pf_ctas = env_pc;
  }
  if (p.Wp != nullptr) {
    VCB_CHECK((reinterpret_cast<uintptr_t>(p.Wp) & 15) == 0, kErrInvalid, "skinny_gemm: packed weights must be 16-byte aligned");
    static int env_pd = -1;
    if (env_pd < 0) { const char* e = getenv("VCB_SKINNY_TMA_DEPTH"); env_pd = e ? atoi(e) : 0; }
    p.trace = trace_slots(kTraceSkinny, grid, p.N);
    KernelScope scope(kCatSkinny, stream);
    switch (nb) {
      case 1: return launch_packed<1>(p, grid, env_pd, stream);
      case 2: return launch_packed<2>(p, grid, env_pd, stream);
      case 3: return launch_packed<3>(p, grid, env_pd, stream);
      default: return launch_packed<4>(p, grid, env_pd, stream);
    }
  }
  int warps = 4;  // 4 warps x 3 stages x 8 KiB = 96 KiB per CTA -> two CTAs per SM
  int depth = 2, group = 4;  // measured best on B200: 4 warps x 2 stages x 8 KiB -> three CTAs per SM
  static int env_w = -1, env_d = -1, env_g = -1;
  if (env_g < 0) { const char* e = getenv("VCB_SKINNY_TMA_GROUP"); env_g = e ? atoi(e) : 0; }
  if (env_g == 2 || env_g == 4) group = env_g;
  if (env_w < 0) { const char* e = getenv("VCB_SKINNY_TMA_WARPS"); env_w = e ? atoi(e) : 0; }
  if (env_d < 0) { const char* e = getenv("VCB_SKINNY_TMA_DEPTH"); env_d = e ? atoi(e) : 0; }
  if (env_w == 2 || env_w == 4 || env_w == 8) warps = env_w;
  if (env_d > 0) depth = env_d;
  for (const auto& u : tune_table())
    if (u.N == p.N && u.K == p.K) {
      if (u.warps > 0) warps = u.warps;
      if (u.depth > 0) depth = u.depth;
    }
  if (p.pn_parts != nullptr || p.cn_rstd != nullptr) { group = 4; warps = 4; depth = 2; }
  CUtensorMap tm;
  VCB_TRY(make_tmap_w3d(&tm, p.W, p.K, p.N, group));
  p.trace = trace_slots(kTraceSkinny, grid, p.N);
  KernelScope scope(kCatSkinny, stream);
  switch (nb) {
    case 1: return launch_tma_w<1>(tm, p, grid, warps, depth, group, stream);
    case 2: return launch_tma_w<2>(tm, p, grid, warps, depth, group, stream);
    case 3: return launch_tma_w<3>(tm, p, grid, warps, depth, group, stream);
    default: return launch_tma_w<4>(tm, p, grid, warps, depth, group, stream);
  }
}

}  // namespace vcb
