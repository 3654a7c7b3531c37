// tcgen05 weight-streaming GEMM for decode / lm_head:  Y[B,N] = X[B,K] · W[N,K]^T,  B <= 32.
// Replaces the per-token nn.Linear calls of HF LlamaModel at decode (modeling_llama.py:171-184,224-) and the lm_head
// (reference vcoder_ds_llava_llama.py:92-93); HBM-bound: each weight byte is read once.
//
// The weights are the *M* operand: a CTA streams a 128-row x K-slice panel of W through a TMA ring (128B-swizzled
// [128 x 64] tiles, the same layout the big GEMM uses for its A operand) and a single thread issues
// tcgen05.mma 128 x NT x 16 (NT = 16 or 32 = padded batch) against the matching [NT x 64] tile of X, accumulating
// in TMEM.  The CUDA cores never touch a weight byte — the consumer side of the stream costs 4 instructions per 16 KiB —
// so the bytes in flight (ring depth x CTAs per SM) are what bounds the kernel, not instruction issue.
// K is split across the CTAs of a thread-block cluster (<= 8); the partial accumulators are reduced through
// distributed shared memory in rank order (deterministic), and rank 0 applies the epilogue (SwiGLU / residual) and stores.
// Programmatic dependent launch: the W tiles of the first ring fill are requested BEFORE griddepcontrol.wait.
// Built only with VCB_EXPERIMENTAL=1 (python -m vcoder_b200.build): measured slower than the TMA-fed kernels.
#ifdef VCB_EXPERIMENTAL
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kBM = 128;   // weight rows per CTA
constexpr int kBK = 64;    // k per stage (128-byte swizzle row)
constexpr int kStages = 4;
constexpr int kThreads = 128;
constexpr int kWBytes = kBM * kBK * 2;  // 16 KiB

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + fast_exp2(-1.4426950408889634f * x)); }

template <int NT>
__global__ void __launch_bounds__(kThreads)
skinny_tc5_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, SkinnyParams p,
                  int ksplit) {
  constexpr int kXBytes = NT * kBK * 2;
  constexpr int kStageBytes = kWBytes + kXBytes;
  extern __shared__ __align__(1024) uint8_t smem[];  // ring | barriers ; the ring doubles as the fp32 partial buffer
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t bars = base + kStages * kStageBytes;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * static_cast<uint32_t>(i); };  // [0,S) full, [S,2S) empty, 2S done
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kStages * kStageBytes + 8 * (2 * kStages + 1));

  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kBM;
  const int rank = ksplit > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int nkb_all = p.K / kBK;
  const int per = (nkb_all + ksplit - 1) / ksplit;
  const int kb0 = min(nkb_all, rank * per), kb1 = min(nkb_all, kb0 + per);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kStages + 1; ++i) mbar_init(bar(i), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmW);
      tma_prefetch_desc(&tmX);
      // first ring fill: weights before the dependency wait, activations after it
      const int pre = min(kStages, nkb);
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(bar(i), kStageBytes);
        tma_load_2d(base + i * kStageBytes, &tmW, bar(i), (kb0 + i) * kBK, row0);
      }
      griddep_wait();
      for (int i = 0; i < pre; ++i) tma_load_2d(base + i * kStageBytes + kWBytes, &tmX, bar(i), (kb0 + i) * kBK, 0);
      int s = pre % kStages;
      uint32_t phase = pre == kStages ? 1u : 0u;  // parity of the ring pass the producer is in
      for (int i = pre; i < nkb; ++i) {
        mbar_wait(bar(kStages + s), phase ^ 1u);
        mbar_expect_tx(bar(s), kStageBytes);
        tma_load_2d(base + s * kStageBytes, &tmW, bar(s), (kb0 + i) * kBK, row0);
        tma_load_2d(base + s * kStageBytes + kWBytes, &tmX, bar(s), (kb0 + i) * kBK, 0);
        if (++s == kStages) { s = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBM, NT);
      int s = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar(s), phase);
        tc_fence_after();
        const uint64_t adesc = make_sw128_kmajor_desc(base + s * kStageBytes);
        const uint64_t bdesc = make_sw128_kmajor_desc(base + s * kStageBytes + kWBytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
        umma_commit(bar(kStages + s));
        if (++s == kStages) { s = 0; phase ^= 1u; }
      }
      umma_commit(bar(2 * kStages));  // accumulator complete
    }
  }
  if (!(warp == 0 && lane == 0)) griddep_wait();  // everybody is past the dependency before touching outputs
  __syncwarp();

  // ---- all four warps: accumulator row (TMEM lane) -> registers
  float v[NT];
  if (nkb > 0) {
    mbar_wait(bar(2 * kStages), 0);
    tc_fence_after();
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    if (NT == 16) {
      uint32_t r[16];
      tmem_ld_32x32_x16(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < 16; ++n) v[n] = __uint_as_float(r[n]);
    } else {
      uint32_t r[32];
      tmem_ld_32x32(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < NT; ++n) v[n] = __uint_as_float(r[n < 32 ? n : 0]);
    }
  } else {
#pragma unroll
    for (int n = 0; n < NT; ++n) v[n] = 0.f;
  }
  const int rloc = warp * 32 + lane;  // row within the 128-row panel

  if (ksplit > 1) {
    // split-K reduction through distributed shared memory, fixed rank order
    __syncthreads();  // the ring is no longer read by the tensor core (done barrier observed by every thread)
    float* part = reinterpret_cast<float*>(smem);  // [128][NT]
#pragma unroll
    for (int n = 0; n < NT; ++n) part[rloc * NT + n] = v[n];
    cluster_sync_all();
    if (rank == 0) {
#pragma unroll
      for (int n = 0; n < NT; ++n) v[n] = 0.f;
      for (int r = 0; r < ksplit; ++r) {
        const uint32_t remote = dsmem_addr(base + static_cast<uint32_t>(rloc * NT) * 4u, static_cast<uint32_t>(r));
#pragma unroll
        for (int n = 0; n < NT; ++n)
          if (n < p.B) v[n] += ld_dsmem_f32(remote + 4u * n);
      }
    }
  }

  if (rank == 0) {
    const int row = row0 + rloc;
    if (p.swiglu) {
      // rows are interleaved 8 gate | 8 up inside every 16-row group: the partner sits 8 lanes away
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const float up = __shfl_down_sync(0xffffffffu, v[n], 8);
        if (n < p.B && (rloc & 15) < 8)
          p.Y[static_cast<long>(n) * p.ldy + (row >> 4) * 8 + (rloc & 7)] = __float2half_rn(silu_f(v[n]) * up);
      }
    } else {
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < p.B) {
          float y = v[n];
          if (p.residual) y += __half2float(p.residual[static_cast<long>(n) * p.ldr + row]);
          p.Y[static_cast<long>(n) * p.ldy + row] = __float2half_rn(y);
        }
      }
    }
  }
  if (ksplit > 1) cluster_sync_all();  // partials must stay valid until rank 0 has read them

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

template <int NT>
int launch_tc5(const CUtensorMap& tmW, const CUtensorMap& tmX, const SkinnyParams& p, int tiles, int ksplit,
               cudaStream_t stream) {
  constexpr int smem = kStages * (kWBytes + NT * kBK * 2) + 256;
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(skinny_tc5_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, ksplit);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 1;
  at[1].val.clusterDim.y = ksplit;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  VCB_CUDA(cudaLaunchKernelEx(&cfg, skinny_tc5_kernel<NT>, tmW, tmX, p, ksplit));
  return kOk;
}

}  // namespace

int make_tmap_2d_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
                     uint32_t box_rows);

bool skinny_tc5_supported(const SkinnyParams& p) {
  return p.N % kBM == 0 && p.K % kBK == 0 && p.norm_w == nullptr && p.B >= 1 && p.B <= 32 && p.ldx % 8 == 0;
}

int skinny_gemm_tc5(const SkinnyParams& p, cudaStream_t stream) {
  VCB_CHECK(skinny_tc5_supported(p), kErrInvalid, "skinny_gemm_tc5: needs N % 128 == 0, K % 64 == 0, B <= 32, no fused norm");
  const int tiles = p.N / kBM;
  const int nkb = p.K / kBK;
  int ksplit = 1;
  while (ksplit < 8 && tiles * ksplit < 2 * device_sm_count() && nkb / (ksplit * 2) >= 2) ksplit *= 2;
  static int env_ks = -1;
  if (env_ks < 0) { const char* e = getenv("VCB_SKINNY_TC5_KSPLIT"); env_ks = e ? atoi(e) : 0; }
  if (env_ks == 1 || env_ks == 2 || env_ks == 4 || env_ks == 8) ksplit = env_ks;
  const int nt = p.B <= 16 ? 16 : 32;
  CUtensorMap tmW, tmX;
  VCB_TRY(make_tmap_2d_f16(&tmW, p.W, p.K, p.N, p.K, kBK, kBM));
  VCB_TRY(make_tmap_2d_f16(&tmX, p.X, p.K, p.B, p.ldx, kBK, nt));
  KernelScope scope(kCatSkinny, stream);
  return nt == 16 ? launch_tc5<16>(tmW, tmX, p, tiles, ksplit, stream) : launch_tc5<32>(tmW, tmX, p, tiles, ksplit, stream);
}

}  // namespace vcb

#else  // !VCB_EXPERIMENTAL
#include "common.h"
#include "kernels.h"
namespace vcb {
bool skinny_tc5_supported(const SkinnyParams&) { return false; }
int skinny_gemm_tc5(const SkinnyParams&, cudaStream_t) {
  set_thread_error("skinny_gemm_tc5: built without VCB_EXPERIMENTAL");
  return kErrInvalid;
}
}  // namespace vcb
#endif
