// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] · W[N,K]^T ),  fp16 in, fp32 accumulate in TMEM.
//
// Replaces every dense nn.Linear / Conv2d-as-GEMM on the VCoder hot path that the reference reaches through
// HF transformers -> cuBLAS/cuDNN (SURVEY.md §2.2 K1,K3,K5,K6,K9,K12):
//   CLIP patch-embed conv  (transformers/models/clip/modeling_clip.py:150-156 -> im2col + this kernel)
//   CLIP q/k/v/out proj, fc1 (+quick_gelu), fc2 (+residual)          (modeling_clip.py:290-345)
//   mm_projector / seg_mm_projector  Linear+GELU(erf)+Linear          (reference vcoder_llava/model/multimodal_projector/builder.py:33-46)
//   LLaMA q/k/v/o proj, gate/up (+SwiGLU), down (+residual) at prefill (transformers/models/llama/modeling_llama.py:171-184,224-)
//
// Structure (persistent, warp-specialised, one CTA per SM, 192 threads):
//   warp 0 / lane 0 : TMA producer  — cp.async.bulk.tensor 128B-swizzled K-major tiles of A (128x64) and W (BNx64)
//   warp 1 / lane 0 : MMA issuer    — 4 x tcgen05.mma (128 x BN x 16) per k-block, accumulators double-buffered in TMEM
//   warps 2..9      : epilogue      — tcgen05.ld TMEM -> registers, bias / activation / residual, fp16 stores
// Three mbarrier pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), static tile round-robin.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <map>
#include <tuple>

#include "common.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 bytes = one swizzle-128B row
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant, half of the tile's columns each)
constexpr int kRasterGroup1 = 16;  // tile rows per band, 1-CTA kernel: minimises 16 x 1 MB (A) + 148/16 x 2 MB (W) per wave at K = 4096
constexpr int kRasterGroup2 = 8;   // CTA-pair kernel: 8 x 2 MB + 74/8 x 2 MB
constexpr int kGemmPairDefault = 1;  // VCB_GEMM_PAIR=0 forces the 1-CTA kernel (A/B: DESIGN.md section 3, profiles/r01_gemm_pair_s3.log)

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual 1 KiB alignment
  static constexpr int kTmemCols = 2 * BN;                                     // two accumulator buffers
};

// x * sigmoid(a x) = x / (1 + 2^(-a x log2e)): one ex2.approx + one rcp.approx (MUFU) instead of a full-precision divide
__device__ __forceinline__ float act_quick_gelu(float x) { return __fdividef(x, 1.0f + fast_exp2(-2.4554669595930157f * x)); }
__device__ __forceinline__ float act_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float act_silu(float x) { return __fdividef(x, 1.0f + fast_exp2(-1.4426950408889634f * x)); }

__device__ __forceinline__ void add_half8(float* v, const uint4& h) {
  float2 a = unpack_half2(h.x), b = unpack_half2(h.y), c = unpack_half2(h.z), d = unpack_half2(h.w);
  v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
}

// Tile rasterisation: tiles are walked in bands of kGroup tile rows, column by column inside a band, so the CTAs that run
// concurrently (one wave = 148 tiles or 74 pair-tiles) cover a squarish block of the output and share A row-blocks AND W
// column-blocks in L2.  Row-major order made every wave stream its own ~150 MB of W through the 126 MB L2 (ncu: 6.7 GB of
// DRAM reads for the 9656 x 22016 x 4096 gate|up GEMM against 0.26 GB of operands).
template <int kGroup>
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int* tm, int* tn) {
  const int band = kGroup * tiles_n;
  const int first_m = (tile / band) * kGroup;
  const int rows = min(tiles_m - first_m, kGroup);
  const int in_band = tile % band;
  *tm = first_m + in_band % rows;
  *tn = in_band / rows;
}

// Epilogue of one accumulator row: `tmem_acc` = TMEM address of this thread's lane at the first column of the accumulator
// buffer, m = global output row, n_base = first global column of the tile.  tcgen05.ld -> bias / activation / residual ->
// fp16 stores (a thread owns one row, 32 columns at a time).
// [c_begin, c_end): the columns of the tile this warp handles.  Two warps share a lane quadrant: with a 16-k-block mainloop
// (the K = 1024 tower GEMMs) four epilogue warps needed ~12.7 k cycles per 128 x 256 tile against 8.2 k of MMAs and held the
// tensor pipe at 64 % (profiles/r02_ncu_gemm_tower_vs_prefill_s2.csv).
template <int BN, int ACT>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpilogue& ep, uint32_t tmem_acc, int m, int M, int N, int n_base,
                                                   int c_begin, int c_end) {
  const bool row_ok = m < M;
  if (ACT == kActRopeKv) {
    // q | k | v of one token row, one 128-column head at a time: columns d and d + 64 of a head rotate together
    // (apply_rotary_pos_emb, modeling_llama.py:124-168; same arithmetic on the fp16-rounded values as rope_kv_kernel)
    int page = 0, slot = 0;
    const float2* csr = ep.rk_cos_sin;
    if (row_ok) {
      const int b = m / ep.rk_T, pos = ep.rk_pos0 + (m - b * ep.rk_T);
      page = ep.rk_block_table[b * ep.rk_max_pages + pos / ep.rk_page_tokens];
      slot = pos % ep.rk_page_tokens;
      csr += static_cast<long>(pos) * 64;
    }
#pragma unroll 1
    for (int hb = c_begin; hb < c_end; hb += 128) {
      const int n0 = n_base + hb;
      if (n0 >= N) break;
      const int sect = n0 / ep.rk_H;  // 0 = q, 1 = k, 2 = v (H is a multiple of the tile width)
      const int head = (n0 - sect * ep.rk_H) >> 7;
      __half* dst;
      if (sect == 0) dst = ep.C + static_cast<long>(m) * ep.ldc + n0;
      else dst = (sect == 1 ? ep.rk_k_pool : ep.rk_v_pool) +
                 ((static_cast<long>(page) * ep.rk_heads + head) * ep.rk_page_tokens + slot) * 128;
#pragma unroll 1
      for (int cc = 0; cc < 64; cc += 32) {
        uint32_t lo[32], hi[32];
        tmem_ld_32x32(tmem_acc + hb + cc, lo);
        tmem_ld_32x32(tmem_acc + hb + 64 + cc, hi);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float o1[8], o2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float x1 = __half2float(__float2half_rn(__uint_as_float(lo[j + e])));
              const float x2 = __half2float(__float2half_rn(__uint_as_float(hi[j + e])));
              if (sect < 2) {
                const float2 cs = __ldg(csr + cc + j + e);
                o1[e] = x1 * cs.x - x2 * cs.y;
                o2[e] = x2 * cs.x + x1 * cs.y;
              } else {
                o1[e] = x1;
                o2[e] = x2;
              }
            }
            *reinterpret_cast<uint4*>(dst + cc + j) =
                make_uint4(pack_half2(o1[0], o1[1]), pack_half2(o1[2], o1[3]), pack_half2(o1[4], o1[5]), pack_half2(o1[6], o1[7]));
            *reinterpret_cast<uint4*>(dst + 64 + cc + j) =
                make_uint4(pack_half2(o2[0], o2[1]), pack_half2(o2[2], o2[3]), pack_half2(o2[4], o2[5]), pack_half2(o2[6], o2[7]));
          }
        }
      }
    }
    return;
  }
  long out_row = m;
  if (ep.out_group) out_row = m + static_cast<long>(m / ep.out_group) * ep.out_skip + ep.out_off;
  const long res_row = ep.res_mod ? static_cast<long>(m % ep.res_mod) + ep.res_off : out_row;
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    const int n0 = n_base + c;
    if (n0 >= N) break;
    uint32_t r[32];
    tmem_ld_32x32(tmem_acc + c, r);
    tmem_ld_wait();
    if (row_ok) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.bias) {
        const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) add_half8(v + 8 * j, __ldg(bp + j));
      }
      if (ACT == kActQuickGelu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_quick_gelu(v[j]);
      } else if (ACT == kActGeluErf) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_gelu_erf(v[j]);
      }
      if (ACT == kActSwiGLU) {
        // columns [16g, 16g+8) = gate, [16g+8, 16g+16) = up  ->  8 outputs per granule
        uint32_t o[8];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float a = act_silu(v[16 * g + e]) * v[16 * g + 8 + e];
            const float b = act_silu(v[16 * g + e + 1]) * v[16 * g + 8 + e + 1];
            o[4 * g + e / 2] = pack_half2(a, b);
          }
        }
        uint4* cp = reinterpret_cast<uint4*>(ep.C + out_row * ep.ldc + (n0 >> 1));
        cp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        cp[1] = make_uint4(o[4], o[5], o[6], o[7]);
      } else {
        if (ep.residual) {
          const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + res_row * ep.ldr + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) add_half8(v + 8 * j, rp[j]);
        }
        uint4* cp = reinterpret_cast<uint4*>(ep.C + out_row * ep.ldc + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cp[j] = make_uint4(pack_half2(v[8 * j + 0], v[8 * j + 1]), pack_half2(v[8 * j + 2], v[8 * j + 3]),
                             pack_half2(v[8 * j + 4], v[8 * j + 5]), pack_half2(v[8 * j + 6], v[8 * j + 7]));
        }
      }
    }
  }
}

template <int BN, int ACT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpilogue ep,
                int M, int N, int K, int tiles_m, int tiles_n) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  const uint32_t bars_addr = tiles_addr + Cfg::kStages * Cfg::kStageBytes;
  // barrier index map: [0,S) full, [S,2S) empty, 2S+{0,1} tmem_full, 2S+{2,3} tmem_empty
  auto bar = [&](int i) -> uint32_t { return bars_addr + 8u * static_cast<uint32_t>(i); };
  constexpr int S = Cfg::kStages;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * S; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(2 * S + 0), 1);
    mbar_init(bar(2 * S + 1), 1);
    mbar_init(bar(2 * S + 2), 8);
    mbar_init(bar(2 * S + 3), 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = K / BK;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tm, tn;
        tile_coords<kRasterGroup1>(tile, tiles_m, tiles_n, &tm, &tn);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar(S + s), phase ^ 1u);
          const uint32_t a_dst = tiles_addr + s * Cfg::kStageBytes;
          mbar_expect_tx(bar(s), Cfg::kStageBytes);
          tma_load_2d(a_dst, &tmA, bar(s), kb * BK, tm * BM);
          tma_load_2d(a_dst + Cfg::kABytes, &tmB, bar(s), kb * BK, tn * BN);
          if (++s == S) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN);
      int s = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(bar(2 * S + 2 + acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar(s), phase);
          tc_fence_after();
          const uint32_t a_addr = tiles_addr + s * Cfg::kStageBytes;
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the 128B swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(bar(S + s));  // frees the smem stage once these MMAs have read it
          if (++s == S) { s = 0; phase ^= 1u; }
        }
        umma_commit(bar(2 * S + acc));  // accumulator complete -> epilogue
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 2..9
    const int q = warp & 3;  // a warp may only touch TMEM lanes [32*(warpid%4), +32)
    // column half of this warp; the RoPE epilogue keeps the 128 columns of a head together (BN = 128: one warp per quadrant)
    const int half = (warp - 2) >> 2;
    const int c_begin = (ACT == kActRopeKv && BN == 128) ? (half == 0 ? 0 : BN) : half * (BN / 2);
    const int c_end = (ACT == kActRopeKv && BN == 128) ? BN : (half + 1) * (BN / 2);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tm, tn;
      tile_coords<kRasterGroup1>(tile, tiles_m, tiles_n, &tm, &tn);
      mbar_wait(bar(2 * S + acc), acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<BN, ACT>(ep, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, tm * BM + q * 32 + lane, M, N,
                                  tn * BN, c_begin, c_end);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(2 * S + 2 + acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// -------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per pair of SMs.  Each CTA stages ITS 128 rows of A and ITS
// 128 of the tile's 256 W rows (32 KiB per k-block instead of 48), the leader's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory and writes each CTA's own TMEM half, and every
// CTA runs the epilogue of its own 128 rows.  Per SM and k-block the shared-memory traffic drops from 192 B/clk (TMA writes
// + MMA reads of the 1-CTA 128 x 256 tile) to 128 B/clk, which is what bounds the 1-CTA kernel.
//   full[s]      (leader only, count 1): leader's producer arrives with expect_tx of BOTH CTAs' stage bytes; both CTAs' TMA
//                loads complete_tx on it (tma_load_2d_2sm)
//   empty[s]     (each CTA, count 1): tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[a] (each CTA, count 1): commit multicast after the tile's last k-block
//   tmem_empty[a](leader only, count 16): the eight epilogue warps of each CTA (the peer's arrive through DSMEM)
// -------------------------------------------------------------------------------------------------
struct PairCfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = 128 * BK * 2;  // this CTA's half of the 256 W rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
  static constexpr int kTmemCols = 512;  // two 256-column accumulator buffers
};

template <int ACT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc5_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpilogue ep,
                     int M, int N, int K, int tiles_m, int tiles_n) {
  using Cfg = PairCfg;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  const uint32_t bars_addr = tiles_addr + Cfg::kStages * Cfg::kStageBytes;
  auto bar = [&](int i) -> uint32_t { return bars_addr + 8u * static_cast<uint32_t>(i); };
  constexpr int S = Cfg::kStages;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * S; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(2 * S + 0), 1);
    mbar_init(bar(2 * S + 1), 1);
    mbar_init(bar(2 * S + 2), 16);
    mbar_init(bar(2 * S + 3), 16);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = K / BK;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int s = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        int tm, tn;
        tile_coords<kRasterGroup2>(tile, tiles_m, tiles_n, &tm, &tn);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar(S + s), phase ^ 1u);
          const uint32_t a_dst = tiles_addr + s * Cfg::kStageBytes;
          if (rank == 0) mbar_expect_tx(bar(s), 2 * Cfg::kStageBytes);
          tma_load_2d_2sm(a_dst, &tmA, bar(s), kb * BK, tm * 256 + static_cast<int>(rank) * BM);
          tma_load_2d_2sm(a_dst + Cfg::kABytes, &tmB, bar(s), kb * BK, tn * BN + static_cast<int>(rank) * 128);
          if (++s == S) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN);
      int s = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(bar(2 * S + 2 + acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar(s), phase);
          tc_fence_after();
          const uint32_t a_addr = tiles_addr + s * Cfg::kStageBytes;
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(bar(S + s), 0b11);  // frees the stage in BOTH CTAs
          if (++s == S) { s = 0; phase ^= 1u; }
        }
        umma_commit_2sm(bar(2 * S + acc), 0b11);  // accumulator complete -> both epilogues
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // two epilogue warps per lane quadrant: 128 of the 256 columns each (one head for RoPE)
    const int c_begin = half * (BN / 2), c_end = (half + 1) * (BN / 2);
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tm, tn;
      tile_coords<kRasterGroup2>(tile, tiles_m, tiles_n, &tm, &tn);
      mbar_wait(bar(2 * S + acc), acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<BN, ACT>(ep, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN,
                                  tm * 256 + static_cast<int>(rank) * BM + q * 32 + lane, M, N, tn * BN, c_begin, c_end);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar(2 * S + 2 + acc));
        else mbar_arrive_cluster(dsmem_addr(bar(2 * S + 2 + acc), 0));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees TMEM or exits while the pair's MMAs / remote arrivals may still touch it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// Descriptor cache: a prefill issues ~450 GEMMs over a few dozen distinct (pointer, shape) pairs — workspaces and weights
// live at fixed addresses for the life of a handle — so every tensor map is encoded by the driver once and then reused.
struct TmapKey {
  const void* ptr;
  uint64_t a, b, c;
  uint32_t box0, box1, kind;
  bool operator<(const TmapKey& o) const {
    return std::tie(ptr, a, b, c, box0, box1, kind) < std::tie(o.ptr, o.a, o.b, o.c, o.box0, o.box1, o.kind);
  }
};
std::map<TmapKey, CUtensorMap>& tmap_cache() {
  static std::map<TmapKey, CUtensorMap> c;
  return c;
}
std::mutex& tmap_mu() {
  static std::mutex m;
  return m;
}
bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(tmap_mu());
  auto it = tmap_cache().find(k);
  if (it == tmap_cache().end()) return false;
  *out = it->second;
  return true;
}
void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(tmap_mu());
  if (tmap_cache().size() > 8192) tmap_cache().clear();  // op-level callers with ever-new buffers
  tmap_cache()[k] = m;
}

int make_tmap_kmajor(CUtensorMap* m, const void* ptr, uint64_t k, uint64_t rows, uint64_t ld_elems,
                     uint32_t box_rows) {
  const TmapKey key{ptr, k, rows, ld_elems, BK, box_rows, 0};
  if (tmap_lookup(key, m)) return kOk;
  EncodeTiledFn fn = get_encode_fn();
  VCB_CHECK(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_thread_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return kErrCuda;
  }
  tmap_store(key, *m);
  return kOk;
}

template <int BN, int ACT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int N, int K,
           cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    VCB_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel<BN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int grid = std::min(tiles_m * tiles_n, device_sm_count());
  KernelScope scope(kCatGemm, stream);
  gemm_tc5_kernel<BN, ACT><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, ep, M, N, K, tiles_m, tiles_n);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

template <int BN>
int launch_act(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int N, int K,
               cudaStream_t stream) {
  switch (ep.act) {
    case kActNone: return launch<BN, kActNone>(tmA, tmB, ep, M, N, K, stream);
    case kActQuickGelu: return launch<BN, kActQuickGelu>(tmA, tmB, ep, M, N, K, stream);
    case kActGeluErf: return launch<BN, kActGeluErf>(tmA, tmB, ep, M, N, K, stream);
    case kActSwiGLU: return launch<BN, kActSwiGLU>(tmA, tmB, ep, M, N, K, stream);
    case kActRopeKv: return launch<BN, kActRopeKv>(tmA, tmB, ep, M, N, K, stream);
    default: set_thread_error("gemm_f16: unknown activation"); return kErrInvalid;
  }
}


template <int ACT>
int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int N, int K,
                cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VCB_CUDA(cudaFuncSetAttribute(gemm_tc5_pair_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles_m = (M + 255) / 256;
  const int tiles_n = N / 256;
  const int pairs = std::min(tiles_m * tiles_n, device_sm_count() / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = PairCfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  KernelScope scope(kCatGemm, stream);
  VCB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc5_pair_kernel<ACT>, tmA, tmB, ep, M, N, K, tiles_m, tiles_n));
  return kOk;
}

int launch_pair_act(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int N, int K,
                    cudaStream_t stream) {
  switch (ep.act) {
    case kActNone: return launch_pair<kActNone>(tmA, tmB, ep, M, N, K, stream);
    case kActQuickGelu: return launch_pair<kActQuickGelu>(tmA, tmB, ep, M, N, K, stream);
    case kActGeluErf: return launch_pair<kActGeluErf>(tmA, tmB, ep, M, N, K, stream);
    case kActSwiGLU: return launch_pair<kActSwiGLU>(tmA, tmB, ep, M, N, K, stream);
    case kActRopeKv: return launch_pair<kActRopeKv>(tmA, tmB, ep, M, N, K, stream);
    default: set_thread_error("gemm_f16: unknown activation"); return kErrInvalid;
  }
}

}  // namespace

// generic 2-D fp16 tensor map with the 128-byte swizzle (box_inner must be 64 elements = 128 bytes)
int make_tmap_2d_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
                     uint32_t box_rows) {
  const TmapKey key{ptr, inner, rows, ld_elems, box_inner, box_rows, 1};
  if (tmap_lookup(key, m)) return kOk;
  EncodeTiledFn fn = get_encode_fn();
  VCB_CHECK(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_thread_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return kErrCuda;
  }
  tmap_store(key, *m);
  return kOk;
}

// KV page pool [pages][heads][64 tokens][128] viewed as [n_rows = pages*heads][64][128]: box = one 64-wide head_dim half
// of one page of one head (64 tokens x 128 B = 8 KiB, 128B-swizzled) for the tcgen05 prefill attention
int make_tmap_pages_f16(CUtensorMap* m, const void* ptr, uint64_t n_rows) {
  const TmapKey key{ptr, n_rows, 0, 0, 64, 64, 2};
  if (tmap_lookup(key, m)) return kOk;
  EncodeTiledFn fn = get_encode_fn();
  VCB_CHECK(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {128, 64, n_rows};
  cuuint64_t strides[2] = {256, 64 * 256};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_thread_error("cuTensorMapEncodeTiled (KV page map) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return kErrCuda;
  }
  tmap_store(key, *m);
  return kOk;
}

// W [N, K] row-major viewed as [K/64][N][64] for the TMA-fed skinny GEMM: box = 4 k-blocks x 16 rows x 64 k (8 KiB)
int make_tmap_w3d(CUtensorMap* m, const void* ptr, uint64_t K, uint64_t N, uint32_t box_kblocks) {
  const TmapKey key{ptr, K, N, 0, box_kblocks, 16, 3};
  if (tmap_lookup(key, m)) return kOk;
  EncodeTiledFn fn = get_encode_fn();
  VCB_CHECK(fn != nullptr, kErrCuda, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {64, N, K / 64};
  cuuint64_t strides[2] = {K * 2, 128};
  cuuint32_t box[3] = {64, 16, box_kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_thread_error("cuTensorMapEncodeTiled (3-D weight map) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return kErrCuda;
  }
  tmap_store(key, *m);
  return kOk;
}

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0) {
      n = 148;
    }
  }
  return n;
}

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
             int block_n, cudaStream_t stream) {
  VCB_CHECK(M > 0 && N > 0 && K > 0, kErrInvalid, "gemm_f16: empty problem");
  VCB_CHECK(K % BK == 0, kErrInvalid, "gemm_f16: K must be a multiple of 64");
  VCB_CHECK(N % 32 == 0, kErrInvalid, "gemm_f16: N must be a multiple of 32");
  VCB_CHECK(lda % 8 == 0 && ldw % 8 == 0, kErrInvalid, "gemm_f16: leading dimensions must be multiples of 8");
  VCB_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(ep.C) & 15) == 0,
            kErrInvalid, "gemm_f16: pointers must be 16-byte aligned");
  VCB_CHECK(ep.C != nullptr && ep.ldc % 8 == 0, kErrInvalid, "gemm_f16: bad output");
  if (ep.act == kActRopeKv) {
    VCB_CHECK(ep.rk_cos_sin && ep.rk_k_pool && ep.rk_v_pool && ep.rk_block_table && ep.rk_T > 0 && ep.rk_page_tokens > 0 &&
                  ep.rk_H == ep.rk_heads * 128 && N == 3 * ep.rk_H && M % ep.rk_T == 0,
              kErrInvalid, "gemm_f16: incomplete RoPE / KV-append epilogue description");
    VCB_CHECK(ep.bias == nullptr && ep.residual == nullptr && ep.out_group == 0, kErrInvalid,
              "gemm_f16: the RoPE / KV-append epilogue takes no bias, residual or row remap");
    if (ep.rk_H % 256 != 0) block_n = 128;  // a tile must not straddle q | k | v
  }
  {
    // block_n == 512 or VCB_GEMM_PAIR=1: CTA-pair kernel (256 x 256 tile per two SMs) where the shape allows it
    static int env_pair = -1;
    if (env_pair < 0) { const char* e = getenv("VCB_GEMM_PAIR"); env_pair = e ? atoi(e) : kGemmPairDefault; }
    // measured on B200 (profiles/r01_gemm_pair_s3.log, r01_gemm_raster_s3.log): the pair kernel is 5..10 % faster per tile on
    // the large shapes, 1-CTA wins on the K = 1024 GEMMs with an activation epilogue and whenever the 256 x 256 tiling
    // quantises into a worse last wave (e.g. M = 4616: 304 pair-tiles on 74 pairs = 5 rounds vs 592 tiles on 148 SMs = 4)
    const bool short_k_act = K < 2048 && ep.act != kActNone;
    bool auto_pair = false;
    if (block_n == 0 && env_pair != 0 && M >= 1024 && N % 256 == 0 && !short_k_act) {
      const int sms = device_sm_count();
      const long t1 = static_cast<long>((M + BM - 1) / BM) * (N / 256), t2 = static_cast<long>((M + 255) / 256) * (N / 256);
      const double eff1 = static_cast<double>(t1) / (((t1 + sms - 1) / sms) * sms);
      const double eff2 = static_cast<double>(t2) / (((t2 + sms / 2 - 1) / (sms / 2)) * (sms / 2));
      auto_pair = eff2 * 1.07 >= eff1;
    }
    const bool want_pair = block_n == 512 || auto_pair;
    if (want_pair && N % 256 == 0) {
      CUtensorMap tmA, tmB;
      VCB_TRY(make_tmap_kmajor(&tmA, A, K, M, lda, BM));
      VCB_TRY(make_tmap_kmajor(&tmB, W, K, N, ldw, 128));
      return launch_pair_act(tmA, tmB, ep, M, N, K, stream);
    }
    VCB_CHECK(block_n != 512, kErrInvalid, "gemm_f16: the CTA-pair kernel needs N % 256 == 0");
  }
  int bn = block_n;
  if (bn == 0) {
    // 256-wide tiles halve the A re-reads; keep 128 when that would leave SMs idle.
    const long tiles256 = static_cast<long>((M + BM - 1) / BM) * ((N + 255) / 256);
    bn = (N % 256 == 0 && tiles256 >= device_sm_count()) ? 256 : 128;
  }
  VCB_CHECK(bn == 128 || bn == 256, kErrInvalid, "gemm_f16: block_n must be 0, 128 or 256");
  CUtensorMap tmA, tmB;
  VCB_TRY(make_tmap_kmajor(&tmA, A, K, M, lda, BM));
  VCB_TRY(make_tmap_kmajor(&tmB, W, K, N, ldw, bn));
  return bn == 256 ? launch_act<256>(tmA, tmB, ep, M, N, K, stream) : launch_act<128>(tmA, tmB, ep, M, N, K, stream);
}

}  // namespace vcb
