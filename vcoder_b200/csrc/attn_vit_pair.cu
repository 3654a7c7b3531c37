// CLIP ViT self-attention on tcgen05 + TMEM, "pair" schedule:  O = softmax(Q K^T * scale) V,  head_dim 64.
// Reference: CLIPVisionTower.forward -> transformers/models/clip/modeling_clip.py:261-345 (eager attention, no mask);
// SURVEY K4 (4,726,784 algorithmic bytes, 1.364 GFLOP per image-layer at S=577).
//
// One CTA per SM works on TWO 128-query tiles (A, B) of one (image, head): K/V tiles are fetched once for both, and two
// softmax warpgroups (one thread per query row, the whole 128-score row of a key block in registers) ping-pong on the MUFU
// while the single MMA-issuing thread alternates QK^T / PV between the tiles.
//   warp 0 / lane 0 : TMA producer  (Q_A, Q_B, 2-deep K ring, 2-deep V ring; 128B-swizzled boxes of the packed QKV buffer)
//   warp 1 / lane 0 : MMA issuer    S_t = Q_t K_j^T (128x128x64, SS) -> TMEM S_t ;  O_t += P_t V_j (128x64x128, SS, V MN-major)
//   warps 2..5      : softmax warpgroup of tile A        warps 6..9 : softmax warpgroup of tile B
// Every score is read from TMEM exactly once (the TMEM read port, 64 B/clk/SM, and the MUFU, 16 exp/clk/SM, are the two
// units this kernel saturates); S_t is released to the next QK^T as soon as it sits in registers; O_t stays in TMEM and is
// only rescaled when the running row max grows by more than 2^8 (lazy rescale).
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kQ = 128;              // query rows per tile
constexpr int kKB = 128;             // keys per block
constexpr int kDh = 64;
constexpr int kTile = kQ * kDh * 2;  // 16 KiB
constexpr int kThreads = 320;
// smem: Q_A Q_B | K[2] | V[2] | P_A P_B (32 KiB each) | barriers
constexpr int kOffK = 2 * kTile, kOffV = 4 * kTile, kOffP = 6 * kTile, kOffBar = 6 * kTile + 2 * 2 * kTile;
constexpr int kSmemBytes = kOffBar + 256;  // 164,096 B
constexpr uint32_t kColS = 0, kColO = 256;  // TMEM: S_A 0, S_B 128, O_A 256, O_B 320
constexpr float kRescaleLog2 = 8.0f;

struct VitPairParams {
  __half* out;
  long out_stride;
  int S, D, n_heads;
  float scale_log2;
  uint32_t v_desc_hi, v_lbo;
  long long* dbg;  // optional clock64 trace of CTA (0,0,0) (tools/trace_vit_pair.py; VCB_VIT_TRACE=<device address>)
};

// barriers: per tile t: S ready, S free, P ready[2 bufs], O done
enum {
  kBarQ = 0,        // +t
  kBarKFull = 2, kBarKEmpty = 4, kBarVFull = 6, kBarVEmpty = 8,
  kBarS = 10,       // +t
  kBarSFree = 12,   // +t   (4 warps)
  kBarP = 14,       // +t   (4 warps)
  kBarO = 16,       // +t
  kNumBars = 18
};

__global__ void __launch_bounds__(kThreads, 1)
attn_vit_pair_kernel(const __grid_constant__ CUtensorMap tm, VitPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t bars = base + kOffBar;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * static_cast<uint32_t>(i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) p.dbg[128] = clock64();
  const int S = p.S;
  const int nb = (S + kKB - 1) / kKB;
  const int row_base = b * S;
  const int q0[2] = {(2 * pair) * kQ, (2 * pair + 1) * kQ};
  const int ntiles = q0[1] < S ? 2 : 1;  // the last pair of an odd tile count has no tile B

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) {
      const bool four = (i >= kBarSFree && i < kBarSFree + 2) || (i >= kBarP && i < kBarP + 2);
      mbar_init(bar(i), four ? 4 : 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm);
      for (int t = 0; t < ntiles; ++t) {
        mbar_expect_tx(bar(kBarQ + t), kTile);
        tma_load_2d(base + t * kTile, &tm, bar(kBarQ + t), h * kDh, row_base + q0[t]);
      }
      for (int j = 0; j < nb; ++j) {
        const int slot = j & 1;
        const uint32_t ph = static_cast<uint32_t>(((j >> 1) - 1) & 1);
        if (j >= 2) mbar_wait(bar(kBarKEmpty + slot), ph);
        mbar_expect_tx(bar(kBarKFull + slot), kTile);
        tma_load_2d(base + kOffK + slot * kTile, &tm, bar(kBarKFull + slot), p.D + h * kDh, row_base + j * kKB);
        if (j >= 2) mbar_wait(bar(kBarVEmpty + slot), ph);
        mbar_expect_tx(bar(kBarVFull + slot), kTile);
        tma_load_2d(base + kOffV + slot * kTile, &tm, bar(kBarVFull + slot), 2 * p.D + h * kDh, row_base + j * kKB);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kQ, kKB);
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(kQ, kDh);
      const uint64_t qdesc[2] = {make_sw128_kmajor_desc(base), make_sw128_kmajor_desc(base + kTile)};
      const uint64_t kdesc[2] = {make_sw128_kmajor_desc(base + kOffK), make_sw128_kmajor_desc(base + kOffK + kTile)};
      const uint64_t vhi = (static_cast<uint64_t>(p.v_desc_hi) << 32) | (static_cast<uint64_t>(p.v_lbo) << 16);
      const uint64_t vdesc[2] = {vhi | static_cast<uint64_t>(((base + kOffV) & 0x3FFFF) >> 4),
                                 vhi | static_cast<uint64_t>(((base + kOffV + kTile) & 0x3FFFF) >> 4)};
      auto issue_s = [&](int t, int j) {  // K_j must be resident (checked by the caller once per j)
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_f16(tmem + kColS + t * kKB, qdesc[t] + 2 * k, kdesc[j & 1] + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar(kBarS + t));
      };
      auto issue_pv = [&](int t, int j) {
        const int nkeys = min(kKB, S - j * kKB);
        const int ksteps = (nkeys + 15) / 16;
        const uint32_t pbuf = base + kOffP + t * 2 * kTile;  // P_t: two 64-key halves
        const uint64_t pd0 = make_sw128_kmajor_desc(pbuf), pd1 = make_sw128_kmajor_desc(pbuf + kTile);
        for (int s = 0; s < ksteps; ++s)
          umma_f16(tmem + kColO + t * kDh, (s < 4 ? pd0 : pd1) + 2 * (s & 3), vdesc[j & 1] + 128 * s, idesc_o,
                   (j > 0 || s > 0) ? 1u : 0u);
        umma_commit(bar(kBarO + t));
      };
      for (int t = 0; t < ntiles; ++t) mbar_wait(bar(kBarQ + t), 0);
      mbar_wait(bar(kBarKFull + 0), 0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_s(t, 0);
      umma_commit(bar(kBarKEmpty + 0));
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) {  // next QK^T of each tile as soon as its scores sit in registers
          mbar_wait(bar(kBarKFull + ((j + 1) & 1)), ((j + 1) >> 1) & 1);
          for (int t = 0; t < ntiles; ++t) {
            mbar_wait(bar(kBarSFree + t), j & 1);
            tc_fence_after();
            issue_s(t, j + 1);
          }
          umma_commit(bar(kBarKEmpty + ((j + 1) & 1)));
        }
        mbar_wait(bar(kBarVFull + (j & 1)), (j >> 1) & 1);
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar(kBarP + t), j & 1);
          tc_fence_after();
          issue_pv(t, j);
        }
        umma_commit(bar(kBarVEmpty + (j & 1)));
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups: one thread per query row
    const int t = (warp - 2) >> 2;  // tile A (warps 2..5) or B (warps 6..9)
    if (t < ntiles) {
      const int q = warp & 3;  // TMEM lane quadrant (hardware rule: warp id % 4)
      const int r = q * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const float sc = p.scale_log2;
      const uint32_t s_addr = tmem + lane_off + kColS + t * kKB;
      const uint32_t o_addr = tmem + lane_off + kColO + t * kDh;
      float Mref = -INFINITY, L = 0.f;

      const bool tr = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && q == 0 && lane == 0;
#define VCB_STAMP(k) do { if (tr) p.dbg[(t * 8 + j) * 8 + (k)] = clock64(); } while (0)
      for (int j = 0; j < nb; ++j) {
        const int nk = min(kKB, S - j * kKB);
        VCB_STAMP(0);
        mbar_wait(bar(kBarS + t), j & 1);
        tc_fence_after();
        VCB_STAMP(1);
        uint32_t v[4][32];
        tmem_ld_32x32(s_addr, v[0]);
        tmem_ld_32x32(s_addr + 32, v[1]);
        tmem_ld_32x32(s_addr + 64, v[2]);
        tmem_ld_32x32(s_addr + 96, v[3]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarSFree + t));
        VCB_STAMP(2);

        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        if (nk == kKB) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(v[c][i]));
              mx1 = fmaxf(mx1, __uint_as_float(v[c][i + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(v[c][i + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(v[c][i + 3]));
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < nk) mx0 = fmaxf(mx0, __uint_as_float(v[c][i]));
          }
        }
        const float mblk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

        // lazy rescale of O (in TMEM): only when the row max grew by more than 2^kRescaleLog2
        float f = 1.f;
        if (j == 0) {
          Mref = mblk;
        } else if ((mblk - Mref) * sc > kRescaleLog2) {
          f = fast_exp2((Mref - mblk) * sc);
          Mref = mblk;
          L *= f;
        }
        // P_{j-1} V_{j-1} must be complete before O is rescaled or the P tile is overwritten (it was issued right after
        // this warpgroup's previous arrival, so in steady state this wait is already satisfied)
        VCB_STAMP(3);
        if (j > 0) {
          mbar_wait(bar(kBarO + t), (j - 1) & 1);
          tc_fence_after();
        }
        VCB_STAMP(4);
        if (j > 0 && __any_sync(0xffffffffu, f != 1.f)) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(o_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32(o_addr + c * 32, ov);
          }
          tmem_st_wait();
        }
        VCB_STAMP(5);
        const float msc = Mref * sc;
        uint8_t* prow = smem + kOffP + t * 2 * kTile + r * 128;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < nk) {
            uint32_t pk[16];
            if (nk >= (c + 1) * 32) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc));
                const float p1 = fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc));
                l0 += p0;
                l1 += p1;
                pk[i >> 1] = pack_half2(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = (c * 32 + i < nk) ? fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc)) : 0.f;
                const float p1 = (c * 32 + i + 1 < nk) ? fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc)) : 0.f;
                l0 += p0;
                l1 += p1;
                pk[i >> 1] = pack_half2(p0, p1);
              }
            }
            // keys [c*32, c*32+32): 64-key half (c>>1), 16-byte chunks (c&1)*4 .. +3 of the 128-byte row
            uint8_t* dst = prow + (c >> 1) * kTile;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const int chunk = ((c & 1) * 4 + w) ^ (r & 7);
              *reinterpret_cast<uint4*>(dst + (chunk << 4)) = make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]);
            }
          }
        }
        L += l0 + l1;
        VCB_STAMP(6);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarP + t));
        VCB_STAMP(7);
      }
#undef VCB_STAMP

      // ---- epilogue: O (TMEM) / l -> fp16
      mbar_wait(bar(kBarO + t), (nb - 1) & 1);
      tc_fence_after();
      const int row = q0[t] + r;
      const float inv = 1.f / L;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(o_addr + c * 32, ov);
        tmem_ld_wait();
        if (row < S) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long>(row_base + row) * p.out_stride + h * kDh + c * 32);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            dst[w] = make_uint4(
                pack_half2(__uint_as_float(ov[8 * w]) * inv, __uint_as_float(ov[8 * w + 1]) * inv),
                pack_half2(__uint_as_float(ov[8 * w + 2]) * inv, __uint_as_float(ov[8 * w + 3]) * inv),
                pack_half2(__uint_as_float(ov[8 * w + 4]) * inv, __uint_as_float(ov[8 * w + 5]) * inv),
                pack_half2(__uint_as_float(ov[8 * w + 6]) * inv, __uint_as_float(ov[8 * w + 7]) * inv));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) p.dbg[129] = clock64();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int make_tmap_2d_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
                     uint32_t box_rows);

int attn_vit_pair(const AttnPrefillParams& a, cudaStream_t stream) {
  VCB_CHECK(a.head_dim == 64, kErrInvalid, "attn_vit_pair: head_dim must be 64");
  VCB_CHECK(a.q_stride == a.kv_stride && a.k == a.q + a.n_heads * 64 && a.v == a.q + 2 * a.n_heads * 64, kErrInvalid,
            "attn_vit_pair: expects the packed [tokens, 3*D] qkv layout");
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(attn_vit_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  const int D = a.n_heads * 64;
  CUtensorMap tm;
  VCB_TRY(make_tmap_2d_f16(&tm, a.q, 3 * D, static_cast<uint64_t>(a.batch) * a.seq_len, a.q_stride, 64, 128));
  VitPairParams p;
  p.out = a.out; p.out_stride = a.out_stride; p.S = a.seq_len; p.D = D; p.n_heads = a.n_heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.v_desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // MN-major V: SBO 1024 B, version 1, SWIZZLE_128B
  p.v_lbo = 0;
  p.dbg = nullptr;
  if (const char* e = getenv("VCB_VIT_TRACE")) p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 16));
  const int q_tiles = (a.seq_len + kQ - 1) / kQ;
  dim3 grid((q_tiles + 1) / 2, a.n_heads, a.batch);
  KernelScope scope(kCatAttnVit, stream);
  attn_vit_pair_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm, p);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb
