// CLIP ViT self-attention on tcgen05 + TMEM (sm_100a):  O = softmax(Q K^T * scale) V,  head_dim 64.
// Reference: CLIPVisionTower.forward -> transformers/models/clip/modeling_clip.py:261-345 (eager attention, no mask);
// SURVEY K4 — the kernel the north star names (4,726,784 algorithmic bytes and 1.364 GFLOP per image-layer at S=577).
//
// One CTA = one (image, head, 128-query tile); TWO CTAs are resident per SM (<= 113 KiB smem, 256 TMEM columns each) so
// that one CTA's exponentials (MUFU-bound) overlap the other's MMA / TMA latencies.
// TMA drops Q and a 2-deep ring of K and V tiles straight from the packed [tokens, 3*D] QKV buffer into 128B-swizzled
// shared memory (one 2-D tensor map, three column offsets).
//   warp 0 / lane 0 : TMA producer
//   warp 1 / lane 0 : MMA issuer    S_j = Q K_j^T  (128x128x64, SS)                      -> TMEM S      (fp32)
//                                    O_j = P_j V_j  (128x64x128, SS, V = MN-major B operand) -> TMEM O[j&1]
//   warps 2..5      : softmax, one thread per query row: tcgen05.ld the score row, running max / sum in fp32,
//                     P_j (fp16) written to swizzled smem as the next MMA's A operand, O_j pulled out of TMEM and
//                     folded into fp32 registers with the online-softmax correction; final O / l stored as fp16.
// Scores and probabilities never leave the SM.
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kQ = 128;              // query rows per CTA
constexpr int kKB = 128;             // keys per block
constexpr int kDh = 64;
constexpr int kTile = kQ * kDh * 2;  // 16 KiB: one [128 x 64] fp16 tile (Q, K_j or V_j); also one 64-key half of P
constexpr int kThreads = 192;
// smem: Q | K[2] | V[2] | P (2 x 16 KiB) | barriers      (dynamic smem base must be 1 KiB aligned: checked at run time)
constexpr int kSmemBytes = kTile * (1 + 2 + 2 + 2) + 256;
constexpr uint32_t kColS = 0, kColO = 128;  // TMEM columns: S at 0, O[2] at 128/192
constexpr uint32_t kTmemCols = 256;

struct VitTc5Params {
  __half* out;
  long out_stride;
  int S, D, n_heads;
  float scale_log2;    // scale * log2(e)
  uint32_t v_desc_hi;  // upper 32 bits of the V (MN-major) smem descriptor: SBO / version / layout
  uint32_t v_lbo;      // LBO field (>>4) of the V descriptor
};

enum {
  kBarQ = 0, kBarKFull = 1, kBarKEmpty = 3, kBarVFull = 5, kBarVEmpty = 7, kBarS = 9, kBarP = 10, kBarO = 11,
  kBarOC = 13, kNumBars = 15
};

__global__ void __launch_bounds__(kThreads, 2)
attn_vit_tc5_kernel(const __grid_constant__ CUtensorMap tm, VitTc5Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();  // the 128B swizzle needs 1 KiB-aligned tiles
  const uint32_t sQ = base, sK = base + kTile, sV = sK + 2 * kTile, sP = sV + 2 * kTile;
  const uint32_t bars = sP + 2 * kTile;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * static_cast<uint32_t>(i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bars - base) + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int q0 = qt * kQ;
  const int nb = (S + kKB - 1) / kKB;
  const int row_base = b * S;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) mbar_init(bar(i), (i == kBarP || i == kBarOC || i == kBarOC + 1) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm);
      mbar_expect_tx(bar(kBarQ), kTile);
      tma_load_2d(sQ, &tm, bar(kBarQ), h * kDh, row_base + q0);
      for (int j = 0; j < nb; ++j) {
        const int slot = j & 1;
        const uint32_t ph = static_cast<uint32_t>(((j >> 1) - 1) & 1);
        if (j >= 2) mbar_wait(bar(kBarKEmpty + slot), ph);  // S_{j-2} has consumed this K slot
        mbar_expect_tx(bar(kBarKFull + slot), kTile);
        tma_load_2d(sK + slot * kTile, &tm, bar(kBarKFull + slot), p.D + h * kDh, row_base + j * kKB);
        if (j >= 2) mbar_wait(bar(kBarVEmpty + slot), ph);  // P_{j-2} V_{j-2} has consumed this V slot
        mbar_expect_tx(bar(kBarVFull + slot), kTile);
        tma_load_2d(sV + slot * kTile, &tm, bar(kBarVFull + slot), 2 * p.D + h * kDh, row_base + j * kKB);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kQ, kKB);      // 128 x 128, both K-major
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(kQ, kDh);  // 128 x 64, B (= V) MN-major
      const uint64_t qdesc = make_sw128_kmajor_desc(sQ);
      auto issue_s = [&](int j) {
        const int slot = j & 1;
        mbar_wait(bar(kBarKFull + slot), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t kdesc = make_sw128_kmajor_desc(sK + slot * kTile);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) umma_f16(tmem + kColS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar(kBarS));
        umma_commit(bar(kBarKEmpty + slot));
      };
      auto issue_pv = [&](int j) {
        const int slot = j & 1;
        const int nkeys = min(kKB, S - j * kKB);
        const int ksteps = (nkeys + 15) / 16;
        for (int s = 0; s < ksteps; ++s) {
          const uint64_t pdesc = make_sw128_kmajor_desc(sP + (s >> 2) * kTile) + 2 * (s & 3);
          const uint32_t vaddr = sV + slot * kTile + s * 2048;  // 16 keys = 16 rows of 128 B
          const uint64_t vdesc = (static_cast<uint64_t>(p.v_desc_hi) << 32) | (static_cast<uint64_t>(p.v_lbo) << 16) |
                                 static_cast<uint64_t>((vaddr & 0x3FFFF) >> 4);
          umma_f16(tmem + kColO + slot * kDh, pdesc, vdesc, idesc_o, s != 0 ? 1u : 0u);
        }
        umma_commit(bar(kBarO + slot));
        umma_commit(bar(kBarVEmpty + slot));
      };
      mbar_wait(bar(kBarQ), 0);
      issue_s(0);
      for (int j = 0; j < nb; ++j) {
        mbar_wait(bar(kBarP), j & 1);  // P_j is in smem and S_j has been read out of TMEM
        mbar_wait(bar(kBarVFull + (j & 1)), (j >> 1) & 1);
        if (j >= 2) mbar_wait(bar(kBarOC + (j & 1)), ((j >> 1) - 1) & 1);  // O_{j-2} has been read out of TMEM
        tc_fence_after();
        issue_pv(j);
        if (j + 1 < nb) issue_s(j + 1);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps: thread = query row
    const int q = warp & 3;        // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;   // row within the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float sc = p.scale_log2;
    float o[kDh];
#pragma unroll
    for (int i = 0; i < kDh; ++i) o[i] = 0.f;
    float M = -INFINITY, L = 0.f, corr_prev = 0.f;

    auto fold_o = [&](int j, float corr) {  // o = o * corr + O_j
      mbar_wait(bar(kBarO + (j & 1)), (j >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem + lane_off + kColO + (j & 1) * kDh, v0);
      tmem_ld_32x32(tmem + lane_off + kColO + (j & 1) * kDh + 32, v1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = fmaf(o[i], corr, __uint_as_float(v0[i]));
        o[32 + i] = fmaf(o[32 + i], corr, __uint_as_float(v1[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarOC + (j & 1)));
    };

    const uint32_t s_addr = tmem + lane_off + kColS;
    for (int j = 0; j < nb; ++j) {
      const int nkeys = min(kKB, S - j * kKB);
      const int nchunks = (nkeys + 31) / 32;
      mbar_wait(bar(kBarS), j & 1);
      tc_fence_after();
      // pass A: block max.  Only the last chunk of the last key block is partial: full chunks take the mask-free path.
      float mx0 = M, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
        const int valid = nkeys - c * 32;
        if (valid >= 32) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(v[i]));
            mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < valid) mx0 = fmaxf(mx0, __uint_as_float(v[i]));
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float corr = fast_exp2((M - mx) * sc);  // M = -inf on the first block -> 0
      const float msc = mx * sc;
      M = mx;
      // the previous block's PV result (also frees the P buffer for this block's probabilities)
      if (j > 0) fold_o(j - 1, corr_prev);
      corr_prev = corr;
      // pass B: probabilities -> fp16 -> swizzled smem (A operand of the PV MMA), row sum
      float l = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
        const int valid = nkeys - c * 32;
        if (valid >= 32) {
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), sc, -msc));
            const float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, -msc));
            l0 += p0;
            l1 += p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
          l += l0 + l1;
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (i < valid) ? fast_exp2(fmaf(__uint_as_float(v[i]), sc, -msc)) : 0.f;
            const float p1 = (i + 1 < valid) ? fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, -msc)) : 0.f;
            l += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
        }
        // keys [c*32, c*32+32) of row r: 64-key half (c>>1), 16-byte chunks (c&1)*4 .. +3 of the 128-byte row
        uint8_t* prow = smem + (sP - base) + (c >> 1) * kTile + r * 128;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int chunk = ((c & 1) * 4 + w) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + (chunk << 4)) = make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]);
        }
      }
      L = L * corr + l;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarP));
    }
    fold_o(nb - 1, corr_prev);
    const int row = q0 + r;
    if (row < S) {
      const float inv = 1.f / L;
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long>(row_base + row) * p.out_stride + h * kDh);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        dst[w] = make_uint4(pack_half2(o[8 * w] * inv, o[8 * w + 1] * inv), pack_half2(o[8 * w + 2] * inv, o[8 * w + 3] * inv),
                            pack_half2(o[8 * w + 4] * inv, o[8 * w + 5] * inv), pack_half2(o[8 * w + 6] * inv, o[8 * w + 7] * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

}  // namespace

int make_tmap_2d_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
                     uint32_t box_rows);

int attn_vit_tc5(const AttnPrefillParams& a, cudaStream_t stream) {
  VCB_CHECK(a.head_dim == 64, kErrInvalid, "attn_vit_tc5: head_dim must be 64");
  VCB_CHECK(a.q_stride == a.kv_stride && a.k == a.q + a.n_heads * 64 && a.v == a.q + 2 * a.n_heads * 64, kErrInvalid,
            "attn_vit_tc5: expects the packed [tokens, 3*D] qkv layout");
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(attn_vit_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  const int D = a.n_heads * 64;
  CUtensorMap tm;
  VCB_TRY(make_tmap_2d_f16(&tm, a.q, 3 * D, static_cast<uint64_t>(a.batch) * a.seq_len, a.q_stride, 64, 128));
  VitTc5Params p;
  p.out = a.out; p.out_stride = a.out_stride; p.S = a.seq_len; p.D = D; p.n_heads = a.n_heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  // V descriptor (MN-major, 128B swizzle): SBO = 1024 B between 8-key groups, version 1, layout SWIZZLE_128B
  // (verified on B200: profiles/r01_vit_tc5_bringup.log)
  p.v_desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  p.v_lbo = 0;
  dim3 grid((a.seq_len + kQ - 1) / kQ, a.n_heads, a.batch);
  KernelScope scope(kCatAttnVit, stream);
  attn_vit_tc5_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm, p);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb
