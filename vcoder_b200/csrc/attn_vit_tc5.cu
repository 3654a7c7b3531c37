// CLIP ViT self-attention on tcgen05 + TMEM (sm_100a):  O = softmax(Q K^T * scale) V,  head_dim 64.
// Reference: CLIPVisionTower.forward -> transformers/models/clip/modeling_clip.py:261-345 (eager attention, no mask);
// SURVEY K4 — the kernel the north star names (4,726,784 algorithmic bytes and 1.364 GFLOP per image-layer at S=577).
//
// One CTA = one (image, head, 128-query tile); TWO CTAs are resident per SM (<= 113 KiB smem, 256 TMEM columns each).
// TMA drops Q and a 2-deep ring of K and V tiles straight from the packed [tokens, 3*D] QKV buffer into 128B-swizzled
// shared memory (one 2-D tensor map, three column offsets).
//   warp 0 / lane 0 : TMA producer
//   warp 1 / lane 0 : MMA issuer    S_j = Q K_j^T  (128x128x64, SS)                          -> TMEM S  (fp32)
//                                    O  += P_j V_j  (128x64x128, SS, V = MN-major B operand)  -> TMEM O  (fp32)
//   warps 2..9      : softmax, two threads per query row (64 key columns each).
// The measured limiter of this kernel class is the TMEM read port (64 B/clk/SM) together with MUFU (16 exp/clk/SM), so
// the data flow is built around reading every score exactly once: a thread pulls its 64 scores into registers (which
// frees the S buffer for the next QK^T immediately), the two half-row maxima are exchanged through smem, and O stays in
// TMEM across key blocks — it is only touched by the CUDA cores when the running max grows by more than 2^8 (lazy
// rescale, as in FlashAttention-4), and once at the end for the 1/l normalisation.
// Built only with VCB_EXPERIMENTAL=1 (python -m vcoder_b200.build): superseded by the pair schedule of attn_vit_pair.cu.
#ifdef VCB_EXPERIMENTAL
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
constexpr int kThreads = 320;        // TMA warp + MMA warp + 8 softmax warps
constexpr int kSoftThreads = 256;
// smem: Q | K[2] | V[2] | P (2 x 16 KiB) | barriers      (dynamic smem base must be 1 KiB aligned: checked at run time)
constexpr int kSmemBytes = kTile * (1 + 2 + 2 + 2) + 768;  // 115,456 B: two CTAs per SM (limit 115,712)
constexpr int kOffXchg = 128;                              // inside the 768-byte tail: [barriers | tmem slot | fp16 row maxima]
constexpr uint32_t kColS = 0, kColO = 128;                 // TMEM columns: S at 0, O at 128
constexpr uint32_t kTmemCols = 256;
constexpr float kRescaleLog2 = 8.0f;  // P stays below 2^8: exact enough in fp16, far from overflow

struct VitTc5Params {
  __half* out;
  long out_stride;
  int S, D, n_heads;
  float scale_log2;    // scale * log2(e)
  uint32_t v_desc_hi;  // upper 32 bits of the V (MN-major) smem descriptor: SBO / version / layout
  uint32_t v_lbo;      // LBO field (>>4) of the V descriptor
  long long* dbg;      // optional clock64 trace of CTA (0,0,0) (bring-up / profiling only)
};

enum {
  kBarQ = 0, kBarKFull = 1, kBarKEmpty = 3, kBarVFull = 5, kBarVEmpty = 7, kBarS = 9, kBarSFree = 10, kBarP = 11,
  kBarO = 12, kNumBars = 13
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
    for (int i = 0; i < kNumBars; ++i) mbar_init(bar(i), (i == kBarP || i == kBarSFree) ? 8 : 1);
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
      // descriptors precomputed once: the issuing thread shares its SM sub-partition with two softmax warps, so every
      // instruction between "P is ready" and the tcgen05.mma issue is exposed latency
      const uint64_t pdesc0 = make_sw128_kmajor_desc(sP), pdesc1 = make_sw128_kmajor_desc(sP + kTile);
      const uint64_t vdesc_hi = (static_cast<uint64_t>(p.v_desc_hi) << 32) | (static_cast<uint64_t>(p.v_lbo) << 16);
      const uint64_t vdesc_s0 = vdesc_hi | static_cast<uint64_t>((sV & 0x3FFFF) >> 4);
      const uint64_t vdesc_s1 = vdesc_hi | static_cast<uint64_t>(((sV + kTile) & 0x3FFFF) >> 4);
      auto issue_pv = [&](int j) {
        const int slot = j & 1;
        const int nkeys = min(kKB, S - j * kKB);
        const uint64_t vd = slot ? vdesc_s1 : vdesc_s0;
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        if (nkeys == kKB) {
#pragma unroll
          for (int s = 0; s < 8; ++s)  // 16 keys per step: +32 B along P's k, +2048 B (16 rows) along V
            umma_f16(tmem + kColO, (s < 4 ? pdesc0 : pdesc1) + 2 * (s & 3), vd + 128 * s, idesc_o, s == 0 ? acc0 : 1u);
        } else {
          const int ksteps = (nkeys + 15) / 16;
          for (int s = 0; s < ksteps; ++s)
            umma_f16(tmem + kColO, (s < 4 ? pdesc0 : pdesc1) + 2 * (s & 3), vd + 128 * s, idesc_o, s == 0 ? acc0 : 1u);
        }
        umma_commit(bar(kBarO));
        umma_commit(bar(kBarVEmpty + slot));
      };
      mbar_wait(bar(kBarQ), 0);
      issue_s(0);
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) {
          mbar_wait(bar(kBarSFree), j & 1);  // the softmax warps hold S_j in registers: S may be overwritten
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(bar(kBarP), j & 1);  // P_j is in smem (and O has been rescaled if that was needed)
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[j * 8 + 6] = clock64();
        mbar_wait(bar(kBarVFull + (j & 1)), (j >> 1) & 1);
        tc_fence_after();
        issue_pv(j);
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[j * 8 + 7] = clock64();
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps: two threads per query row
    const int q = warp & 3;            // TMEM lane quadrant of this warp (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;  // which 64 key columns of each block / which 32 columns of O this thread owns
    const int r = q * 32 + lane;       // row within the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float sc = p.scale_log2;
    const uint32_t s_addr = tmem + lane_off + kColS + half * 64;
    const uint32_t o_addr = tmem + lane_off + kColO + half * 32;
    uint8_t* my_row = smem + (sP - base) + half * kTile + r * 128;            // this thread's 64 keys of P
    __half* xchg = reinterpret_cast<__half*>(smem + (bars - base) + kOffXchg);     // [2 halves][128 rows] fp16 row maxima
    float Mref = -INFINITY, L = 0.f;

    const bool trace = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
#define VCB_STAMP(slot) do { if (trace) p.dbg[(j) * 8 + (slot)] = clock64(); } while (0)
    for (int j = 0; j < nb; ++j) {
      const int nk = min(kKB, S - j * kKB) - half * 64;  // valid keys among this thread's 64 columns (may be <= 0)
      VCB_STAMP(0);
      mbar_wait(bar(kBarS), j & 1);
      tc_fence_after();
      VCB_STAMP(1);
      uint32_t v[2][32];
      tmem_ld_32x32(s_addr, v[0]);
      tmem_ld_32x32(s_addr + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarSFree));  // S_{j+1} may start while we exponentiate S_j
      VCB_STAMP(2);

      // local max over this thread's valid columns
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (nk >= 64) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
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
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nk) mx0 = fmaxf(mx0, __uint_as_float(v[c][i]));
        }
      }
      const float mloc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

      // exchange the half-row maxima through dedicated fp16 slots (both threads then use the same reference max).
      // The slot may be overwritten once the partner has read the previous block's value, i.e. after every warp has
      // arrived on bar_p of block j-1.
      if (j > 0) mbar_wait(bar(kBarP), (j - 1) & 1);
      const __half mh = __float2half_rn(fminf(fmaxf(mloc, -60000.f), 60000.f));
      xchg[half * kQ + r] = mh;
      asm volatile("bar.sync 1, %0;" ::"n"(kSoftThreads) : "memory");
      const float mblk = fmaxf(__half2float(mh), __half2float(xchg[(half ^ 1) * kQ + r]));
      VCB_STAMP(3);

      // lazy rescale: keep the old reference max unless the row max grew by more than 2^kRescaleLog2
      float f = 1.f;
      if (j == 0) {
        Mref = mblk;
      } else if ((mblk - Mref) * sc > kRescaleLog2) {
        f = fast_exp2((Mref - mblk) * sc);
        Mref = mblk;
        L *= f;
      }
      bool waited_o = false;
      if (j > 0 && __any_sync(0xffffffffu, f != 1.f)) {  // warp-uniform: tcgen05.ld/st are warp-collective
        mbar_wait(bar(kBarO), (j - 1) & 1);  // P_{j-1} V_{j-1} has landed in O
        tc_fence_after();
        waited_o = true;
        uint32_t ov[32];
        tmem_ld_32x32(o_addr, ov);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
        tmem_st_32x32(o_addr, ov);
        tmem_st_wait();
      }
      const float msc = Mref * sc;
      VCB_STAMP(4);

      // probabilities (fp16 pairs) into registers first: the MUFU work overlaps the PV MMA of the previous block
      uint32_t pk[2][16];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (nk >= (c + 1) * 32) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc));
            const float p1 = fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc));
            l0 += p0;
            l1 += p1;
            pk[c][i >> 1] = pack_half2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < nk) ? fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc)) : 0.f;
            const float p1 = (c * 32 + i + 1 < nk) ? fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc)) : 0.f;
            l0 += p0;
            l1 += p1;
            pk[c][i >> 1] = pack_half2(p0, p1);
          }
        }
      }
      // the P tile is free once P_{j-1} V_{j-1} has completed
      if (j > 0 && !waited_o) {
        mbar_wait(bar(kBarO), (j - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 < nk) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int chunk = (c * 4 + w) ^ (r & 7);
            *reinterpret_cast<uint4*>(my_row + (chunk << 4)) =
                make_uint4(pk[c][4 * w], pk[c][4 * w + 1], pk[c][4 * w + 2], pk[c][4 * w + 3]);
          }
        }
      }
      L += l0 + l1;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarP));
      VCB_STAMP(5);
    }
#undef VCB_STAMP

    // ---- epilogue: O (TMEM) / l  -> fp16
    mbar_wait(bar(kBarO), (nb - 1) & 1);
    tc_fence_after();
    uint32_t ov[32];
    tmem_ld_32x32(o_addr, ov);
    tmem_ld_wait();
    // total row sum = the two halves' partial sums (both relative to the same reference max).  The P tile is free now.
    float* xb = reinterpret_cast<float*>(smem + (sP - base));
    xb[half * kQ + r] = L;
    asm volatile("bar.sync 1, %0;" ::"n"(kSoftThreads) : "memory");
    L += xb[(half ^ 1) * kQ + r];
    const int row = q0 + r;
    if (row < S) {
      const float inv = 1.f / L;
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long>(row_base + row) * p.out_stride + h * kDh + half * 32);
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
  p.dbg = nullptr;
  if (const char* e = getenv("VCB_VIT_TRACE")) {  // profiling aid: device buffer of >= 64 int64 (address in hex)
    p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 16));
  }
  dim3 grid((a.seq_len + kQ - 1) / kQ, a.n_heads, a.batch);
  KernelScope scope(kCatAttnVit, stream);
  attn_vit_tc5_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm, p);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb

#else  // !VCB_EXPERIMENTAL
#include "common.h"
#include "kernels.h"
namespace vcb {
int attn_vit_tc5(const AttnPrefillParams& p, cudaStream_t stream) { return attn_vit_pair(p, stream); }
}  // namespace vcb
#endif
