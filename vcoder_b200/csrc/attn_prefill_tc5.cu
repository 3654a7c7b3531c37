// LLaMA causal prefill attention over the paged KV cache on tcgen05 + TMEM (sm_100a), head_dim 128:
//   O = softmax_fp32(Q K^T / sqrt(128) + causal) V          (transformers/models/llama/modeling_llama.py:199-222, eager
//   attention, reached from VCoder*LlavaLlamaForCausalLM.forward, vcoder_ds_llava_llama.py:81).
// Replaces the mma.sync / cp.async kernel of attn_prefill.cu (kept for A/B and as the fallback for odd shapes).
//
// One CTA = one (sequence, head, 128-query tile); key blocks are ONE KV page (64 keys), so a block is two TMA boxes per
// operand straight out of the page pool ([page][head][64 tokens][128] viewed as a 3-D tensor, 128B-swizzled 64-wide
// halves of head_dim).  TWO CTAs are resident per SM (112 KiB shared memory, 256 TMEM columns each) and hide each other's
// softmax / MMA / barrier latencies.
//   warp 5 / lane 0 : TMA producer (Q once; 2-deep rings of K and V pages; page ids from the block table)
//   warp 4 / lane 0 : MMA issuer    S_j = Q K_j^T   (128 x 64 x 128: 8 k-steps over the two head_dim halves, SS) -> TMEM S
//                                    O  += P_j V_j   (two 128 x 64 x 64 MMAs, one per head_dim half; V is the MN-major B
//                                                     operand exactly as TMA delivers a page)                    -> TMEM O
//   warps 0..3      : softmax, one thread per query row: the 64 scores of a block are read from TMEM once, masked
//                     causally (only the diagonal blocks need it), exponentiated (exp2, log2e folded into the scale) and
//                     parked as fp16 in a swizzled smem tile = the A operand of the PV MMA.  O stays in TMEM and is only
//                     rescaled when the running row max grew by more than 2^8 (lazy rescale).
// Causal tile skipping: a query tile visits only the key blocks up to its own diagonal; heavy tiles are scheduled first.
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kQ = 128;    // query rows per CTA
constexpr int kKB = 64;    // keys per block = tokens per KV page
constexpr int kDh = 128;
constexpr int kSub = 64;   // head_dim elements per 128-byte swizzled row
constexpr int kQSub = kQ * kSub * 2;   // 16 KiB: [128 rows][64 dh]
constexpr int kKVSub = kKB * kSub * 2; // 8 KiB:  [64 keys][64 dh]
constexpr int kThreads = 192;          // 4 softmax warps + MMA warp + TMA warp
// the warp scheduler prefers the highest ready warp id: the single-thread MMA issuer / TMA producer sit ABOVE the softmax
// warps (as warp 1 the issuer of the ViT kernel started a PV MMA thousands of cycles after its operands were ready)
constexpr int kMmaWarp = 4, kTmaWarp = 5;
// smem: Q (2 halves) | K ring [2][2 halves] | V ring [2][2 halves] | P [128][64 keys] | barriers
constexpr int kOffK = 2 * kQSub, kOffV = kOffK + 4 * kKVSub, kOffP = kOffV + 4 * kKVSub, kOffBar = kOffP + kQ * kKB * 2;
constexpr int kSmemBytes = kOffBar + 768;  // 115,456 B: two CTAs per SM
constexpr uint32_t kColS = 0, kColO = 64;  // TMEM columns: S [0,64), O [64,192)
constexpr uint32_t kTmemCols = 256;
constexpr float kRescaleLog2 = 8.0f;

struct PrefillTc5Params {
  __half* out;
  long out_stride;
  const int32_t* block_table;
  int max_pages, n_heads;
  int S;       // query rows per sequence
  int P0;      // position of query row 0 (> 0: a cached prefix is reused)
  float scale_log2;
  uint32_t v_desc_hi;
};

enum {
  kBarQ = 0, kBarKFull = 1, kBarKEmpty = 3, kBarVFull = 5, kBarVEmpty = 7, kBarS = 9, kBarSFree = 10, kBarP = 11,
  kBarO = 12, kNumBars = 13
};

__global__ void __launch_bounds__(kThreads, 2)
attn_prefill_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, PrefillTc5Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();  // the 128B swizzle needs 1 KiB-aligned tiles
  const uint32_t sQ = base, sK = base + kOffK, sV = base + kOffV, sP = base + kOffP;
  const uint32_t bars = base + kOffBar;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * static_cast<uint32_t>(i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);  // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int S = p.S, P0 = p.P0;
  const int q0 = qt * kQ;
  const int KV = P0 + S;  // keys visible to the last query row of the sequence
  // key blocks this tile needs: up to the block that holds the position of its last (valid) row
  const int last_pos = P0 + min(q0 + kQ, S) - 1;
  const int nb = last_pos / kKB + 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) mbar_init(bar(i), (i == kBarP || i == kBarSFree) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kTmaWarp) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(bar(kBarQ), 2 * kQSub);
      tma_load_2d(sQ, &tmQ, bar(kBarQ), h * kDh, b * S + q0);
      tma_load_2d(sQ + kQSub, &tmQ, bar(kBarQ), h * kDh + kSub, b * S + q0);
      const int32_t* bt = p.block_table + static_cast<long>(b) * p.max_pages;
      for (int j = 0; j < nb; ++j) {
        const int slot = j & 1;
        const uint32_t ph = static_cast<uint32_t>(((j >> 1) - 1) & 1);
        const int pg = bt[j] * p.n_heads + h;  // row of the [pages*heads][64][128] view
        if (j >= 2) mbar_wait(bar(kBarKEmpty + slot), ph);  // S_{j-2} has consumed this K slot
        mbar_expect_tx(bar(kBarKFull + slot), 2 * kKVSub);
        tma_load_3d(sK + slot * 2 * kKVSub, &tmK, bar(kBarKFull + slot), 0, 0, pg);
        tma_load_3d(sK + slot * 2 * kKVSub + kKVSub, &tmK, bar(kBarKFull + slot), kSub, 0, pg);
        if (j >= 2) mbar_wait(bar(kBarVEmpty + slot), ph);  // P_{j-2} V_{j-2} has consumed this V slot
        mbar_expect_tx(bar(kBarVFull + slot), 2 * kKVSub);
        tma_load_3d(sV + slot * 2 * kKVSub, &tmV, bar(kBarVFull + slot), 0, 0, pg);
        tma_load_3d(sV + slot * 2 * kKVSub + kKVSub, &tmV, bar(kBarVFull + slot), kSub, 0, pg);
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kQ, kKB);       // 128 x 64, both operands K-major
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(kQ, kSub);  // 128 x 64, B (= V half) MN-major
      const uint64_t qdesc[2] = {make_sw128_kmajor_desc(sQ), make_sw128_kmajor_desc(sQ + kQSub)};
      const uint64_t pdesc = make_sw128_kmajor_desc(sP);
      const uint64_t vhi = static_cast<uint64_t>(p.v_desc_hi) << 32;
      auto issue_s = [&](int j) {
        const int slot = j & 1;
        mbar_wait(bar(kBarKFull + slot), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t kdesc[2] = {make_sw128_kmajor_desc(sK + slot * 2 * kKVSub),
                                   make_sw128_kmajor_desc(sK + slot * 2 * kKVSub + kKVSub)};
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)  // 16 head_dim elements per step: +32 B inside a 64-wide half
          umma_f16(tmem + kColS, qdesc[k >> 2] + 2 * (k & 3), kdesc[k >> 2] + 2 * (k & 3), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar(kBarS));
        umma_commit(bar(kBarKEmpty + slot));
      };
      auto issue_pv = [&](int j) {
        const int slot = j & 1;
        const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint64_t vd = vhi | static_cast<uint64_t>(((sV + slot * 2 * kKVSub + hf * kKVSub) & 0x3FFFF) >> 4);
#pragma unroll
          for (int s = 0; s < kKB / 16; ++s)  // 16 keys per step: +32 B along P's k, +2048 B (16 rows) along V
            umma_f16(tmem + kColO + hf * kSub, pdesc + 2 * s, vd + 128 * s, idesc_o, s == 0 ? acc0 : 1u);
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
        mbar_wait(bar(kBarVFull + (j & 1)), (j >> 1) & 1);
        tc_fence_after();
        issue_pv(j);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps: one thread per query row
    const int q = warp & 3;  // TMEM lane quadrant of this warp (hardware rule: warp id % 4)
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float sc = p.scale_log2;
    const uint32_t s_addr = tmem + lane_off + kColS;
    const uint32_t o_addr = tmem + lane_off + kColO;
    uint8_t* prow = smem + kOffP + r * 128;
    const int qpos = P0 + q0 + r;  // position of this thread's query row (rows past S compute garbage that is never stored)
    float Mref = -INFINITY, L = 0.f;

    for (int j = 0; j < nb; ++j) {
      // valid keys of this block for this row: key position j*64 + c <= qpos (and < KV, implied for stored rows)
      const int nk = max(0, min(kKB, min(qpos, KV - 1) - j * kKB + 1));
      mbar_wait(bar(kBarS), j & 1);
      tc_fence_after();
      uint32_t v[2][32];
      tmem_ld_32x32(s_addr, v[0]);
      tmem_ld_32x32(s_addr + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarSFree));  // S_{j+1} may start while we exponentiate S_j

      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (nk == kKB) {
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
      const float mblk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));  // -inf when the whole block is masked for this row

      // The last page of a sequence is only partly written: its tail rows hold whatever the pool held before (possibly
      // NaN bit patterns), and 0 * NaN would poison O.  Zero those V rows in shared memory before the PV MMA reads them
      // (a row is 128 contiguous bytes in each head_dim half; the swizzle permutes inside a row only).
      const int nkeys_blk = KV - j * kKB;
      if (nkeys_blk < kKB) {  // CTA-uniform; at most once per CTA
        mbar_wait(bar(kBarVFull + (j & 1)), (j >> 1) & 1);
        uint8_t* vb = smem + kOffV + (j & 1) * 2 * kKVSub;
        for (int hf = 0; hf < 2; ++hf)
          for (int off = nkeys_blk * 128 + static_cast<int>(threadIdx.x) * 16; off < kKVSub; off += 128 * 16)
            *reinterpret_cast<uint4*>(vb + hf * kKVSub + off) = make_uint4(0, 0, 0, 0);
      }

      // lazy rescale of O (in TMEM): only when the row max grew by more than 2^kRescaleLog2
      float f = 1.f;
      if (j == 0) {
        Mref = mblk;  // block 0 always holds key 0 <= qpos: finite
      } else if ((mblk - Mref) * sc > kRescaleLog2) {
        f = fast_exp2((Mref - mblk) * sc);
        Mref = mblk;
        L *= f;
      }
      // P_{j-1} V_{j-1} must be complete before O is rescaled or the P tile is overwritten
      if (j > 0) {
        mbar_wait(bar(kBarO), (j - 1) & 1);
        tc_fence_after();
      }
      if (j > 0 && __any_sync(0xffffffffu, f != 1.f)) {  // warp-uniform: tcgen05.ld/st are warp-collective
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32(o_addr + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
          tmem_st_32x32(o_addr + c * 32, ov);
        }
        tmem_st_wait();
      }
      const float msc = Mref * sc;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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
        // keys [c*32, c*32+32) = 16-byte chunks c*4 .. c*4+3 of this row's 128 bytes (masked keys are written as zeros:
        // the PV MMA always runs over the whole 64-key block)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int chunk = (c * 4 + w) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + (chunk << 4)) = make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]);
        }
      }
      L += l0 + l1;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarP));
    }

    // ---- epilogue: O (TMEM) / l -> fp16
    mbar_wait(bar(kBarO), (nb - 1) & 1);
    tc_fence_after();
    const int row = q0 + r;
    const float inv = 1.f / L;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32(o_addr + c * 32, ov);
      tmem_ld_wait();
      if (row < S) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long>(b) * S + row) * p.out_stride + h * kDh + c * 32);
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

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

}  // namespace

int make_tmap_2d_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
                     uint32_t box_rows);
int make_tmap_pages_f16(CUtensorMap* m, const void* ptr, uint64_t n_rows);

int attn_llama_prefill_tc5(const AttnPrefillParams& a, cudaStream_t stream) {
  VCB_CHECK(a.head_dim == 128 && a.page_tokens == kKB, kErrInvalid, "attn_llama_prefill_tc5: head_dim 128, 64-token pages");
  VCB_CHECK(a.q_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && a.out_stride % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
            kErrInvalid, "attn_llama_prefill_tc5: q / out must be 16-byte aligned with strides % 8 == 0");
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(attn_prefill_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t q_cols = static_cast<uint64_t>(a.n_heads) * kDh;
  VCB_TRY(make_tmap_2d_f16(&tmQ, a.q, q_cols, static_cast<uint64_t>(a.batch) * a.seq_len, a.q_stride, kSub, kQ));
  // pools viewed as [pages * heads][64 tokens][128]: when the pool size is not known (operator-level call) the view is
  // sized generously — TMA only bounds-checks against the view, the page ids come from the block table
  const uint64_t rows = a.pool_pages > 0 ? static_cast<uint64_t>(a.pool_pages) * a.n_heads : (1ull << 22);
  VCB_TRY(make_tmap_pages_f16(&tmK, a.k_pool, rows));
  VCB_TRY(make_tmap_pages_f16(&tmV, a.v_pool, rows));
  PrefillTc5Params p;
  p.out = a.out; p.out_stride = a.out_stride; p.block_table = a.block_table; p.max_pages = a.max_pages;
  p.n_heads = a.n_heads; p.S = a.seq_len; p.P0 = a.q_pos0;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.v_desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // MN-major V: SBO 1024 B, version 1, SWIZZLE_128B (as attn_vit_*)
  dim3 grid((a.seq_len + kQ - 1) / kQ, a.n_heads, a.batch);
  KernelScope scope(kCatAttnPrefill, stream);
  attn_prefill_tc5_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmQ, tmK, tmV, p);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb
