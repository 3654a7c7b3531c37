// Decode attention over the paged KV cache (flash-decoding: split along the context, then combine).
// Replaces HF eager attention at decode: softmax_fp32(q·K^T / sqrt(128)) · V over EVERY cached position — the
// reference substitutes an all-ones mask of length past+1 on each decode step
// (vcoder_llava/model/vcoder_ds_llava_arch.py:130-133; arithmetic modeling_llama.py:199-222).
//
// HBM-bound (B·ctx·2·H·2 bytes per layer per step), so the kernel is organised around bytes in flight:
// one CTA = (256-key split, head, sequence); a half-warp owns every 8th key of a 64-key page and its 16 lanes cover
// one 128-dim row with 16-byte loads; per page each thread issues its 8 K-row and 8 V-row loads back to back (16
// independent 16-byte loads in flight) and only then reduces.  Every half-warp runs its own online softmax, so the
// main loop has no block-level synchronisation; the 8 partial (max, sum, acc) states are merged once through smem.
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kChunk = 256;  // default keys per split (AttnDecodeParams::chunk; VCB_ATTN_DEC_CHUNK = 128 | 256 | 512)
constexpr int kPT = 64;      // tokens per KV page (checked on the host)
constexpr int kDecThreads = 128;
constexpr int kDH = 128;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Merge of the flash-decoding partials WITHOUT a second kernel: every CTA of a (sequence, head) publishes its partial
// state, fences and takes a ticket; the CTA that draws the last ticket reads all partials back (L2) and writes the fp16
// output.  The partials are always merged in split order, so the result does not depend on which CTA happens to be last.
// Saves the combine launch and its dependency boundary (2.4 + 0.6 us per layer at B = 8, ctx 1271).
__device__ __forceinline__ void merge_by_last_cta(const AttnDecodeParams& p, int splits, int b, int h, int tid, int* s_last) {
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int done = atomicAdd(p.sync_counters + b * p.n_heads + h, 1);
    *s_last = (done == splits - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*s_last == 0) return;
  __threadfence();
  const long bh = static_cast<long>(b) * p.n_heads + h;
  float Mg = -INFINITY;
  for (int sp = 0; sp < splits; ++sp) Mg = fmaxf(Mg, __ldcg(p.part_ml + (bh * splits + sp) * 2));
  float num = 0.f, den = 0.f;
  for (int sp = 0; sp < splits; ++sp) {
    const long pi = bh * splits + sp;
    const float mm = __ldcg(p.part_ml + pi * 2);
    if (mm == -INFINITY) continue;
    const float w = exp2f(mm - Mg);
    num += w * __ldcg(p.part_o + pi * kDH + tid);
    den += w * __ldcg(p.part_ml + pi * 2 + 1);
  }
  p.out[bh * kDH + tid] = __float2half_rn(num / den);
  if (tid == 0) p.sync_counters[b * p.n_heads + h] = 0;  // ready for the next launch (stream-ordered)
}

// CLUSTER = true: the `splits` CTAs of one (sequence, head) form a thread-block cluster; the partial (max, sum, acc) states
// stay in shared memory and CTA 0 of the cluster merges them through distributed shared memory and writes the fp16 output
// itself — no partial buffers in HBM, no combine kernel (one launch and one dependency boundary less per layer).
template <bool CLUSTER, int MINB>
__global__ void __launch_bounds__(kDecThreads, MINB)
attn_decode_split_kernel(AttnDecodeParams p, int splits) {
  __shared__ float s_acc[8][kDH];
  __shared__ float s_m[8], s_l[8];
  __shared__ float s_fin[kDH + 2];  // CLUSTER: this split's merged acc[128], max, sum (read by CTA 0 of the cluster)
  __shared__ int s_last;
  griddep_launch();
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  trace_mark(p.trace, cta_lin, 0);
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15;
  const int k0 = split * p.chunk;
  if (p.prefetch && tid == 0) {
    // While the QKV GEMM drains: pull this split's K/V pages (16 KiB contiguous per page and head, written in earlier
    // steps) into L2.  The context length read here may still be the previous step's — a lower bound is all this needs.
    const int ctx0 = *reinterpret_cast<const volatile int32_t*>(p.ctx_len);
    if (k0 < ctx0) {
      const int nt = (min(p.chunk, ctx0 - k0) + kPT - 1) / kPT;
      const int32_t* bt0 = p.block_table + b * p.max_pages + k0 / kPT;
      for (int t = 0; t < nt; ++t) {
        const long off = (static_cast<long>(bt0[t]) * p.n_heads + h) * kPT * kDH;
        bulk_prefetch_l2(p.k_pool + off, kPT * kDH * 2);
        bulk_prefetch_l2(p.v_pool + off, kPT * kDH * 2);
      }
    }
  }
  griddep_wait();  // q and the current token's K/V row come from the kernels just before this one
  trace_mark(p.trace, cta_lin, 1);
  const int ctx = *p.ctx_len;
  const long pidx = (static_cast<long>(b) * p.n_heads + h) * splits + split;
  if (k0 >= ctx) {  // empty split of a graph sized for a longer context (never split 0: ctx >= 1)
    if constexpr (CLUSTER) {
      if (tid == 0) { s_fin[kDH] = -INFINITY; s_fin[kDH + 1] = 0.f; }
      cluster_sync_all();  // publish
      cluster_sync_all();  // CTA 0 has read
    } else {
      if (tid == 0) { p.part_ml[pidx * 2] = -INFINITY; p.part_ml[pidx * 2 + 1] = 0.f; }
      if (p.sync_counters != nullptr) merge_by_last_cta(p, splits, b, h, tid, &s_last);
    }
    return;
  }
  const int nkeys = min(p.chunk, ctx - k0);

  // Fused RoPE (p.cos_sin != null): q arrives un-rotated, and the current token's k / v are taken from the QKV row,
  // rotated here and appended to the paged cache by the one CTA whose split contains the current position
  // (modeling_llama.py:124-168 apply_rotary_pos_emb; positions = cache length, vcoder_ds_llava_llama.py:120-142).
  const bool fuse_rope = p.cos_sin != nullptr;
  const int pos = ctx - 1;
  float cs_c[8], cs_s[8];
  if (fuse_rope) {
    const float2* csr = p.cos_sin + static_cast<long>(pos) * 64 + (l16 & 7) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float2 v = __ldg(csr + e); cs_c[e] = v.x; cs_s[e] = v.y; }
  }
  // rotate-half on a 128-dim row spread over 16 lanes (8 dims each): lane l pairs with lane l^8
  auto rotate = [&](float* x) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float other = __shfl_xor_sync(0xffffffffu, x[e], 8);
      x[e] = (l16 < 8) ? (x[e] * cs_c[e] - other * cs_s[e]) : (x[e] * cs_c[e] + other * cs_s[e]);
    }
  };
  const __half* qkv_row = p.q + static_cast<long>(b) * p.q_stride;
  float qf[8];
  {
    const uint4 qv = *reinterpret_cast<const uint4*>(qkv_row + h * kDH + l16 * 8);
    unpack8(qv, qf);
    if (fuse_rope) rotate(qf);
    const float sc = p.scale * 1.4426950408889634f;  // scores in the log2 domain
#pragma unroll
    for (int i = 0; i < 8; ++i) qf[i] *= sc;
  }
  const int32_t* bt = p.block_table + b * p.max_pages + k0 / kPT;
  // where the current token lives inside this split (only meaningful when pos is in [k0, k0 + nkeys))
  const bool own_pos = fuse_rope && pos >= k0 && pos < k0 + p.chunk;
  const int pos_tile = (pos - k0) / kPT, pos_slot = (pos - k0) % kPT;

  float m = -INFINITY, l = 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int ntiles = (nkeys + kPT - 1) / kPT;
  for (int t = 0; t < ntiles; ++t) {
    const long page_off = (static_cast<long>(bt[t]) * p.n_heads + h) * kPT * kDH + l16 * 8;
    const __half* kb = p.k_pool + page_off;
    const __half* vb = p.v_pool + page_off;
    uint4 kv[8], vv[8];
    const bool fresh_tile = own_pos && t == pos_tile;  // CTA-uniform
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int slot = hw + 8 * i;
      const bool ok = t * kPT + slot < nkeys && !(fresh_tile && slot == pos_slot);
      kv[i] = ok ? ldg_stream16(kb + slot * kDH) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int slot = hw + 8 * i;
      const bool ok = t * kPT + slot < nkeys && !(fresh_tile && slot == pos_slot);
      vv[i] = ok ? ldg_stream16(vb + slot * kDH) : make_uint4(0, 0, 0, 0);
    }
    if (fresh_tile && (tid >> 5) == ((pos_slot & 7) >> 1)) {  // the warp that holds the owning half-warp (warp-uniform)
      // current token: k (rotated) and v come from the QKV row and are appended to the cache here
      const int H = p.n_heads * kDH;
      uint4 kraw = *reinterpret_cast<const uint4*>(qkv_row + H + h * kDH + l16 * 8);
      const uint4 vraw = *reinterpret_cast<const uint4*>(qkv_row + 2 * H + h * kDH + l16 * 8);
      float kf[8];
      unpack8(kraw, kf);
      rotate(kf);
      kraw = make_uint4(pack_half2(kf[0], kf[1]), pack_half2(kf[2], kf[3]), pack_half2(kf[4], kf[5]), pack_half2(kf[6], kf[7]));
      if (hw == (pos_slot & 7)) {
        const int i_own = pos_slot >> 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i == i_own) { kv[i] = kraw; vv[i] = vraw; }
        }
        *reinterpret_cast<uint4*>(p.k_pool_w + page_off + static_cast<long>(pos_slot) * kDH) = kraw;
        *reinterpret_cast<uint4*>(p.v_pool_w + page_off + static_cast<long>(pos_slot) * kDH) = vraw;
      }
    }
    float s[8];
    float mt = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float kf[8];
      unpack8(kv[i], kf);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qf[e], kf[e], d);
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      s[i] = (t * kPT + hw + 8 * i < nkeys) ? d : -INFINITY;
      mt = fmaxf(mt, s[i]);
    }
    const float m_new = fmaxf(m, mt);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float corr = exp2f(m - m_use);  // m = -inf -> 0
    l *= corr;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float pr = exp2f(s[i] - m_use);  // masked -> 0
      l += pr;
      float vf[8];
      unpack8(vv[i], vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, vf[e], acc[e]);
    }
    m = m_new;
  }

  // ---- merge the 8 half-warp states
#pragma unroll
  for (int e = 0; e < 8; ++e) s_acc[hw][l16 * 8 + e] = acc[e];
  if (l16 == 0) { s_m[hw] = m; s_l[hw] = l; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w]);
  float o = 0.f, L = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float wgt = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - M);
    o += wgt * s_acc[w][tid];
    L += wgt * s_l[w];
  }
  if constexpr (CLUSTER) {
    s_fin[tid] = o;
    if (tid == 0) { s_fin[kDH] = M; s_fin[kDH + 1] = L; }
    cluster_sync_all();
    if (split == 0) {  // cluster rank == blockIdx.x: the cluster spans the whole x dimension
      const uint32_t fin = smem_u32(s_fin);
      float Mg = M;
      for (int r = 1; r < splits; ++r) Mg = fmaxf(Mg, ld_dsmem_f32(dsmem_addr(fin + kDH * 4, r)));
      float num = o * exp2f(M - Mg), den = L * exp2f(M - Mg);  // split 0 is never empty: M is finite
      for (int r = 1; r < splits; ++r) {
        const float mr = ld_dsmem_f32(dsmem_addr(fin + kDH * 4, r));
        if (mr == -INFINITY) continue;
        const float w = exp2f(mr - Mg);
        num += w * ld_dsmem_f32(dsmem_addr(fin + tid * 4, r));
        den += w * ld_dsmem_f32(dsmem_addr(fin + (kDH + 1) * 4, r));
      }
      p.out[static_cast<long>(b) * p.n_heads * kDH + h * kDH + tid] = __float2half_rn(num / den);
    }
    cluster_sync_all();  // the other CTAs' shared memory stays alive until CTA 0 has read it
  } else {
    p.part_o[pidx * kDH + tid] = o;
    if (tid == 0) { p.part_ml[pidx * 2] = M; p.part_ml[pidx * 2 + 1] = L; }
    if (p.sync_counters != nullptr) merge_by_last_cta(p, splits, b, h, tid, &s_last);
  }
  trace_mark(p.trace, cta_lin, 2);
}

__global__ void __launch_bounds__(kDH)
attn_decode_combine_kernel(AttnDecodeParams p, int splits) {
  griddep_launch();
  trace_mark(p.trace2, blockIdx.x, 0);
  griddep_wait();
  trace_mark(p.trace2, blockIdx.x, 1);
  const int bh = blockIdx.x;
  const int b = bh / p.n_heads, h = bh % p.n_heads;
  const int d = threadIdx.x;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, p.part_ml[(static_cast<long>(bh) * splits + s) * 2]);
  float num = 0.f, den = 0.f;
  for (int s = 0; s < splits; ++s) {
    const long pi = static_cast<long>(bh) * splits + s;
    const float m = p.part_ml[pi * 2];
    if (m == -INFINITY) continue;
    const float w = exp2f(m - M);
    num += w * p.part_o[pi * kDH + d];
    den += w * p.part_ml[pi * 2 + 1];
  }
  p.out[static_cast<long>(b) * p.n_heads * kDH + h * kDH + d] = __float2half_rn(num / den);
  trace_mark(p.trace2, blockIdx.x, 2);
}


#ifdef VCB_EXPERIMENTAL
// ------------------------------------------------------------------------------------------------------------------
// Bulk-copy fed variant (default).  One CTA = (128-key split, head, sequence): the split's two K pages and two V pages
// (16 KiB contiguous each: one page of one head) are fetched by four cp.async.bulk copies issued BEFORE the
// grid-dependency wait — cached pages do not depend on the QKV GEMM that precedes this kernel — so the KV stream overlaps
// the predecessor's tail and the bytes in flight per SM (3 CTAs x 64 KiB) no longer depend on registers.  The half-warps
// then run the same online softmax as above out of shared memory.  The CTA that finishes LAST for a (sequence, head)
// merges the partial states (fixed split order: deterministic) and writes the fp16 output: no combine kernel.
constexpr int kChunkT = 128;
constexpr int kPagesT = kChunkT / kPT;        // 2
constexpr int kPageBytes = kPT * kDH * 2;     // 16 KiB
constexpr int kSmemT = 2 * kPagesT * kPageBytes + 64;

__global__ void __launch_bounds__(kDecThreads, 3)
attn_decode_tma_kernel(AttnDecodeParams p, int splits) {
  extern __shared__ __align__(128) uint8_t dsm[];  // K pages | V pages | barriers
  __shared__ float s_acc[8][kDH];
  __shared__ float s_m[8], s_l[8];
  __shared__ int s_last;
  griddep_launch();
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  trace_mark(p.trace, cta_lin, 0);
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15;
  const int k0 = split * kChunkT;
  const uint32_t sbase = smem_u32(dsm);
  const uint32_t bars = sbase + 2 * kPagesT * kPageBytes;
  const int32_t* bt = p.block_table + b * p.max_pages + k0 / kPT;
  // Context length as of the previous step (the counters are advanced at the very end of a step, long before this
  // kernel starts): a lower bound of this step's length, which is all the prefetch decision needs
  const int ctx_lo = *reinterpret_cast<const volatile int32_t*>(p.ctx_len);
  int issued = 0;  // pages whose copies are in flight (thread 0 only)
  if (tid == 0) {
    for (int t = 0; t < kPagesT; ++t) mbar_init(bars + 8 * t, 1);
    fence_mbar_init();
    // (without the fused RoPE / append, the current token's row is written by the rope kernel right before this one:
    //  nothing may be fetched ahead of the wait then)
    for (int t = 0; t < kPagesT && p.cos_sin != nullptr; ++t) {
      if (k0 + t * kPT < ctx_lo) {
        const long off = (static_cast<long>(bt[t]) * p.n_heads + h) * kPT * kDH;
        mbar_expect_tx(bars + 8 * t, 2 * kPageBytes);
        bulk_load_g2s(sbase + t * kPageBytes, p.k_pool + off, kPageBytes, bars + 8 * t);
        bulk_load_g2s(sbase + (kPagesT + t) * kPageBytes, p.v_pool + off, kPageBytes, bars + 8 * t);
        issued = t + 1;
      }
    }
  }
  griddep_wait();  // q and the current token's K/V row come from the kernels just before this one
  trace_mark(p.trace, cta_lin, 1);
  const int ctx = *p.ctx_len;
  const long pidx = (static_cast<long>(b) * p.n_heads + h) * splits + split;
  const int nkeys = min(kChunkT, ctx - k0);  // <= 0: empty split of a graph sized for a longer context
  float M = -INFINITY, Lsum = 0.f, o = 0.f;
  if (nkeys > 0) {
    const int ntiles = (nkeys + kPT - 1) / kPT;
    if (tid == 0) {
      // a page that only holds the current token (ctx crossed a page boundary in this step) was not prefetched: nothing
      // cached to read there, but the barrier must still complete
      for (int t = issued; t < ntiles; ++t) {
        const long off = (static_cast<long>(bt[t]) * p.n_heads + h) * kPT * kDH;
        mbar_expect_tx(bars + 8 * t, 2 * kPageBytes);
        bulk_load_g2s(sbase + t * kPageBytes, p.k_pool + off, kPageBytes, bars + 8 * t);
        bulk_load_g2s(sbase + (kPagesT + t) * kPageBytes, p.v_pool + off, kPageBytes, bars + 8 * t);
      }
    }
    const bool fuse_rope = p.cos_sin != nullptr;
    const int pos = ctx - 1;
    float cs_c[8], cs_s[8];
    if (fuse_rope) {
      const float2* csr = p.cos_sin + static_cast<long>(pos) * 64 + (l16 & 7) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float2 v = __ldg(csr + e); cs_c[e] = v.x; cs_s[e] = v.y; }
    }
    auto rotate = [&](float* x) {  // rotate-half on a 128-dim row spread over 16 lanes: lane l pairs with lane l^8
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float other = __shfl_xor_sync(0xffffffffu, x[e], 8);
        x[e] = (l16 < 8) ? (x[e] * cs_c[e] - other * cs_s[e]) : (x[e] * cs_c[e] + other * cs_s[e]);
      }
    };
    const __half* qkv_row = p.q + static_cast<long>(b) * p.q_stride;
    float qf[8];
    {
      const uint4 qv = *reinterpret_cast<const uint4*>(qkv_row + h * kDH + l16 * 8);
      unpack8(qv, qf);
      if (fuse_rope) rotate(qf);
      const float sc = p.scale * 1.4426950408889634f;  // scores in the log2 domain
#pragma unroll
      for (int i = 0; i < 8; ++i) qf[i] *= sc;
    }
    const bool own_pos = fuse_rope && pos >= k0 && pos < k0 + kChunkT;
    const int pos_tile = (pos - k0) / kPT, pos_slot = (pos - k0) % kPT;
    // current token (fused mode): rotated k and v come from the QKV row; every half-warp computes them (cheap, keeps the
    // shuffles warp-uniform), the owner of the slot substitutes them for the stale cache row and appends them to the pool
    uint4 kcur = make_uint4(0, 0, 0, 0), vcur = make_uint4(0, 0, 0, 0);
    if (own_pos) {
      const int H = p.n_heads * kDH;
      kcur = *reinterpret_cast<const uint4*>(qkv_row + H + h * kDH + l16 * 8);
      vcur = *reinterpret_cast<const uint4*>(qkv_row + 2 * H + h * kDH + l16 * 8);
      float kf[8];
      unpack8(kcur, kf);
      rotate(kf);
      kcur = make_uint4(pack_half2(kf[0], kf[1]), pack_half2(kf[2], kf[3]), pack_half2(kf[4], kf[5]), pack_half2(kf[6], kf[7]));
      if (hw == (pos_slot & 7)) {
        const long off = (static_cast<long>(bt[pos_tile]) * p.n_heads + h) * kPT * kDH + static_cast<long>(pos_slot) * kDH + l16 * 8;
        *reinterpret_cast<uint4*>(p.k_pool_w + off) = kcur;
        *reinterpret_cast<uint4*>(p.v_pool_w + off) = vcur;
      }
    }
    float m = -INFINITY, l = 0.f;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(bars + 8 * t, 0);
      const uint8_t* kp = dsm + t * kPageBytes + l16 * 16;
      const uint8_t* vp = dsm + (kPagesT + t) * kPageBytes + l16 * 16;
      const bool fresh_tile = own_pos && t == pos_tile;
      float s[8];
      float mt = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int slot = hw + 8 * i;
        uint4 kv = *reinterpret_cast<const uint4*>(kp + slot * (kDH * 2));
        if (fresh_tile && slot == pos_slot) kv = kcur;
        float kf[8];
        unpack8(kv, kf);
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d = fmaf(qf[e], kf[e], d);
        d += __shfl_xor_sync(0xffffffffu, d, 8);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        s[i] = (t * kPT + slot < nkeys) ? d : -INFINITY;  // stale rows behind the context may hold anything: index-masked
        mt = fmaxf(mt, s[i]);
      }
      const float m_new = fmaxf(m, mt);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m - m_use);  // m = -inf -> 0
      l *= corr;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int slot = hw + 8 * i;
        if (t * kPT + slot < nkeys) {  // half-warp uniform; a masked row's V must not be touched (0 * NaN)
          const float pr = exp2f(s[i] - m_use);
          l += pr;
          uint4 vv = *reinterpret_cast<const uint4*>(vp + slot * (kDH * 2));
          if (fresh_tile && slot == pos_slot) vv = vcur;
          float vf[8];
          unpack8(vv, vf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, vf[e], acc[e]);
        }
      }
      m = m_new;
    }
    // ---- merge the 8 half-warp states
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[hw][l16 * 8 + e] = acc[e];
    if (l16 == 0) { s_m[hw] = m; s_l[hw] = l; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w]);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float wgt = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - M);
      o += wgt * s_acc[w][tid];
      Lsum += wgt * s_l[w];
    }
    p.part_o[pidx * kDH + tid] = o;
  }
  if (tid == 0) { p.part_ml[pidx * 2] = M; p.part_ml[pidx * 2 + 1] = Lsum; }
  merge_by_last_cta(p, splits, b, h, tid, &s_last);
  trace_mark(p.trace, cta_lin, 2);
}

#endif  // VCB_EXPERIMENTAL

}  // namespace

namespace {
// VCB_ATTN_DEC: default "stream" = the balanced page-streaming kernel of attn_decode_stream.cu (+ its combine kernel).
// "split": the register-staged split kernel of this file + combine kernel (round 1).  "merge": the same split kernel whose last CTA per
// (sequence, head) merges the partials (threadfence + ticket): measured SLOWER (38.9 vs 33.3 + 2.4 + 0.6 us per layer at
// B = 8, ctx 1271; profiles/r02_trace_decode_merge_last_cta.csv) — the fence + atomic tail of 1536 CTAs costs more than one
// small launch.  "tma": the bulk-copy fed kernel, also slower (46 us: a CTA loads its 64 KiB, computes and leaves, so its
// shared memory carries no loads during the tail).
int decode_attn_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VCB_ATTN_DEC");
    v = (e == nullptr) ? 3 : (e[0] == 't' ? 2 : (e[0] == 'm' ? 0 : (e[0] == 's' && e[1] == 'p' ? 1 : 3)));
  }
  return v;
}
bool decode_attn_ldg() { return decode_attn_impl() != 2; }
int decode_attn_chunk() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("VCB_ATTN_DEC_CHUNK");
    const int c = e ? atoi(e) : kChunk;
    v = (c == 128 || c == 512) ? c : kChunk;
  }
  return v;
}
int decode_attn_minb() {  // resident CTAs per SM the split kernel is compiled for (register cap 168 vs 128)
  static int v = 0;
  if (v == 0) { const char* e = getenv("VCB_ATTN_DEC_MINB"); v = (e && atoi(e) == 4) ? 4 : 3; }
  return v;
}
}  // namespace

int attn_decode_splits(int max_ctx) {
#ifdef VCB_EXPERIMENTAL
  const int chunk = decode_attn_ldg() ? decode_attn_chunk() : kChunkT;
#else
  const int chunk = decode_attn_chunk();
#endif
  return (max_ctx + chunk - 1) / chunk;
}

int attn_llama_decode(const AttnDecodeParams& p_in, cudaStream_t stream) {
  AttnDecodeParams p = p_in;
  static int env_pf = -1;  // VCB_ATTN_DEC_PF=0 turns off: L2-prefetch the split's KV pages before the grid-dependency wait
  if (env_pf < 0) { const char* e = getenv("VCB_ATTN_DEC_PF"); env_pf = e ? atoi(e) : 1; }  // measured: DESIGN.md section 3.1
  p.prefetch = env_pf;
  p.chunk = decode_attn_chunk();
  if (decode_attn_impl() == 3) return attn_llama_decode_stream(p_in, stream);
  const int splits = attn_decode_splits(p.max_ctx);
  VCB_CHECK(splits >= 1, kErrInvalid, "attn_llama_decode: empty context");
  VCB_CHECK(p.page_tokens == kPT, kErrInvalid, "attn_llama_decode: KV page must hold 64 tokens");
  dim3 grid(splits, p.n_heads, p.batch);
#ifdef VCB_EXPERIMENTAL
  if (!decode_attn_ldg()) {
    VCB_CHECK(p.sync_counters != nullptr, kErrInvalid, "attn_llama_decode: sync_counters missing");
    static bool attr = false;
    if (!attr) {
      VCB_CUDA(cudaFuncSetAttribute(attn_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT));
      attr = true;
    }
    p.trace = trace_slots(kTraceAttnSplit, splits * p.n_heads * p.batch, splits);
    KernelScope scope(kCatAttnDecode, stream, 1);
    VCB_CUDA(launch_pdl(attn_decode_tma_kernel, grid, dim3(kDecThreads), kSmemT, stream, p, splits));
    return kOk;
  }
  // VCB_ATTN_DEC_CLUSTER=1: one kernel, partials merged through DSMEM inside a cluster of `splits` CTAs.  Measured SLOWER on
  // B200 (profiles/r02_trace_decode_*.csv: 48.7 us vs 33.3 + 2.4 us per layer at B = 8, ctx 1271): co-scheduling six
  // 128-thread CTAs per cluster costs more residency than the combine launch saves.  Off.
  static int env_cl = -1;
  if (env_cl < 0) { const char* e = getenv("VCB_ATTN_DEC_CLUSTER"); env_cl = e ? atoi(e) : 0; }
  p.trace = trace_slots(kTraceAttnSplit, splits * p.n_heads * p.batch, splits);
  if (env_cl != 0 && splits <= 8) {  // portable cluster size
    KernelScope scope(kCatAttnDecode, stream, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kDecThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = static_cast<unsigned>(splits);
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    VCB_CUDA(cudaLaunchKernelEx(&cfg, attn_decode_split_kernel<true, 3>, p, splits));
    return kOk;
  }
#endif  // VCB_EXPERIMENTAL (bulk-copy fed and cluster variants: measured slower)
  auto split_kernel = decode_attn_minb() == 4 ? attn_decode_split_kernel<false, 4> : attn_decode_split_kernel<false, 3>;
  if (decode_attn_impl() == 0 && p.sync_counters != nullptr) {  // one kernel: the last CTA of a (sequence, head) merges
    KernelScope scope(kCatAttnDecode, stream, 1);
    VCB_CUDA(launch_pdl(split_kernel, grid, dim3(kDecThreads), 0, stream, p, splits));
    return kOk;
  }
  p.sync_counters = nullptr;
  p.trace2 = trace_slots(kTraceAttnCombine, p.batch * p.n_heads, splits);
  KernelScope scope(kCatAttnDecode, stream, 2);
  VCB_CUDA(launch_pdl(split_kernel, grid, dim3(kDecThreads), 0, stream, p, splits));
  VCB_CUDA(launch_pdl(attn_decode_combine_kernel, dim3(p.batch * p.n_heads), dim3(kDH), 0, stream, p, splits));
  return kOk;
}

}  // namespace vcb
