// Decode attention over the paged KV cache, balanced streaming form (flash-decoding with an even split of the PAGES).
// Replaces HF eager attention at decode: softmax_fp32(q·K^T / sqrt(128)) · V over EVERY cached position — the reference
// substitutes an all-ones mask of length past+1 on each decode step (vcoder_llava/model/vcoder_ds_llava_arch.py:130-133;
// arithmetic modeling_llama.py:199-222) — plus RoPE and the KV append of the current token (K10).
//
// HBM-bound: B·ctx·2·H·2 bytes per layer per step.  The split kernel of attn_decode.cu launches one CTA per (256-key
// split, head, sequence): 1536 CTAs on ~444 resident slots = 2.9 waves whose tail and whose per-CTA start-up (load ->
// wait -> compute, nothing in flight while a CTA winds down) leave it at 5.0 TB/s (33-35 us per layer at B = 8, ctx 1271
// against 25.5 us at the copy peak; profiles/r02_trace_decode_stream.csv).  Here the work is the flat list of KV pages
// [sequence][head][page] and every CTA of ONE resident wave takes an equal, contiguous share of it (about 11.5 pages),
// streaming them through a two-page ring of cp.async.bulk copies (one page of one head = 16 KiB contiguous for K and for
// V) that never drains between pages.  A share crosses at most one (sequence, head) boundary per `pps` pages; a CTA emits
// one partial (max, sum, acc) state per (sequence, head) it touches, and a small combine kernel merges the 1-3 partials of
// every (sequence, head) in share order (deterministic).
//   * fused RoPE + append (decode step): q and the current token's k are rotated here; the CTA that owns the current
//     position's page substitutes the fresh k / v for the stale cache row and appends them to the pool.
//   * programmatic dependent launch: the first two pages of a share are requested BEFORE the grid-dependency wait (cached
//     pages do not depend on the QKV GEMM), except on the one step in 64 where the context enters a new page.
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kPT = 64;                    // tokens per KV page (checked on the host)
constexpr int kDH = 128;
constexpr int kThreads = 128;
constexpr int kRing = 2;                   // pages in flight per CTA
constexpr int kPageBytes = kPT * kDH * 2;  // 16 KiB
constexpr int kSmem = 2 * kRing * kPageBytes + 64;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// share of CTA c: flat pages [c * ppc, min(total, (c + 1) * ppc)), flat page P = (b * heads + h) * pps + page
struct Shares {
  int pps, total, ppc;
  __device__ Shares(int ctx, int bh_count, int n_ctas) {
    pps = (ctx + kPT - 1) / kPT;
    total = bh_count * pps;
    ppc = (total + n_ctas - 1) / n_ctas;
  }
};

__global__ void __launch_bounds__(kThreads, 3)
attn_decode_stream_kernel(AttnDecodeParams p, int max_slots) {
  extern __shared__ __align__(128) uint8_t dsm[];  // K pages [kRing] | V pages [kRing] | barriers
  __shared__ float s_acc[8][kDH];
  __shared__ float s_m[8], s_l[8];
  griddep_launch();
  trace_mark(p.trace, blockIdx.x, 0);
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15;
  const int c = blockIdx.x, G = gridDim.x;
  const int BH = p.batch * p.n_heads;
  const uint32_t sbase = smem_u32(dsm);
  const uint32_t bars = sbase + 2 * kRing * kPageBytes;
  const bool fuse_rope = p.cos_sin != nullptr;

  auto page_src = [&](int P, int pps) -> long {  // element offset of flat page P in the pools
    const int bh = P / pps, pg = P - bh * pps;
    const int b = bh / p.n_heads, h = bh - b * p.n_heads;
    return (static_cast<long>(p.block_table[b * p.max_pages + pg]) * p.n_heads + h) * kPT * kDH;
  };
  auto issue = [&](int i, int P, int pps) {  // i-th page of this share -> ring slot i % kRing
    const int s = i % kRing;
    const long off = page_src(P, pps);
    mbar_expect_tx(bars + 8 * s, 2 * kPageBytes);
    bulk_load_g2s(sbase + s * kPageBytes, p.k_pool + off, kPageBytes, bars + 8 * s);
    bulk_load_g2s(sbase + (kRing + s) * kPageBytes, p.v_pool + off, kPageBytes, bars + 8 * s);
  };

  // ---- before the grid-dependency wait: barriers, and (usually) the first pages of the share
  const int ctx_lo = *reinterpret_cast<const volatile int32_t*>(p.ctx_len);  // this step's length or the previous step's
  bool prefetched = false;
  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) mbar_init(bars + 8 * s, 1);
    fence_mbar_init();
    // the share boundaries depend on ceil(ctx / 64): they are final unless this step may enter a new page
    if (fuse_rope && ctx_lo >= 1 && (ctx_lo % kPT) != 0) {
      const Shares sh(ctx_lo, BH, G);
      const int P0 = c * sh.ppc, P1 = min(sh.total, P0 + sh.ppc);
      for (int i = 0; i < kRing && P0 + i < P1; ++i) issue(i, P0 + i, sh.pps);
      prefetched = true;
    }
  }
  griddep_wait();  // q and the current token's K/V row come from the kernels just before this one
  trace_mark(p.trace, blockIdx.x, 1);
  const int ctx = *p.ctx_len;
  const Shares sh(ctx, BH, G);
  const int P0 = c * sh.ppc, P1 = min(sh.total, P0 + sh.ppc);
  const int n_my = P1 - P0;
  if (n_my <= 0) {
    trace_mark(p.trace, blockIdx.x, 2);
    return;  // more CTAs than pages (tiny contexts): nothing to do, the combine kernel knows
  }
  if (tid == 0 && !prefetched)
    for (int i = 0; i < kRing && i < n_my; ++i) issue(i, P0 + i, sh.pps);

  const int pos = ctx - 1;
  const int pos_pg = pos / kPT, pos_slot = pos % kPT;
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

  int cur_bh = -1;
  float qf[8];
  uint4 kcur = make_uint4(0, 0, 0, 0), vcur = make_uint4(0, 0, 0, 0);
  float m = -INFINITY, l = 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  // partial state of (sequence, head) `bh` -> slot (c - first CTA that touches bh)
  auto flush = [&](int bh) {
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
    const int slot = c - (bh * sh.pps) / sh.ppc;
    const long pi = static_cast<long>(bh) * max_slots + slot;
    p.part_o[pi * kDH + tid] = o;
    if (tid == 0) { p.part_ml[pi * 2] = M; p.part_ml[pi * 2 + 1] = L; }
    __syncthreads();  // s_acc / s_m are reused by the next (sequence, head) of this share
  };

  for (int i = 0; i < n_my; ++i) {
    const int P = P0 + i;
    const int bh = P / sh.pps, pg = P - bh * sh.pps;
    if (bh != cur_bh) {  // CTA-uniform
      if (cur_bh >= 0) flush(cur_bh);
      cur_bh = bh;
      m = -INFINITY; l = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      const int b = bh / p.n_heads, h = bh - b * p.n_heads;
      const __half* qkv_row = p.q + static_cast<long>(b) * p.q_stride;
      const uint4 qv = *reinterpret_cast<const uint4*>(qkv_row + h * kDH + l16 * 8);
      unpack8(qv, qf);
      if (fuse_rope) rotate(qf);
      const float sc = p.scale * 1.4426950408889634f;  // scores in the log2 domain
#pragma unroll
      for (int e = 0; e < 8; ++e) qf[e] *= sc;
      if (fuse_rope) {
        // current token: rotated k and v from the QKV row (every half-warp computes them: keeps the shuffles uniform)
        const int H = p.n_heads * kDH;
        kcur = *reinterpret_cast<const uint4*>(qkv_row + H + h * kDH + l16 * 8);
        vcur = *reinterpret_cast<const uint4*>(qkv_row + 2 * H + h * kDH + l16 * 8);
        float kf[8];
        unpack8(kcur, kf);
        rotate(kf);
        kcur = make_uint4(pack_half2(kf[0], kf[1]), pack_half2(kf[2], kf[3]), pack_half2(kf[4], kf[5]), pack_half2(kf[6], kf[7]));
      }
    }
    const bool fresh = fuse_rope && pg == pos_pg;  // this page holds the current position
    if (fresh && hw == (pos_slot & 7)) {            // append the current token's k / v to the pool (once per (b, h))
      const int b = bh / p.n_heads, h = bh - b * p.n_heads;
      const long off = (static_cast<long>(p.block_table[b * p.max_pages + pg]) * p.n_heads + h) * kPT * kDH +
                       static_cast<long>(pos_slot) * kDH + l16 * 8;
      *reinterpret_cast<uint4*>(p.k_pool_w + off) = kcur;
      *reinterpret_cast<uint4*>(p.v_pool_w + off) = vcur;
    }
    const int s = i % kRing;
    mbar_wait(bars + 8 * s, static_cast<uint32_t>((i / kRing) & 1));
    const uint8_t* kp = dsm + s * kPageBytes + l16 * 16;
    const uint8_t* vp = dsm + (kRing + s) * kPageBytes + l16 * 16;
    const int nkeys = min(kPT, ctx - pg * kPT);  // valid keys of this page (>= 1)
    float sv[8];
    float mt = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = hw + 8 * j;
      uint4 kv = *reinterpret_cast<const uint4*>(kp + slot * (kDH * 2));
      if (fresh && slot == pos_slot) kv = kcur;
      float kf[8];
      unpack8(kv, kf);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qf[e], kf[e], d);
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      sv[j] = (slot < nkeys) ? d : -INFINITY;  // rows behind the context may hold anything: masked by index
      mt = fmaxf(mt, sv[j]);
    }
    const float m_new = fmaxf(m, mt);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float corr = exp2f(m - m_use);  // m = -inf -> 0
    l *= corr;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = hw + 8 * j;
      if (slot < nkeys) {  // half-warp uniform; a masked row's V must not be touched (0 * NaN)
        const float pr = exp2f(sv[j] - m_use);
        l += pr;
        uint4 vv = *reinterpret_cast<const uint4*>(vp + slot * (kDH * 2));
        if (fresh && slot == pos_slot) vv = vcur;
        float vf[8];
        unpack8(vv, vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, vf[e], acc[e]);
      }
    }
    m = m_new;
    __syncthreads();  // every thread is done with ring slot s before it is refilled
    if (tid == 0 && i + kRing < n_my) issue(i + kRing, P + kRing, sh.pps);
  }
  flush(cur_bh);
  trace_mark(p.trace, blockIdx.x, 2);
}

// merges the partials of one (sequence, head): the CTAs first..last of the stream kernel, in that order
__global__ void __launch_bounds__(kDH)
attn_decode_stream_combine_kernel(AttnDecodeParams p, int max_slots, int n_ctas) {
  griddep_launch();
  trace_mark(p.trace2, blockIdx.x, 0);
  griddep_wait();
  trace_mark(p.trace2, blockIdx.x, 1);
  const int bh = blockIdx.x, d = threadIdx.x;
  const Shares sh(*p.ctx_len, p.batch * p.n_heads, n_ctas);
  const int first = (bh * sh.pps) / sh.ppc, last = ((bh + 1) * sh.pps - 1) / sh.ppc;
  const int n = last - first + 1;
  const long base = static_cast<long>(bh) * max_slots;
  float M = -INFINITY;
  for (int s = 0; s < n; ++s) M = fmaxf(M, p.part_ml[(base + s) * 2]);
  float num = 0.f, den = 0.f;
  for (int s = 0; s < n; ++s) {
    const float mm = p.part_ml[(base + s) * 2];
    if (mm == -INFINITY) continue;
    const float w = exp2f(mm - M);
    num += w * p.part_o[(base + s) * kDH + d];
    den += w * p.part_ml[(base + s) * 2 + 1];
  }
  p.out[static_cast<long>(bh) * kDH + d] = __float2half_rn(num / den);
  trace_mark(p.trace2, blockIdx.x, 2);
}

int stream_grid() {  // one resident wave
  static int g = 0;
  if (g == 0) {
    cudaFuncSetAttribute(attn_decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, attn_decode_stream_kernel, kThreads, kSmem) != cudaSuccess || occ < 1) occ = 1;
    g = occ * device_sm_count();
  }
  return g;
}

}  // namespace

// partial-state slots per (sequence, head): a share is at least total/G pages, so at most G/BH + 2 shares touch one
int attn_decode_stream_slots(int batch, int n_heads) { return stream_grid() / (batch * n_heads) + 2; }

int attn_llama_decode_stream(const AttnDecodeParams& p_in, cudaStream_t stream) {
  AttnDecodeParams p = p_in;
  VCB_CHECK(p.page_tokens == kPT, kErrInvalid, "attn_llama_decode_stream: KV page must hold 64 tokens");
  const int G = stream_grid();
  const int max_slots = attn_decode_stream_slots(p.batch, p.n_heads);
  p.trace = trace_slots(kTraceAttnSplit, G, max_slots);
  p.trace2 = trace_slots(kTraceAttnCombine, p.batch * p.n_heads, max_slots);
  KernelScope scope(kCatAttnDecode, stream, 2);
  VCB_CUDA(launch_pdl(attn_decode_stream_kernel, dim3(G), dim3(kThreads), kSmem, stream, p, max_slots));
  VCB_CUDA(launch_pdl(attn_decode_stream_combine_kernel, dim3(p.batch * p.n_heads), dim3(kDH), 0, stream, p, max_slots, G));
  return kOk;
}

}  // namespace vcb
