// Decode attention over the paged KV cache (flash-decoding: split along the context, then combine).
// Replaces HF eager attention at decode: softmax_fp32(q·K^T / sqrt(128)) · V over EVERY cached position — the
// reference substitutes an all-ones mask of length past+1 on each decode step
// (vcoder_llava/model/vcoder_ds_llava_arch.py:130-133; arithmetic modeling_llama.py:199-222).
// HBM-bound: each K/V byte of the (b, head) stream is read once with 16-byte loads; 16 lanes cover one 128-dim row.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kChunk = 256;  // keys per split
constexpr int kDecThreads = 128;
constexpr int kDH = 128;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__global__ void __launch_bounds__(kDecThreads)
attn_decode_split_kernel(AttnDecodeParams p, int splits) {
  __shared__ float s_p[kChunk];
  __shared__ float s_red[8][kDH];
  __shared__ float s_wred[4];
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15, lane = tid & 31, warp = tid >> 5;
  const int ctx = *p.ctx_len;
  const int k0 = split * kChunk;
  const long pidx = (static_cast<long>(b) * p.n_heads + h) * splits + split;
  if (k0 >= ctx) {
    if (tid == 0) { p.part_ml[pidx * 2] = -INFINITY; p.part_ml[pidx * 2 + 1] = 0.f; }
    return;
  }
  const int kend = min(k0 + kChunk, ctx);
  const int nkeys = kend - k0;

  float qf[8];
  {
    const uint4 qv = *reinterpret_cast<const uint4*>(p.q + static_cast<long>(b) * p.q_stride + h * kDH + l16 * 8);
    unpack8(qv, qf);
    const float sc = p.scale * 1.4426950408889634f;
#pragma unroll
    for (int i = 0; i < 8; ++i) qf[i] *= sc;
  }
  const int32_t* bt = p.block_table + b * p.max_pages;
  const int PT = p.page_tokens;
  auto row_ptr = [&](const __half* pool, int key) -> const uint4* {
    const int page = bt[key / PT];
    const int slot = key % PT;
    return reinterpret_cast<const uint4*>(pool + ((static_cast<long>(page) * p.n_heads + h) * PT + slot) * kDH + l16 * 8);
  };

  // ---- phase 1: scores (log2 domain) -> smem
  for (int base = 0; base < nkeys; base += 32) {  // warp-uniform trip count (shuffles inside)
    const int kb = base + hw;
    uint4 kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = kb + u * 8;
      kv[u] = (kk < nkeys) ? __ldg(row_ptr(p.k_pool, k0 + kk)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float kf[8];
      unpack8(kv[u], kf);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(qf[i], kf[i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      const int kk = kb + u * 8;
      if (l16 == 0 && kk < nkeys) s_p[kk] = d;
    }
  }
  __syncthreads();

  // ---- phase 2: chunk max, probabilities, chunk sum
  float m = -INFINITY;
  for (int i = tid; i < nkeys; i += kDecThreads) m = fmaxf(m, s_p[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_wred[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(s_wred[0], s_wred[1]), fmaxf(s_wred[2], s_wred[3]));
  __syncthreads();
  float l = 0.f;
  for (int i = tid; i < nkeys; i += kDecThreads) {
    const float e = exp2f(s_p[i] - m);
    s_p[i] = e;
    l += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) s_wred[warp] = l;
  __syncthreads();
  l = s_wred[0] + s_wred[1] + s_wred[2] + s_wred[3];

  // ---- phase 3: o = sum_k p_k V_k  (each half-warp takes every 8th key, lanes own 8 dims)
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < nkeys; base += 32) {
    const int kb = base + hw;
    uint4 vv[4];
    float pr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = kb + u * 8;
      const bool ok = kk < nkeys;
      vv[u] = ok ? __ldg(row_ptr(p.v_pool, k0 + kk)) : make_uint4(0, 0, 0, 0);
      pr[u] = ok ? s_p[kk] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float vf[8];
      unpack8(vv[u], vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr[u], vf[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_red[hw][l16 * 8 + i] = acc[i];
  __syncthreads();
  {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_red[w][tid];
    p.part_o[pidx * kDH + tid] = v;
  }
  if (tid == 0) { p.part_ml[pidx * 2] = m; p.part_ml[pidx * 2 + 1] = l; }
}

__global__ void __launch_bounds__(kDH)
attn_decode_combine_kernel(AttnDecodeParams p, int splits) {
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
}

}  // namespace

int attn_decode_splits(int max_ctx) { return (max_ctx + kChunk - 1) / kChunk; }

int attn_llama_decode(const AttnDecodeParams& p, cudaStream_t stream) {
  const int splits = attn_decode_splits(p.max_ctx);
  VCB_CHECK(splits >= 1, kErrInvalid, "attn_llama_decode: empty context");
  dim3 grid(splits, p.n_heads, p.batch);
  KernelScope scope(kCatAttnDecode, stream, 2);
  attn_decode_split_kernel<<<grid, kDecThreads, 0, stream>>>(p, splits);
  VCB_CUDA(cudaGetLastError());
  attn_decode_combine_kernel<<<p.batch * p.n_heads, kDH, 0, stream>>>(p, splits);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb
