// HBM-bound helpers around the GEMMs: im2col for the CLIP patch embedding, LayerNorm / RMSNorm, RoPE + paged-KV
// append, the visual/text splice gather, embedding gather, argmax, CLS-row compaction.  All 16-byte vectorised,
// fp32 internal math, fp16 in/out.  Reference call sites are cited per kernel.
#include <cmath>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) t += scratch[i];
  return t;
}

// ---- CLIPVisionEmbeddings.patch_embedding as a GEMM: im2col with K order (c, ky, kx), zero-padded to k_pad
//      (transformers/models/clip/modeling_clip.py:150-156, :207-208)
__global__ void im2col_kernel(const __half* __restrict__ px, __half* __restrict__ out, int image, int patch, int npr,
                              int k_real, int k_pad) {
  const int pidx = blockIdx.x;  // patch within image
  const int b = blockIdx.y;
  const int py = pidx / npr, pxx = pidx % npr;
  __half* orow = out + (static_cast<long>(b) * npr * npr + pidx) * k_pad;
  const __half* img = px + static_cast<long>(b) * 3 * image * image;
  const int pp = patch * patch;
  for (int k = threadIdx.x; k < k_pad; k += blockDim.x) {
    __half v = __float2half(0.f);
    if (k < k_real) {
      const int c = k / pp, r = k % pp;
      const int ky = r / patch, kx = r % patch;
      v = img[(static_cast<long>(c) * image + py * patch + ky) * image + pxx * patch + kx];
    }
    orow[k] = v;
  }
}

// ---- CLS row: class_embedding + position_embedding[0]  (modeling_clip.py:210-215)
__global__ void cls_rows_kernel(const __half* __restrict__ cls, const __half* __restrict__ pos, __half* hidden,
                                int tokens, int D) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    hidden[static_cast<long>(b) * tokens * D + d] = __float2half_rn(__half2float(cls[d]) + __half2float(pos[d]));
  }
}

// ---- LayerNorm (pre_layrnorm, layer_norm1/2: nn.LayerNorm eps 1e-5). One warp per row, two-pass.
constexpr int kLnMaxChunks = 8;  // D <= 2048
__global__ void __launch_bounds__(128)
layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bvec,
                 __half* __restrict__ y, long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nch = D / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
  float v[kLnMaxChunks][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      unpack8(xr[c], v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  const float mean = warp_sum(sum) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * D);
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float wv[8], bv[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), wv);
      unpack8(__ldg(reinterpret_cast<const uint4*>(bvec) + c), bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * wv[j] + bv[j];
      yr[c] = pack8(o);
    }
  }
}

// ---- LlamaRMSNorm: fp32 x * rsqrt(mean(x^2)+eps) * w   (modeling_llama.py:53-70). One CTA per row.
constexpr int kRmsThreads = 128;
constexpr int kRmsMaxChunks = 8;  // H <= 8192
__global__ void __launch_bounds__(kRmsThreads)
rmsnorm_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ y, int H, float eps,
               unsigned long long* trace) {
  __shared__ float scratch[kRmsThreads / 32];
  griddep_launch();
  trace_mark(trace, blockIdx.x, 0);
  const long row = blockIdx.x;
  const int nch = H / 8;
  // the norm weight does not depend on the previous kernel: fetch it before the grid-dependency wait so that only ONE
  // dependent global round trip (the activations) sits on the decode critical path
  uint4 wq[kRmsMaxChunks];
#pragma unroll
  for (int i = 0; i < kRmsMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRmsThreads;
    wq[i] = (c < nch) ? __ldg(reinterpret_cast<const uint4*>(w) + c) : make_uint4(0, 0, 0, 0);
  }
  griddep_wait();
  trace_mark(trace, blockIdx.x, 1);
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * H);
  uint4 xq[kRmsMaxChunks];
#pragma unroll
  for (int i = 0; i < kRmsMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRmsThreads;
    xq[i] = (c < nch) ? xr[c] : make_uint4(0, 0, 0, 0);
  }
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kRmsMaxChunks; ++i) {
    float v[8];
    unpack8(xq[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
  }
  const float rstd = rsqrtf(block_sum<kRmsThreads>(sq, scratch) / H + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * H);
#pragma unroll
  for (int i = 0; i < kRmsMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRmsThreads;
    if (c < nch) {
      float v[8], wv[8], o[8];
      unpack8(xq[i], v);
      unpack8(wq[i], wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[j] * rstd * wv[j];
      yr[c] = pack8(o);
    }
  }
  trace_mark(trace, blockIdx.x, 2);
}

// ---- RoPE (rotate-half, theta 1e4, positions arange(past, past+T)) + append K/V to the paged cache
//      (modeling_llama.py:124-168 apply_rotary_pos_emb; reference passes no position_ids:
//       vcoder_ds_llava_llama.py:120-142).  One CTA per token row; 8 threads per head.
__global__ void rope_kv_kernel(__half* qkv, int T, int H, int n_heads, const float2* __restrict__ cs, int pos0,
                               const int32_t* __restrict__ pos0_dev, __half* k_pool, __half* v_pool,
                               const int32_t* __restrict__ block_table, int max_pages, int PT) {
  griddep_launch();
  // position, page and the cos/sin row were fixed before the previous kernel started (the counters are advanced at the end
  // of a decode step, the block table is uploaded before the step): only the qkv row depends on the predecessor
  const long m = blockIdx.x;
  const int b = static_cast<int>(m / T), tt = static_cast<int>(m % T);
  const int pos = (pos0_dev ? *pos0_dev : pos0) + tt;
  const int page = block_table[b * max_pages + pos / PT];
  const int slot = pos % PT;
  __half* row = qkv + m * 3L * H;
  const float2* csr = cs + static_cast<long>(pos) * 64;
  float c[8], s[8];
  {
    const int j0 = (threadIdx.x & 7) * 8;  // blockDim is a multiple of 8: a thread always handles the same j0
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float2 v = __ldg(csr + j0 + j); c[j] = v.x; s[j] = v.y; }
  }
  griddep_wait();
  for (int i = threadIdx.x; i < n_heads * 8; i += blockDim.x) {
    const int h = i >> 3, j0 = (i & 7) * 8;
    // q: in place
    {
      uint4* p1 = reinterpret_cast<uint4*>(row + h * 128 + j0);
      uint4* p2 = reinterpret_cast<uint4*>(row + h * 128 + 64 + j0);
      float x1[8], x2[8], o1[8], o2[8];
      unpack8(*p1, x1);
      unpack8(*p2, x2);
#pragma unroll
      for (int j = 0; j < 8; ++j) { o1[j] = x1[j] * c[j] - x2[j] * s[j]; o2[j] = x2[j] * c[j] + x1[j] * s[j]; }
      *p1 = pack8(o1);
      *p2 = pack8(o2);
    }
    const long coff = ((static_cast<long>(page) * n_heads + h) * PT + slot) * 128;
    // k: rotate, write to cache
    {
      const uint4* p1 = reinterpret_cast<const uint4*>(row + H + h * 128 + j0);
      const uint4* p2 = reinterpret_cast<const uint4*>(row + H + h * 128 + 64 + j0);
      float x1[8], x2[8], o1[8], o2[8];
      unpack8(*p1, x1);
      unpack8(*p2, x2);
#pragma unroll
      for (int j = 0; j < 8; ++j) { o1[j] = x1[j] * c[j] - x2[j] * s[j]; o2[j] = x2[j] * c[j] + x1[j] * s[j]; }
      *reinterpret_cast<uint4*>(k_pool + coff + j0) = pack8(o1);
      *reinterpret_cast<uint4*>(k_pool + coff + 64 + j0) = pack8(o2);
    }
    // v: copy
    {
      *reinterpret_cast<uint4*>(v_pool + coff + j0) = *reinterpret_cast<const uint4*>(row + 2 * H + h * 128 + j0);
      *reinterpret_cast<uint4*>(v_pool + coff + 64 + j0) =
          *reinterpret_cast<const uint4*>(row + 2 * H + h * 128 + 64 + j0);
    }
  }
}

// ---- splice: text rows gathered from embed_tokens, visual rows from the projector outputs, pad rows zero
//      (prepare_inputs_labels_for_multimodal: vcoder_ds_llava_arch.py:175-312, vcoder_llava_arch.py:182-294; K7)
__global__ void splice_gather_kernel(const int32_t* __restrict__ plan, const __half* __restrict__ embed,
                                     const __half* __restrict__ feats, __half* __restrict__ out, int H) {
  const long r = blockIdx.x;
  const int v = plan[r];
  uint4* o = reinterpret_cast<uint4*>(out + r * H);
  const int nch = H / 8;
  if (v == -1) {
    for (int c = threadIdx.x; c < nch; c += blockDim.x) o[c] = make_uint4(0, 0, 0, 0);
    return;
  }
  const __half* src = v >= 0 ? embed + static_cast<long>(v) * H : feats + static_cast<long>(-v - 2) * H;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  for (int c = threadIdx.x; c < nch; c += blockDim.x) o[c] = __ldg(s + c);
}

__global__ void embed_gather_kernel(const int32_t* __restrict__ ids, const __half* __restrict__ embed,
                                    __half* __restrict__ out, int H) {
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.x;
  const uint4* s = reinterpret_cast<const uint4*>(embed + static_cast<long>(ids[b]) * H);
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<long>(b) * H);
  for (int c = threadIdx.x; c < H / 8; c += blockDim.x) o[c] = __ldg(s + c);
}

// ---- greedy: next = argmax(logits[:, -1]) over fp16 logits, lowest index wins ties (GenerationMixin greedy; K13)
constexpr int kArgThreads = 256;
// embedding gather of the decode step + the producer half of layer 0's input RMSNorm (see SkinnyParams::pn_*)
__global__ void __launch_bounds__(128)
embed_gather_norm_kernel(const int32_t* __restrict__ ids, const __half* __restrict__ embed, __half* __restrict__ out,
                         const __half* __restrict__ norm_w, __half* __restrict__ xn_out, float* __restrict__ rstd, float eps,
                         int H) {
  __shared__ float scratch[4];
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(embed + static_cast<long>(ids[b]) * H);
  const uint4* wq = reinterpret_cast<const uint4*>(norm_w);
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<long>(b) * H);
  uint4* on = reinterpret_cast<uint4*>(xn_out + static_cast<long>(b) * H);
  float sq = 0.f;
  for (int c = threadIdx.x; c < H / 8; c += blockDim.x) {
    const uint4 v = __ldg(src + c);
    o[c] = v;
    float x[8], w[8], y[8];
    unpack8(v, x);
    unpack8(__ldg(wq + c), w);
#pragma unroll
    for (int j = 0; j < 8; ++j) { sq += x[j] * x[j]; y[j] = x[j] * w[j]; }
    on[c] = pack8(y);
  }
  const float tot = block_sum<128>(sq, scratch);
  if (threadIdx.x == 0) rstd[b] = rsqrtf(tot / static_cast<float>(H) + eps);
}

__global__ void __launch_bounds__(kArgThreads)
argmax_kernel(const __half* __restrict__ logits, int V, int32_t* next, int32_t* trace, int trace_stride,
              const int32_t* __restrict__ step_dev) {
  __shared__ float s_v[kArgThreads / 32];
  __shared__ int s_i[kArgThreads / 32];
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.x;
  const __half* row = logits + static_cast<long>(b) * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += kArgThreads) {
    const float v = __half2float(row[i]);
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kArgThreads / 32; ++w) {
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) { best = s_v[w]; bi = s_i[w]; }
    }
    next[b] = bi;
    if (trace) trace[static_cast<long>(b) * trace_stride + (step_dev ? *step_dev : 0)] = bi;
  }
}

// ---- feature_select 'patch': drop the CLS row (clip_encoder.py:32-33)
__global__ void compact_patches_kernel(const __half* __restrict__ hidden, __half* __restrict__ out, int tokens, int D) {
  const int p = blockIdx.x, b = blockIdx.y;
  const uint4* s = reinterpret_cast<const uint4*>(hidden + (static_cast<long>(b) * tokens + 1 + p) * D);
  uint4* o = reinterpret_cast<uint4*>(out + (static_cast<long>(b) * (tokens - 1) + p) * D);
  for (int c = threadIdx.x; c < D / 8; c += blockDim.x) o[c] = __ldg(s + c);
}

// counters: [0] cache_len (= position of the token being decoded), [1] tokens produced, [2] cache_len + 1
__global__ void advance_kernel(int32_t* counters) {
  griddep_launch();
  griddep_wait();
  if (threadIdx.x == 0) { counters[0] += 1; counters[1] += 1; counters[2] += 1; }
}

}  // namespace

int im2col_patches(const __half* pixels, __half* out, int B, int image, int patch, int k_pad, cudaStream_t s) {
  const int npr = image / patch;
  dim3 grid(npr * npr, B);
  KernelScope scope(kCatMisc, s);
  im2col_kernel<<<grid, 128, 0, s>>>(pixels, out, image, patch, npr, 3 * patch * patch, k_pad);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

int write_cls_rows(const __half* cls, const __half* pos, __half* hidden, int B, int tokens, int D, cudaStream_t s) {
  KernelScope scope(kCatMisc, s);
  cls_rows_kernel<<<B, 128, 0, s>>>(cls, pos, hidden, tokens, D);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

int layernorm_f16(const __half* x, const __half* w, const __half* b, __half* y, long rows, int D, float eps,
                  cudaStream_t s) {
  VCB_CHECK(D % 8 == 0 && D <= kLnMaxChunks * 256, kErrInvalid, "layernorm_f16: D must be a multiple of 8, <= 2048");
  KernelScope scope(kCatNorm, s);
  layernorm_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, s>>>(x, w, b, y, rows, D, eps);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

int rmsnorm_f16(const __half* x, const __half* w, __half* y, long rows, int H, float eps, cudaStream_t s) {
  VCB_CHECK(H % 8 == 0 && H <= kRmsMaxChunks * kRmsThreads * 8, kErrInvalid, "rmsnorm_f16: H must be <= 8192, % 8");
  KernelScope scope(kCatNorm, s);
  unsigned long long* tr = rows <= 64 ? trace_slots(kTraceRmsnorm, static_cast<int>(rows), H) : nullptr;
  VCB_CUDA(launch_pdl(rmsnorm_kernel, dim3(static_cast<unsigned>(rows)), dim3(kRmsThreads), 0, s, x, w, y, H, eps, tr));
  return kOk;
}

int rope_kv_write(__half* qkv, int B, int T, int H, int n_heads, const float2* cos_sin, int pos0,
                  const int32_t* pos0_dev, __half* k_pool, __half* v_pool, const int32_t* block_table, int max_pages,
                  int page_tokens, cudaStream_t s) {
  VCB_CHECK(H == n_heads * 128, kErrInvalid, "rope_kv_write: head_dim must be 128");
  const int threads = std::min(256, n_heads * 8);
  KernelScope scope(kCatRope, s);
  VCB_CUDA(launch_pdl(rope_kv_kernel, dim3(static_cast<unsigned>(static_cast<long>(B) * T)), dim3(threads), 0, s, qkv, T, H,
                      n_heads, cos_sin, pos0, pos0_dev, k_pool, v_pool, block_table, max_pages, page_tokens));
  return kOk;
}

int splice_gather(const int32_t* plan, const __half* embed, const __half* feats, __half* out, long rows, int H,
                  cudaStream_t s) {
  KernelScope scope(kCatMisc, s);
  splice_gather_kernel<<<static_cast<unsigned>(rows), 128, 0, s>>>(plan, embed, feats, out, H);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

int embed_gather(const int32_t* ids, const __half* embed, __half* out, int B, int H, cudaStream_t s) {
  KernelScope scope(kCatMisc, s);
  VCB_CUDA(launch_pdl(embed_gather_kernel, dim3(B), dim3(128), 0, s, ids, embed, out, H));
  return kOk;
}

int embed_gather_norm(const int32_t* ids, const __half* embed, __half* out, const __half* norm_w, __half* xn_out,
                      float* rstd, float eps, int B, int H, cudaStream_t s) {
  KernelScope scope(kCatMisc, s);
  VCB_CUDA(launch_pdl(embed_gather_norm_kernel, dim3(B), dim3(128), 0, s, ids, embed, out, norm_w, xn_out, rstd, eps, H));
  return kOk;
}

int argmax_rows(const __half* logits, int B, int V, int32_t* next, int32_t* trace, int trace_stride,
                const int32_t* step_dev, cudaStream_t s) {
  KernelScope scope(kCatMisc, s);
  VCB_CUDA(launch_pdl(argmax_kernel, dim3(B), dim3(kArgThreads), 0, s, logits, V, next, trace, trace_stride, step_dev));
  return kOk;
}

int compact_patches(const __half* hidden, __half* out, int B, int tokens, int D, cudaStream_t s) {
  dim3 grid(tokens - 1, B);
  KernelScope scope(kCatMisc, s);
  compact_patches_kernel<<<grid, 128, 0, s>>>(hidden, out, tokens, D);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

int advance_counters(int32_t* counters, cudaStream_t s) {
  KernelScope scope(kCatMisc, s);
  VCB_CUDA(launch_pdl(advance_kernel, dim3(1), dim3(32), 0, s, counters));
  return kOk;
}

// cos/sin table as HF computes it: inv_freq = theta^(-2j/d) in fp32, angle = pos * inv_freq in fp32
// (modeling_llama.py:96-113, 124-141)
void build_rope_table(float2* t, int max_pos, int head_dim, float theta) {
  const int half = head_dim / 2;
  for (int j = 0; j < half; ++j) {
    const float inv_freq = 1.0f / std::pow(theta, static_cast<float>(2 * j) / static_cast<float>(head_dim));
    for (int p = 0; p < max_pos; ++p) {
      const float ang = static_cast<float>(p) * inv_freq;
      t[static_cast<long>(p) * half + j] = make_float2(static_cast<float>(std::cos(static_cast<double>(ang))),
                                                       static_cast<float>(std::sin(static_cast<double>(ang))));
    }
  }
}

}  // namespace vcb
