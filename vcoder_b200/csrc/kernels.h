// Internal launcher declarations for the non-GEMM kernels of libvcoder_b200.
#pragma once
#include "common.h"

namespace vcb {

// ---- attention -----------------------------------------------------------------------------------
struct AttnPrefillParams {
  // dense path (ViT): q/k/v are column blocks of one [B*S, 3*D] buffer
  const __half* q = nullptr;
  const __half* k = nullptr;
  const __half* v = nullptr;
  long q_stride = 0;   // elements between consecutive tokens of q
  long kv_stride = 0;  // same for k/v (dense path)
  // paged path (LLaMA): K/V pools laid out [page][head][page_tokens][head_dim]
  const __half* k_pool = nullptr;
  const __half* v_pool = nullptr;
  const int32_t* block_table = nullptr;  // [batch][max_pages]
  int max_pages = 0;
  int page_tokens = 0;
  int pool_pages = 0;     // pages in the pools (sizes the TMA view of the tcgen05 kernel); 0 = unknown
  __half* out = nullptr;  // [B*S, heads*head_dim]
  long out_stride = 0;
  int batch = 0, n_heads = 0, seq_len = 0, head_dim = 0;
  // paged causal path only: query row i of a sequence sits at position q_pos0 + i and attends to cached keys
  // [0, q_pos0 + i]; seq_len counts the query rows (f3: prefill of a suffix behind a reused KV prefix)
  int q_pos0 = 0;
  float scale = 1.f;
};
int attn_vit(const AttnPrefillParams& p, cudaStream_t stream);       // dispatcher: tcgen05 kernel when it applies
int attn_vit_tc5(const AttnPrefillParams& p, cudaStream_t stream);   // tcgen05 + TMEM (attn_vit_tc5.cu)
int attn_vit_legacy(const AttnPrefillParams& p, cudaStream_t stream);  // mma.sync flash attention
int attn_vit_pair(const AttnPrefillParams& p, cudaStream_t stream);    // tcgen05, two query tiles per CTA (attn_vit_pair.cu)
int attn_llama_prefill(const AttnPrefillParams& p, cudaStream_t stream);       // dispatcher (VCB_PREFILL_ATTN=tc5|legacy)
int attn_llama_prefill_tc5(const AttnPrefillParams& p, cudaStream_t stream);   // tcgen05 + TMEM + TMA (attn_prefill_tc5.cu)

struct AttnDecodeParams {
  const __half* q = nullptr;  // [B, 3H] (rotated q in the first H columns), stride q_stride
  long q_stride = 0;
  const __half* k_pool = nullptr;
  const __half* v_pool = nullptr;
  const int32_t* block_table = nullptr;
  int max_pages = 0, page_tokens = 0;
  const int32_t* ctx_len = nullptr;  // device scalar: number of cached tokens INCLUDING the current one
  int max_ctx = 0;                   // host upper bound used to size the grid
  // fused RoPE + KV append (decode): q is un-rotated, k/v of the current token are read from the QKV row (k at +H,
  // v at +2H), rotated with cos_sin[pos] and written to the cache by the attention kernel itself
  const float2* cos_sin = nullptr;
  __half* k_pool_w = nullptr;
  __half* v_pool_w = nullptr;
  float* part_o = nullptr;           // [B*heads, splits, 128]
  float* part_ml = nullptr;          // [B*heads, splits, 2]
  __half* out = nullptr;             // [B, H]
  int32_t* sync_counters = nullptr;  // [B*heads] zero-initialised arrival counters (last-CTA merge of the bulk-copy kernel)
  int batch = 0, n_heads = 0;
  int prefetch = 0;  // L2-prefetch the split's KV pages before the grid-dependency wait (VCB_ATTN_DEC_PF)
  int chunk = 256;   // keys per flash-decoding split (register-staged kernel)
  float scale = 1.f;
  unsigned long long* trace = nullptr;   // timeline slots of the split kernel (common.h trace_slots), else null
  unsigned long long* trace2 = nullptr;  // ... of the combine kernel
};
int attn_llama_decode(const AttnDecodeParams& p, cudaStream_t stream);         // dispatcher (VCB_ATTN_DEC=stream|split|...)
int attn_llama_decode_stream(const AttnDecodeParams& p, cudaStream_t stream);  // balanced page-streaming kernel (attn_decode_stream.cu)
int attn_decode_stream_slots(int batch, int n_heads);  // partial-state slots per (sequence, head) of the streaming kernel
int attn_decode_splits(int max_ctx);

// ---- skinny GEMM (decode / lm_head): Y[B,N] = X[B,K] * W[N,K]^T ---------------------------------------
struct SkinnyParams {
  const __half* X = nullptr;  // [B, K], ldx
  long ldx = 0;
  const __half* W = nullptr;  // [N, K] row-major
  const __half* Wp = nullptr; // optional tile-major copy of W made by pack_skinny_weights: the weight stream then reads
                              // this instead of W (linear 8 KiB bulk copies instead of 16 x 512-byte tensor-map boxes)
  __half* Y = nullptr;        // [B, N_out], ldy
  long ldy = 0;
  const __half* residual = nullptr;  // [B, N], ldr (may alias Y)
  long ldr = 0;
  int B = 0, N = 0, K = 0;
  int swiglu = 0;  // W rows 8/8 interleaved gate|up; N_out = N/2
  int tpc = 1;     // multi-tile kernel (skinny_stream.cu): consecutive 16-row tiles per CTA (set by the launcher)
  int mode = 1;    // chunk->warp mapping (see skinny.cu)
  int prefetch = 0;  // cp.async.bulk.prefetch.L2 ahead of the demand loads (measured slower: off)
  int pf_groups = 0;  // TMA kernel: k-groups per warp prefetched into L2 before the grid-dependency wait (first wave only)
  int pf_ctas = 0;    // CTAs that count as the first wave
  int double_buffer = 0;  // keep the next batch of weight loads in flight during the MMAs (costs registers)
  // fused LlamaRMSNorm on X (decode): Y = rsqrt(mean(x^2)+eps) * W (x o norm_w); null = X is used as is
  const __half* norm_w = nullptr;
  float norm_eps = 0.f;
  // RMSNorm chained through the GEMMs of a decode step (no norm kernel in between), TMA kernel only:
  //  producer side (o-proj / down-proj, Y = new residual stream x): also emits the NEXT norm's weighted input
  //    pn_out[b, c] = fp16(x[b, c] * pn_w[c]) and this CTA's share of sum(x^2) (pn_parts[blockIdx.x * 32 + b]); the CTA
  //    that finishes last folds the shares into pn_rstd[b] = rsqrt(mean(x^2) + pn_eps) (pn_ticket: zeroed arrival counter)
  //  consumer side (qkv / gate|up / lm_head): X is such a weighted input; the fp32 accumulators are scaled by cn_rstd[b]
  //    in the epilogue
  const __half* pn_w = nullptr;
  __half* pn_out = nullptr;
  long pn_ld = 0;
  float* pn_parts = nullptr;
  float* pn_rstd = nullptr;
  int32_t* pn_ticket = nullptr;
  float pn_eps = 0.f;
  const float* cn_rstd = nullptr;
  unsigned long long* trace = nullptr;  // timeline slots (common.h trace_slots), else null
};
int skinny_gemm(const SkinnyParams& p, cudaStream_t stream);
int tune_skinny(int N, int K, int warps, int depth);  // ring shape of the TMA kernel for one (N, K); 0 = default
int skinny_gemm_stream(const SkinnyParams& p, cudaStream_t stream);  // balanced multi-tile TMA kernel (skinny_stream.cu): the default
int skinny_gemm_tma(const SkinnyParams& p, cudaStream_t stream);  // TMA / bulk-copy fed kernel (skinny_tma.cu); needs K % 64 == 0
// W[N,K] row-major -> the streaming layout of the decode GEMM: [N/16][K/64][16 rows][128 B, 16-byte chunks pre-swizzled]
int pack_skinny_weights(const __half* W, __half* Wp, int N, int K, cudaStream_t stream);

// ---- elementwise / normalisation / data movement ----------------------------------------------------------
int im2col_patches(const __half* pixels, __half* out, int B, int image, int patch, int k_pad, cudaStream_t s);
int write_cls_rows(const __half* cls, const __half* pos, __half* hidden, int B, int tokens, int D, cudaStream_t s);
int layernorm_f16(const __half* x, const __half* w, const __half* b, __half* y, long rows, int D, float eps,
                  cudaStream_t s);
int rmsnorm_f16(const __half* x, const __half* w, __half* y, long rows, int H, float eps, cudaStream_t s);
// rows of qkv [M,3H]: rotate q in place, write rotated k and v into the paged cache.
// position of row m = pos0 (host) or *pos0_dev (device scalar, decode) + (m % T)
int rope_kv_write(__half* qkv, int B, int T, int H, int n_heads, const float2* cos_sin, int pos0,
                  const int32_t* pos0_dev, __half* k_pool, __half* v_pool, const int32_t* block_table, int max_pages,
                  int page_tokens, cudaStream_t s);
// out[r] = plan[r] >= 0 ? embed[plan[r]] : plan[r] == -1 ? 0 : feats[-plan[r]-2]
int splice_gather(const int32_t* plan, const __half* embed, const __half* feats, __half* out, long rows, int H,
                  cudaStream_t s);
int embed_gather(const int32_t* ids, const __half* embed, __half* out, int B, int H, cudaStream_t s);
// embed_gather + the first RMSNorm's producer half: out = embed[ids], xn_out = fp16(out * norm_w), rstd[b] = rsqrt(mean(out[b]^2) + eps)
int embed_gather_norm(const int32_t* ids, const __half* embed, __half* out, const __half* norm_w, __half* xn_out,
                      float* rstd, float eps, int B, int H, cudaStream_t s);
// first-index argmax over fp16 logits [B,V]; writes next[b] and (optionally) trace[b*trace_stride + *step]
int argmax_rows(const __half* logits, int B, int V, int32_t* next, int32_t* trace, int trace_stride,
                const int32_t* step_dev, cudaStream_t s);
// do_sample=True: temperature -> top-k (ties kept) -> top-p -> inverse-CDF draw at u = Philox(seed; step, row) on the
// bitonic-sorted row (sample.cu).  params_dev: 20-byte block written by pack_sample_params; step = *step_dev or step_host.
int sample_rows(const __half* logits, int B, int V, const void* params_dev, const int32_t* step_dev, int step_host,
                int32_t* next, int32_t* trace, int trace_stride, int32_t* kept_out, cudaStream_t s);
void pack_sample_params(float temperature, int top_k, float top_p, uint64_t seed, void* out20);
uint32_t philox_first_host(uint64_t seed, uint32_t step, uint32_t row);
// process_images (mm_utils.py:28-40) for one uint8 HWC RGB image (host or device) -> fp16 [3,size,size] on the device:
// optional expand2square with pad_rgb, PIL-exact bicubic resize of the shortest edge to `size`, centre crop, rescale,
// normalize (preprocess.cu).  Synchronous on `st`.
int preprocess_u8(const uint8_t* img_any, int height, int width, const uint8_t* pad_rgb, int size, const float* mean,
                  const float* stdv, __half* out_dev, cudaStream_t st);
int compact_patches(const __half* hidden, __half* out, int B, int tokens, int D, cudaStream_t s);
int advance_counters(int32_t* counters, cudaStream_t s);
void build_rope_table(float2* host_table, int max_pos, int head_dim, float theta);

}  // namespace vcb
