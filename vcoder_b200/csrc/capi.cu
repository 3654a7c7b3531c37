// extern "C" surface of libvcoder_b200 (declarations + reference citations: include/vcoder_b200.h).
#include "vcoder_b200.h"

#include "common.h"

namespace vcb {
static thread_local std::string g_err;
void set_thread_error(const std::string& m) { g_err = m; }
const char* get_thread_error() { return g_err.c_str(); }
}  // namespace vcb

using namespace vcb;

static int require_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_thread_error("no CUDA device visible: libvcoder_b200 has no CPU fallback");
    return kErrNoDevice;
  }
  return kOk;
}

extern "C" {

#ifdef VCB_EXPERIMENTAL
const char* vcb_version(void) { return "vcoder_b200 0.2.0 (sm_100a, experimental variants)"; }
#else
const char* vcb_version(void) { return "vcoder_b200 0.2.0 (sm_100a)"; }
#endif
const char* vcb_last_error(void) { return get_thread_error(); }

int vcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vcb_op_gemm_f16(const vcb_half* A, int32_t lda, const vcb_half* W, int32_t ldw, int32_t M, int32_t N, int32_t K,
                    const vcb_half* bias, const vcb_half* residual, int32_t ldr, int32_t act, vcb_half* C,
                    int32_t ldc, int32_t block_n, void* stream) {
  VCB_TRY(require_device());
  GemmEpilogue ep;
  ep.C = reinterpret_cast<__half*>(C);
  ep.ldc = ldc;
  ep.bias = reinterpret_cast<const __half*>(bias);
  ep.residual = reinterpret_cast<const __half*>(residual);
  ep.ldr = ldr;
  ep.act = act;
  return gemm_f16(reinterpret_cast<const __half*>(A), lda, reinterpret_cast<const __half*>(W), ldw, M, N, K, ep,
                  block_n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
#include <algorithm>
#include <vector>

#include "engine.h"
#include "kernels.h"

struct vcb_engine {
  vcb::Engine* e;
};

namespace {
inline const __half* H(const vcb_half* p) { return reinterpret_cast<const __half*>(p); }
inline __half* H(vcb_half* p) { return reinterpret_cast<__half*>(p); }
inline cudaStream_t ST(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

int vcb_op_skinny_gemm_f16(const vcb_half* X, int32_t ldx, const vcb_half* W, int32_t B, int32_t N, int32_t K,
                           const vcb_half* residual, int32_t ldr, int32_t swiglu, vcb_half* Y, int32_t ldy,
                           void* stream) {
  VCB_TRY(require_device());
  SkinnyParams s;
  s.X = H(X); s.ldx = ldx; s.W = H(W); s.Y = H(Y); s.ldy = ldy; s.residual = H(residual); s.ldr = ldr;
  s.B = B; s.N = N; s.K = K; s.swiglu = swiglu;
  return skinny_gemm(s, ST(stream));
}

int vcb_op_skinny_gemm_onetile_f16(const vcb_half* X, int32_t ldx, const vcb_half* W, int32_t B, int32_t N, int32_t K,
                                   const vcb_half* residual, int32_t ldr, int32_t swiglu, vcb_half* Y, int32_t ldy,
                                   void* stream) {
  VCB_TRY(require_device());
  SkinnyParams s;
  s.X = H(X); s.ldx = ldx; s.W = H(W); s.Y = H(Y); s.ldy = ldy; s.residual = H(residual); s.ldr = ldr;
  s.B = B; s.N = N; s.K = K; s.swiglu = swiglu;
  return skinny_gemm_tma(s, ST(stream));
}

int vcb_op_pack_skinny_weights(const vcb_half* W, int32_t N, int32_t K, vcb_half* Wp, void* stream) {
  VCB_TRY(require_device());
  return pack_skinny_weights(H(W), H(Wp), N, K, ST(stream));
}

int vcb_op_skinny_gemm_packed_f16(const vcb_half* X, int32_t ldx, const vcb_half* Wp, int32_t B, int32_t N, int32_t K,
                                  const vcb_half* residual, int32_t ldr, int32_t swiglu, vcb_half* Y, int32_t ldy,
                                  void* stream) {
  VCB_TRY(require_device());
  VCB_CHECK(Wp != nullptr, kErrInvalid, "vcb_op_skinny_gemm_packed_f16: Wp is null");
  SkinnyParams s;
  s.X = H(X); s.ldx = ldx; s.W = H(Wp); s.Wp = H(Wp); s.Y = H(Y); s.ldy = ldy; s.residual = H(residual); s.ldr = ldr;
  s.B = B; s.N = N; s.K = K; s.swiglu = swiglu;
  return skinny_gemm_tma(s, ST(stream));
}

int vcb_op_skinny_rmsnorm_gemm_f16(const vcb_half* X, int32_t ldx, const vcb_half* norm_w, float eps, const vcb_half* W,
                                   int32_t B, int32_t N, int32_t K, const vcb_half* residual, int32_t ldr, int32_t swiglu,
                                   vcb_half* Y, int32_t ldy, void* stream) {
  VCB_TRY(require_device());
  VCB_CHECK(norm_w != nullptr, kErrInvalid, "vcb_op_skinny_rmsnorm_gemm_f16: norm_w is null");
  SkinnyParams s;
  s.X = H(X); s.ldx = ldx; s.W = H(W); s.Y = H(Y); s.ldy = ldy; s.residual = H(residual); s.ldr = ldr;
  s.B = B; s.N = N; s.K = K; s.swiglu = swiglu; s.norm_w = H(norm_w); s.norm_eps = eps;
  return skinny_gemm(s, ST(stream));
}

int vcb_op_layernorm_f16(const vcb_half* x, const vcb_half* w, const vcb_half* b, vcb_half* y, int64_t rows, int32_t D,
                         float eps, void* stream) {
  VCB_TRY(require_device());
  return layernorm_f16(H(x), H(w), H(b), H(y), rows, D, eps, ST(stream));
}

int vcb_op_rmsnorm_f16(const vcb_half* x, const vcb_half* w, vcb_half* y, int64_t rows, int32_t Hd, float eps,
                       void* stream) {
  VCB_TRY(require_device());
  return rmsnorm_f16(H(x), H(w), H(y), rows, Hd, eps, ST(stream));
}

int vcb_op_attention_vit_f16(const vcb_half* qkv, int32_t B, int32_t S, int32_t heads, float scale, vcb_half* out,
                             void* stream) {
  VCB_TRY(require_device());
  const int D = heads * 64;
  AttnPrefillParams a;
  a.q = H(qkv); a.k = a.q + D; a.v = a.q + 2 * D;
  a.q_stride = 3 * D; a.kv_stride = 3 * D;
  a.out = H(out); a.out_stride = D;
  a.batch = B; a.n_heads = heads; a.seq_len = S; a.head_dim = 64; a.scale = scale;
  return attn_vit(a, ST(stream));
}

int vcb_op_attention_prefill_f16(const vcb_half* q, int64_t q_stride, const vcb_half* k_pool, const vcb_half* v_pool,
                                 const int32_t* block_table, int32_t max_pages, int32_t B, int32_t heads, int32_t Tq,
                                 int32_t q_pos0, float scale, vcb_half* out, int64_t out_stride, void* stream) {
  VCB_TRY(require_device());
  VCB_CHECK(q != nullptr && k_pool != nullptr && v_pool != nullptr && block_table != nullptr && out != nullptr, kErrInvalid,
            "vcb_op_attention_prefill_f16: null argument");
  VCB_CHECK(B >= 1 && heads >= 1 && Tq >= 1 && q_pos0 >= 0 && (q_pos0 + Tq + 63) / 64 <= max_pages, kErrInvalid,
            "vcb_op_attention_prefill_f16: bad shape");
  AttnPrefillParams a;
  a.q = H(q); a.q_stride = q_stride;
  a.k_pool = H(k_pool); a.v_pool = H(v_pool); a.block_table = block_table; a.max_pages = max_pages; a.page_tokens = 64;
  a.out = H(out); a.out_stride = out_stride;
  a.batch = B; a.n_heads = heads; a.seq_len = Tq; a.q_pos0 = q_pos0; a.head_dim = 128; a.scale = scale;
  return attn_llama_prefill(a, ST(stream));
}

int vcb_op_attention_decode_f16(const vcb_half* qkv, vcb_half* k_pool, vcb_half* v_pool, const int32_t* block_table,
                                int32_t max_pages, int32_t B, int32_t heads, int32_t ctx_len, int32_t max_ctx,
                                int32_t fuse_rope, float rope_theta, float scale, vcb_half* out, void* stream) {
  VCB_TRY(require_device());
  VCB_CHECK(qkv != nullptr && k_pool != nullptr && v_pool != nullptr && block_table != nullptr && out != nullptr, kErrInvalid,
            "vcb_op_attention_decode_f16: null argument");
  VCB_CHECK(B >= 1 && heads >= 1 && ctx_len >= 1 && max_ctx >= ctx_len && (max_ctx + 63) / 64 <= max_pages, kErrInvalid,
            "vcb_op_attention_decode_f16: bad shape");
  cudaStream_t st = ST(stream);
  const int splits = std::max(attn_decode_splits(max_ctx), attn_decode_stream_slots(B, heads));
  const size_t n_po = static_cast<size_t>(B) * heads * splits * 128 * 4, n_pml = static_cast<size_t>(B) * heads * splits * 2 * 4;
  const size_t n_cs = fuse_rope ? static_cast<size_t>(ctx_len) * 64 * sizeof(float2) : 0;
  const size_t n_sync = (static_cast<size_t>(B) * heads * 4 + 255) & ~static_cast<size_t>(255);
  char* scratch = nullptr;
  VCB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), n_po + n_pml + n_cs + 256 + n_sync, st));
  cudaMemsetAsync(scratch + n_po + n_pml + 256 + n_cs, 0, n_sync, st);
  AttnDecodeParams a;
  a.part_o = reinterpret_cast<float*>(scratch);
  a.part_ml = reinterpret_cast<float*>(scratch + n_po);
  int32_t* ctx_dev = reinterpret_cast<int32_t*>(scratch + n_po + n_pml);
  std::vector<float2> tab;
  int rc = kOk;
  cudaError_t ce = cudaMemcpyAsync(ctx_dev, &ctx_len, 4, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess && fuse_rope) {
    tab.resize(static_cast<size_t>(ctx_len) * 64);
    build_rope_table(tab.data(), ctx_len, 128, rope_theta);
    float2* cs_dev = reinterpret_cast<float2*>(scratch + n_po + n_pml + 256);
    ce = cudaMemcpyAsync(cs_dev, tab.data(), n_cs, cudaMemcpyHostToDevice, st);
    a.cos_sin = cs_dev; a.k_pool_w = H(k_pool); a.v_pool_w = H(v_pool);
  }
  if (ce == cudaSuccess) {
    a.q = H(qkv); a.q_stride = 3L * heads * 128; a.k_pool = H(k_pool); a.v_pool = H(v_pool);
    a.block_table = block_table; a.max_pages = max_pages; a.page_tokens = 64;
    a.sync_counters = reinterpret_cast<int32_t*>(scratch + n_po + n_pml + 256 + n_cs);
    a.ctx_len = ctx_dev; a.max_ctx = max_ctx; a.out = H(out); a.batch = B; a.n_heads = heads; a.scale = scale;
    rc = attn_llama_decode(a, st);
  }
  cudaFreeAsync(scratch, st);
  const cudaError_t se = cudaStreamSynchronize(st);  // ctx_len / tab are host stack / heap buffers
  VCB_CUDA(ce);
  VCB_TRY(rc);
  VCB_CUDA(se);
  return kOk;
}

int vcb_splice_plan(int32_t variant, const int64_t* ids, int32_t B, int32_t T, int32_t n_patches, int32_t has_seg,
                    const uint8_t* depth_nonzero, int32_t* plan_out, int64_t plan_cap, int32_t* out_len) {
  VCB_CHECK(ids != nullptr && out_len != nullptr && B >= 1 && T >= 1, kErrInvalid, "vcb_splice_plan: bad arguments");
  VCB_CHECK(variant >= VCB_VARIANT_VCODER && variant <= VCB_VARIANT_VCODER_IT, kErrInvalid, "bad variant");
  std::vector<int32_t> plan;
  int len = 0;
  VCB_TRY(splice_plan_host(variant, ids, B, T, n_patches, has_seg, depth_nonzero, &plan, &len, nullptr));
  *out_len = len;
  if (plan_out) {
    VCB_CHECK(static_cast<int64_t>(plan.size()) <= plan_cap, kErrNoMem, "vcb_splice_plan: plan_out too small");
    std::copy(plan.begin(), plan.end(), plan_out);
  }
  return kOk;
}

int vcb_create(const vcb_config* cfg, int32_t device, vcb_handle* out) {
  VCB_TRY(require_device());
  VCB_CHECK(out != nullptr, kErrInvalid, "vcb_create: null out");
  Engine* e = nullptr;
  VCB_TRY(engine_create(cfg, device, &e));
  *out = new vcb_engine{e};
  return kOk;
}

int vcb_destroy(vcb_handle h) {
  if (h) {
    delete h->e;
    delete h;
  }
  return kOk;
}

int vcb_bind_weights(vcb_handle h, const vcb_tensor* tensors, int32_t n) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  return engine_bind(h->e, tensors, n);
}

int vcb_tower_forward(vcb_handle h, const vcb_half* pixels, int32_t n, vcb_half* out, void* stream) {
  VCB_CHECK(h != nullptr && pixels != nullptr && out != nullptr, kErrInvalid, "vcb_tower_forward: null argument");
  Engine* e = h->e;
  VCB_CUDA(cudaSetDevice(e->device));
  VCB_TRY(engine_check_bound(e, true, false));
  VCB_CHECK(n >= 1, kErrInvalid, "vcb_tower_forward: n must be >= 1");
  const size_t img = 3ul * e->cfg.vit_image_size * e->cfg.vit_image_size * 2;
  const size_t orow = static_cast<size_t>(e->P) * e->cfg.vit_hidden * 2;
  StreamGuard sg(e, ST(stream));
  e->feat_key = 0;  // pix / vh are overwritten: the prefix cache must not skip the towers on the next prefill
  // any number of images: the handle's vision workspaces hold 3*max_batch of them, larger batches run in chunks of that
  // size (BASELINE configs[4] sweeps the tower to 256 images per call)
  for (int i0 = 0; i0 < n; i0 += e->vit_max_images) {
    const int m = std::min(e->vit_max_images, n - i0);
    VCB_TRY(copy_any(e->pix.p, reinterpret_cast<const char*>(pixels) + i0 * img, m * img, sg.st));
    VCB_TRY(engine_tower(e, e->pix.as<__half>(), m, sg.st));
    // feature_select 'patch': drop CLS; compact into vn then copy out (out may be host memory)
    VCB_TRY(compact_patches(e->vh.as<__half>(), e->vn.as<__half>(), m, e->S, e->cfg.vit_hidden, sg.st));
    VCB_TRY(copy_any(reinterpret_cast<char*>(out) + i0 * orow, e->vn.p, m * orow, sg.st));
  }
  return kOk;
}

int vcb_encode(vcb_handle h, int32_t modality, const vcb_half* pixels, int32_t n, vcb_half* out, void* stream) {
  VCB_CHECK(h != nullptr && pixels != nullptr && out != nullptr, kErrInvalid, "vcb_encode: null argument");
  VCB_CHECK(modality >= 0 && modality <= 2, kErrInvalid, "vcb_encode: modality must be 0 (rgb), 1 (seg) or 2 (depth)");
  Engine* e = h->e;
  VCB_CHECK(modality == 0 || e->cfg.variant != VCB_VARIANT_LLAVA, kErrInvalid, "vcb_encode: LLaVA has no seg projector");
  VCB_CHECK(modality != 2 || e->cfg.variant == VCB_VARIANT_VCODER_DS, kErrInvalid,
            "vcb_encode: encode_depth_images exists on VCoder-DS only");
  VCB_CUDA(cudaSetDevice(e->device));
  VCB_TRY(engine_check_bound(e, true, true));
  VCB_CHECK(n >= 1, kErrInvalid, "vcb_encode: n must be >= 1");
  const size_t img = 3ul * e->cfg.vit_image_size * e->cfg.vit_image_size * 2;
  const size_t orow = static_cast<size_t>(e->P) * e->cfg.llm_hidden * 2;
  StreamGuard sg(e, ST(stream));
  e->feat_key = 0;  // pix / vh / feats are overwritten
  for (int i0 = 0; i0 < n; i0 += e->vit_max_images) {  // chunks of the workspace capacity, as vcb_tower_forward
    const int m = std::min(e->vit_max_images, n - i0);
    VCB_TRY(copy_any(e->pix.p, reinterpret_cast<const char*>(pixels) + i0 * img, m * img, sg.st));
    VCB_TRY(engine_tower(e, e->pix.as<__half>(), m, sg.st));
    VCB_TRY(engine_project(e, 0, m * e->S, modality == 0 ? e->mm : e->seg, sg.st));
    VCB_TRY(compact_patches(e->feats.as<__half>(), e->pmid.as<__half>(), m, e->S, e->cfg.llm_hidden, sg.st));
    VCB_TRY(copy_any(reinterpret_cast<char*>(out) + i0 * orow, e->pmid.p, m * orow, sg.st));
  }
  return kOk;
}

int vcb_prefill(vcb_handle h, const int64_t* ids, int32_t B, int32_t T, const vcb_half* images, const vcb_half* segs,
                const vcb_half* depths, int32_t* next_ids, vcb_half* last_logits, vcb_half* embeds, int32_t* spliced_len,
                void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  int len = 0;
  StreamGuard sg(h->e, ST(stream));
  VCB_TRY(engine_prefill(h->e, ids, B, T, images, segs, depths, false, next_ids, last_logits, embeds, &len, nullptr,
                         nullptr, sg.st));
  if (spliced_len) *spliced_len = len;
  return kOk;
}

int vcb_forward_logits(vcb_handle h, const int64_t* ids, int32_t B, int32_t T, const vcb_half* images,
                       const vcb_half* segs, const vcb_half* depths, vcb_half* logits_dev, int64_t logits_halves,
                       int32_t* spliced_len, void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  VCB_CHECK(logits_dev != nullptr && logits_halves > 0, kErrInvalid, "vcb_forward_logits: no logits buffer");
  int len = 0;
  StreamGuard sg(h->e, ST(stream));
  PrefillExtra ex;
  ex.all_logits = logits_dev;
  ex.all_logits_halves = static_cast<long>(logits_halves);
  VCB_TRY(engine_prefill(h->e, ids, B, T, images, segs, depths, false, nullptr, nullptr, nullptr, &len, nullptr, nullptr,
                         sg.st, &ex));
  if (spliced_len) *spliced_len = len;
  return kOk;
}

int vcb_splice_embeds(vcb_handle h, const int64_t* ids, int32_t B, int32_t T, const vcb_half* images,
                      const vcb_half* segs, const vcb_half* depths, vcb_half* embeds, int32_t* spliced_len, void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  VCB_CHECK(embeds != nullptr, kErrInvalid, "vcb_splice_embeds: no embeds buffer");
  int len = 0;
  StreamGuard sg(h->e, ST(stream));
  PrefillExtra ex;
  ex.embeds_only = true;
  VCB_TRY(engine_prefill(h->e, ids, B, T, images, segs, depths, false, nullptr, nullptr, embeds, &len, nullptr, nullptr,
                         sg.st, &ex));
  if (spliced_len) *spliced_len = len;
  return kOk;
}

int vcb_prefill_cached(vcb_handle h, const int64_t* ids, int32_t B, int32_t T, const vcb_half* images,
                       const vcb_half* segs, const vcb_half* depths, int32_t* next_ids, vcb_half* last_logits,
                       int32_t* spliced_len, int32_t* reused_len, int32_t* towers_skipped, void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  int len = 0, reused = 0, skipped = 0;
  StreamGuard sg(h->e, ST(stream));
  VCB_TRY(engine_prefill(h->e, ids, B, T, images, segs, depths, true, next_ids, last_logits, nullptr, &len, &reused,
                         &skipped, sg.st));
  if (spliced_len) *spliced_len = len;
  if (reused_len) *reused_len = reused;
  if (towers_skipped) *towers_skipped = skipped;
  return kOk;
}

int vcb_prefix_match(const int32_t* cached, int32_t cached_len, const int32_t* plan, int32_t plan_len, int32_t B,
                     int32_t features_same, int32_t limit, int32_t* out_len) {
  VCB_CHECK(cached != nullptr && plan != nullptr && out_len != nullptr && B >= 1 && cached_len >= 0 && plan_len >= 0,
            kErrInvalid, "vcb_prefix_match: bad arguments");
  *out_len = prefix_match_host(cached, cached_len, plan, plan_len, B, features_same, limit);
  return kOk;
}

int vcb_decode(vcb_handle h, int32_t n_steps, int32_t* out_ids, vcb_half* logits_trace, void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  StreamGuard sg(h->e, ST(stream));
  return engine_decode(h->e, n_steps, out_ids, logits_trace, sg.st);
}

int vcb_generate(vcb_handle h, const int64_t* ids, int32_t B, int32_t T, const vcb_half* images, const vcb_half* segs,
                 const vcb_half* depths, int32_t max_new_tokens, int64_t* out_ids, void* stream) {
  VCB_CHECK(h != nullptr && out_ids != nullptr, kErrInvalid, "vcb_generate: null argument");
  VCB_CHECK(max_new_tokens >= 1, kErrInvalid, "vcb_generate: max_new_tokens must be >= 1");
  Engine* e = h->e;
  StreamGuard sg(e, ST(stream));
  VCB_TRY(engine_prefill(e, ids, B, T, images, segs, depths, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                         sg.st));
  VCB_TRY(engine_decode(e, max_new_tokens - 1, nullptr, nullptr, sg.st));
  std::vector<int32_t> tmp(static_cast<size_t>(B) * max_new_tokens);
  VCB_CUDA(cudaMemcpy2DAsync(tmp.data(), static_cast<size_t>(max_new_tokens) * 4, e->trace.p,
                             static_cast<size_t>(e->cfg.max_seq_len) * 4, static_cast<size_t>(max_new_tokens) * 4, B,
                             cudaMemcpyDeviceToHost, sg.st));
  VCB_CUDA(cudaStreamSynchronize(sg.st));
  for (size_t i = 0; i < tmp.size(); ++i) out_ids[i] = tmp[i];
  return kOk;
}

int vcb_set_sampling(vcb_handle h, int32_t do_sample, float temperature, int32_t top_k, float top_p, uint64_t seed,
                     void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  StreamGuard sg(h->e, ST(stream));
  return engine_set_sampling(h->e, do_sample, temperature, top_k, top_p, seed, sg.st);
}

int vcb_decode_stream(vcb_handle h, int32_t max_steps, int32_t lookahead, vcb_token_callback cb, void* user,
                      int32_t* n_accepted, void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  StreamGuard sg(h->e, ST(stream));
  return engine_decode_stream(h->e, max_steps, lookahead, cb, user, n_accepted, sg.st);
}

int vcb_op_sample_f16(const vcb_half* logits, int32_t B, int32_t V, float temperature, int32_t top_k, float top_p,
                      uint64_t seed, int32_t step, int32_t* next_ids, int32_t* kept, void* stream) {
  VCB_TRY(require_device());
  VCB_CHECK(logits != nullptr && next_ids != nullptr && B >= 1, kErrInvalid, "vcb_op_sample_f16: bad arguments");
  VCB_CHECK(temperature > 0.f && top_k >= 0 && top_p >= 0.f && top_p <= 1.f, kErrInvalid,
            "vcb_op_sample_f16: temperature > 0, top_k >= 0, 0 <= top_p <= 1");
  unsigned char blk[32] = {0};
  pack_sample_params(temperature, top_k, top_p, seed, blk);
  void* prm = nullptr;
  VCB_CUDA(cudaMallocAsync(&prm, 32, ST(stream)));
  VCB_CUDA(cudaMemcpyAsync(prm, blk, 32, cudaMemcpyHostToDevice, ST(stream)));
  const int rc = sample_rows(H(logits), B, V, prm, nullptr, step, next_ids, nullptr, 0, kept, ST(stream));
  cudaFreeAsync(prm, ST(stream));
  VCB_CUDA(cudaStreamSynchronize(ST(stream)));  // blk is a stack buffer
  return rc;
}

int vcb_op_preprocess_u8(const uint8_t* img, int32_t height, int32_t width, const uint8_t* pad_rgb, int32_t out_size,
                         const float* mean3, const float* std3, vcb_half* out, void* stream) {
  VCB_TRY(require_device());
  return preprocess_u8(img, height, width, pad_rgb, out_size, mean3, std3, H(out), ST(stream));
}

float vcb_philox_uniform(uint64_t seed, uint32_t step, uint32_t row) {
  return static_cast<float>(philox_first_host(seed, step, row) >> 8) * (1.0f / 16777216.0f);
}

int vcb_bench_weight_stream(vcb_handle h, int32_t B, int32_t rounds, float* ms_total, int32_t* launches, void* stream) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  StreamGuard sg(h->e, ST(stream));
  return engine_bench_weight_stream(h->e, B, rounds, ms_total, launches, sg.st);
}

int vcb_reset(vcb_handle h) {
  VCB_CHECK(h != nullptr, kErrInvalid, "null handle");
  return engine_reset(h->e);
}

}  // extern "C"

extern "C" {

uint64_t vcb_launch_count(void) { return vcb::launch_count(); }

int vcb_profile_enable(int32_t on) {
  vcb::prof_enable(on != 0);
  return vcb::kOk;
}

int vcb_tune_skinny(int32_t N, int32_t K, int32_t warps, int32_t depth) { return vcb::tune_skinny(N, K, warps, depth); }

int vcb_trace_begin(int32_t max_ctas) { return vcb::trace_begin(max_ctas); }

int vcb_trace_read(int32_t* launches_out, int32_t max_launches, int32_t* n_launches, uint64_t* stamps_out, int64_t cap) {
  VCB_CHECK(launches_out != nullptr && n_launches != nullptr && stamps_out != nullptr, vcb::kErrInvalid, "vcb_trace_read: null argument");
  return vcb::trace_read(launches_out, max_launches, n_launches, reinterpret_cast<unsigned long long*>(stamps_out), cap);
}

int vcb_profile_read(float* ms_by_category, uint64_t* launches_by_category, int32_t n_categories) {
  VCB_CHECK(ms_by_category != nullptr && launches_by_category != nullptr && n_categories == vcb::kNumCat, vcb::kErrInvalid,
            "vcb_profile_read: need arrays of VCB_PROFILE_CATEGORIES entries");
  return vcb::prof_read(ms_by_category, launches_by_category);
}

}  // extern "C"
