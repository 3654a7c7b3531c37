// Engine: weight packing, workspaces, paged KV cache, and the launch sequences for the vision towers, projectors,
// splice, LLaMA prefill and graph-captured greedy decode.  This is the B200-side replacement of everything
// `model.generate(input_ids, images=, segs=, depths=)` executes in the reference:
//   CLIPVisionTower.forward            vcoder_llava/model/multimodal_encoder/clip_encoder.py:39-51
//   encode_* + splice                  vcoder_llava/model/vcoder_ds_llava_arch.py:106-314, vcoder_llava_arch.py:131-296
//   VCoder*LlavaLlamaForCausalLM.forward + greedy loop   language_model/vcoder_ds_llava_llama.py:57-142
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace vcb {

namespace {

constexpr int kNormChainDefault = 0;   // VCB_DECODE_NORM_CHAIN=1: decode step without rmsnorm kernels (measured slower: DESIGN.md 3.1)
constexpr bool kPackedDefault = false;  // VCB_SKINNY_PACKED unset: decode streams the row-major weights through the tensor map

int dev_alloc(Engine* e, DevBuf* b, size_t bytes) {
  bytes = (bytes + 255) & ~static_cast<size_t>(255);
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, bytes);
  if (err != cudaSuccess) {
    set_thread_error(std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(err));
    return kErrNoMem;
  }
  b->p = p;
  b->bytes = bytes;
  e->allocs.push_back(p);
  return kOk;
}

bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
bool starts_with(const std::string& s, const std::string& pre) { return s.compare(0, pre.size(), pre) == 0; }

long numel(const vcb_tensor& t) {
  long n = 1;
  for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
  return n;
}

}  // namespace

int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  VCB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
  return kOk;
}

StreamGuard::StreamGuard(Engine* e, cudaStream_t caller) : st(caller), e_(e), caller_(caller), swapped_(false) {
  if (caller == nullptr || caller == cudaStreamLegacy) {
    cudaEventRecord(e->ev_in, caller);
    cudaStreamWaitEvent(e->own_stream, e->ev_in, 0);
    st = e->own_stream;
    swapped_ = true;
  }
}
StreamGuard::~StreamGuard() {
  if (swapped_) {
    cudaEventRecord(e_->ev_out, e_->own_stream);
    cudaStreamWaitEvent(caller_, e_->ev_out, 0);
  }
}

Engine::~Engine() {
  if (graph) cudaGraphExecDestroy(graph);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (ev_in) cudaEventDestroy(ev_in);
  if (ev_out) cudaEventDestroy(ev_out);
  for (cudaEvent_t ev : stream_ev) if (ev) cudaEventDestroy(ev);
  if (stream_host) cudaFreeHost(stream_host);
  for (void* p : allocs) cudaFree(p);
  if (plan_host) cudaFreeHost(plan_host);
  if (bt_host) cudaFreeHost(bt_host);
  if (depth_host) cudaFreeHost(depth_host);
  if (fp_host) cudaFreeHost(fp_host);
}

// ------------------------------------------------------------------------------------------------------------
// create
// ------------------------------------------------------------------------------------------------------------
int engine_create(const vcb_config* c, int device, Engine** out) {
  VCB_CHECK(c != nullptr && out != nullptr, kErrInvalid, "vcb_create: null argument");
  VCB_CHECK(c->struct_size == static_cast<int32_t>(sizeof(vcb_config)), kErrInvalid, "vcb_create: struct_size mismatch");
  VCB_CHECK(c->variant >= VCB_VARIANT_VCODER && c->variant <= VCB_VARIANT_VCODER_IT, kErrInvalid, "bad variant");
  VCB_CHECK(c->vit_hidden == c->vit_heads * 64, kErrInvalid, "CLIP head_dim must be 64");
  VCB_CHECK(c->llm_hidden == c->llm_heads * 128, kErrInvalid, "LLaMA head_dim must be 128");
  VCB_CHECK(c->vit_hidden % 64 == 0 && c->vit_ffn % 64 == 0 && c->llm_hidden % 64 == 0 && c->llm_ffn % 64 == 0,
            kErrInvalid, "hidden / ffn sizes must be multiples of 64");
  VCB_CHECK(c->llm_vocab % 16 == 0, kErrInvalid, "vocab must be a multiple of 16");
  VCB_CHECK(c->vit_image_size % c->vit_patch == 0, kErrInvalid, "image size must be a multiple of the patch size");
  VCB_CHECK(c->kv_page_tokens == 64, kErrInvalid, "kv_page_tokens must be 64");
  VCB_CHECK(c->max_batch >= 1 && c->max_batch <= 32, kErrInvalid, "max_batch must be in [1, 32]");
  VCB_CHECK(c->max_seq_len >= 16, kErrInvalid, "max_seq_len too small");
  VCB_CUDA(cudaSetDevice(device));
  Engine* e = new Engine();
  e->cfg = *c;
  e->device = device;
  if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    delete e;
    set_thread_error("stream / event creation failed");
    return kErrCuda;
  }
  const int npr = c->vit_image_size / c->vit_patch;
  e->P = npr * npr;
  e->S = e->P + 1;
  e->Kpe = ((3 * c->vit_patch * c->vit_patch + 63) / 64) * 64;
  e->vit_used = c->vit_select_layer >= 0 ? c->vit_select_layer : c->vit_layers + 1 + c->vit_select_layer;
  if (e->vit_used < 0 || e->vit_used > c->vit_layers) {
    delete e;
    set_thread_error("vit_select_layer out of range");
    return kErrInvalid;
  }
  e->max_pages_per_seq = (c->max_seq_len + c->kv_page_tokens - 1) / c->kv_page_tokens;
  e->n_pages = e->max_pages_per_seq * c->max_batch;
  e->vit_max_images = 3 * c->max_batch;
  e->max_rows = static_cast<long>(c->max_batch) * c->max_seq_len;

  const long D = c->vit_hidden, F = c->vit_ffn, H = c->llm_hidden, FF = c->llm_ffn, V = c->llm_vocab;
  const long NI = e->vit_max_images, S = e->S, B = c->max_batch;
  auto A = [&](DevBuf* b, long halves) { return dev_alloc(e, b, static_cast<size_t>(halves) * 2); };
  int rc = kOk;
#define ALLOC(buf, n)                \
  if (rc == kOk) rc = A(&(buf), (n));
  // weights
  ALLOC(e->patch_w, D * e->Kpe);
  ALLOC(e->cls, D);
  ALLOC(e->pos, S * D);
  ALLOC(e->preln_w, D);
  ALLOC(e->preln_b, D);
  e->vit.resize(e->vit_used);
  for (auto& l : e->vit) {
    ALLOC(l.qkv_w, 3 * D * D); ALLOC(l.qkv_b, 3 * D); ALLOC(l.out_w, D * D); ALLOC(l.out_b, D);
    ALLOC(l.ln1_w, D); ALLOC(l.ln1_b, D); ALLOC(l.ln2_w, D); ALLOC(l.ln2_b, D);
    ALLOC(l.fc1_w, F * D); ALLOC(l.fc1_b, F); ALLOC(l.fc2_w, D * F); ALLOC(l.fc2_b, D);
  }
  for (ProjW* p : {&e->mm, &e->seg}) {
    ALLOC(p->w1, H * D); ALLOC(p->b1, H); ALLOC(p->w2, H * H); ALLOC(p->b2, H);
  }
  ALLOC(e->embed, V * H);
  ALLOC(e->final_norm, H);
  ALLOC(e->lm_head, V * H);
  e->llm.resize(c->llm_layers);
  {
    // second, tile-major copy of every matrix the decode step streams: costs the LLM weight bytes again
    // (13.2 GB at 7B, 25.7 GB at 13B of 180 GB) and turns the stream into linear 8 KiB bulk copies
    const char* ev = getenv("VCB_SKINNY_PACKED");
    e->use_packed = ev ? atoi(ev) != 0 : kPackedDefault;
  }
  for (auto& l : e->llm) {
    ALLOC(l.qkv_w, 3 * H * H); ALLOC(l.o_w, H * H); ALLOC(l.gu_w, 2 * FF * H); ALLOC(l.down_w, H * FF);
    ALLOC(l.ln1, H); ALLOC(l.ln2, H);
    if (e->use_packed) { ALLOC(l.qkv_p, 3 * H * H); ALLOC(l.o_p, H * H); ALLOC(l.gu_p, 2 * FF * H); ALLOC(l.down_p, H * FF); }
  }
  if (e->use_packed) { ALLOC(e->lm_head_p, V * H); }
  // vision workspaces
  ALLOC(e->pix, NI * 3 * c->vit_image_size * c->vit_image_size);
  ALLOC(e->im2col, NI * e->P * e->Kpe);
  ALLOC(e->vh, NI * S * D); ALLOC(e->vn, NI * S * D); ALLOC(e->vqkv, NI * S * 3 * D); ALLOC(e->vattn, NI * S * D);
  ALLOC(e->vmid, NI * S * F);
  ALLOC(e->pmid, NI * S * H); ALLOC(e->feats, NI * S * H);
  // LLM workspaces
  ALLOC(e->x, e->max_rows * H); ALLOC(e->xn, e->max_rows * H); ALLOC(e->qkv, e->max_rows * 3 * H);
  ALLOC(e->attn, e->max_rows * H); ALLOC(e->mid, e->max_rows * FF);
  ALLOC(e->xd, B * H); ALLOC(e->xnd, B * H); ALLOC(e->qkvd, B * 3 * H); ALLOC(e->attnd, B * H); ALLOC(e->midd, B * FF);
  ALLOC(e->logits, B * V);
  const int max_splits = attn_decode_splits(c->max_seq_len);
  if (rc == kOk) rc = dev_alloc(e, &e->next_ids, B * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->trace, B * c->max_seq_len * 4);
  // flash-decoding partial states: [B*heads][splits] for the split kernel, [B*heads][slots(B)] for the page-streaming
  // kernel (slots(B) = resident CTAs / (B*heads) + 2, so B*heads*slots(B) <= resident CTAs + 2*B*heads for every B)
  const long part_entries = std::max<long>(B * c->llm_heads * max_splits,
                                           static_cast<long>(attn_decode_stream_slots(1, c->llm_heads)) * c->llm_heads + 2L * B * c->llm_heads);
  if (rc == kOk) rc = dev_alloc(e, &e->part_o, part_entries * 128 * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->part_ml, part_entries * 2 * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->attn_sync, B * c->llm_heads * 4);
  if (rc == kOk) cudaMemset(e->attn_sync.p, 0, e->attn_sync.bytes);
  if (rc == kOk) rc = dev_alloc(e, &e->plan_dev, e->max_rows * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->depth_sums, B * 4);
  const size_t page_halves = static_cast<size_t>(c->llm_heads) * c->kv_page_tokens * 128;
  if (rc == kOk) rc = dev_alloc(e, &e->k_pool, static_cast<size_t>(c->llm_layers) * e->n_pages * page_halves * 2);
  if (rc == kOk) rc = dev_alloc(e, &e->v_pool, static_cast<size_t>(c->llm_layers) * e->n_pages * page_halves * 2);
  if (rc == kOk) rc = dev_alloc(e, &e->block_table, static_cast<size_t>(B) * e->max_pages_per_seq * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->rope, static_cast<size_t>(c->max_seq_len) * 64 * sizeof(float2));
  if (rc == kOk) rc = dev_alloc(e, &e->counters, 16);
  if (rc == kOk) rc = dev_alloc(e, &e->sample_params, 32);
  if (rc == kOk) rc = dev_alloc(e, &e->fp_dev, 8);
  if (rc == kOk) rc = dev_alloc(e, &e->norm_parts0, (96 + 8) * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->norm_parts_a, static_cast<size_t>(H / 16) * 32 * 4);
  if (rc == kOk) rc = dev_alloc(e, &e->norm_parts_b, static_cast<size_t>(H / 16) * 32 * 4);
  if (rc == kOk) {
    cudaMemset(e->norm_parts0.p, 0, e->norm_parts0.bytes);
    cudaMemset(e->norm_parts_a.p, 0, e->norm_parts_a.bytes);
    cudaMemset(e->norm_parts_b.p, 0, e->norm_parts_b.bytes);
  }
#undef ALLOC
  if (rc != kOk) { delete e; return rc; }
  if (cudaMallocHost(&e->plan_host, e->max_rows * 4) != cudaSuccess ||
      cudaMallocHost(&e->bt_host, static_cast<size_t>(B) * e->max_pages_per_seq * 4) != cudaSuccess ||
      cudaMallocHost(&e->depth_host, B * 4) != cudaSuccess ||
      cudaMallocHost(&e->fp_host, 8) != cudaSuccess ||
      cudaMallocHost(&e->stream_host, static_cast<size_t>(Engine::kStreamRing) * B * 4) != cudaSuccess) {
    delete e;
    set_thread_error("cudaMallocHost failed");
    return kErrNoMem;
  }
  {
    std::vector<float2> tab(static_cast<size_t>(c->max_seq_len) * 64);
    build_rope_table(tab.data(), c->max_seq_len, 128, c->llm_rope_theta);
    if (cudaMemcpy(e->rope.p, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
      delete e;
      set_thread_error("rope table upload failed");
      return kErrCuda;
    }
    cudaMemset(e->patch_w.p, 0, e->patch_w.bytes);
    cudaMemset(e->counters.p, 0, 16);
    // KV pages start as zeros: the unwritten tail of a sequence's last page is multiplied by exact-zero probabilities in
    // the prefill attention MMA, which must never meet a NaN bit pattern left over in fresh device memory
    cudaMemset(e->k_pool.p, 0, e->k_pool.bytes);
    cudaMemset(e->v_pool.p, 0, e->v_pool.bytes);
    for (cudaEvent_t& ev : e->stream_ev) {
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
        delete e;
        set_thread_error("event creation failed");
        return kErrCuda;
      }
    }
  }
  *out = e;
  return kOk;
}

// ------------------------------------------------------------------------------------------------------------
// weights: names are the reference model's state_dict() keys
// ------------------------------------------------------------------------------------------------------------
namespace {
const char* kVT = "model.vision_tower.vision_tower.vision_model.";

int put(const vcb_tensor& t, const DevBuf& dst, long offset_halves, long expect) {
  VCB_CHECK(numel(t) == expect, kErrInvalid, std::string("shape mismatch for ") + t.name);
  VCB_CUDA(cudaMemcpy(dst.as<__half>() + offset_halves, t.data, static_cast<size_t>(expect) * 2, cudaMemcpyDefault));
  return kOk;
}
}  // namespace

int engine_bind(Engine* e, const vcb_tensor* ts, int n) {
  VCB_CHECK(e != nullptr && (ts != nullptr || n == 0), kErrInvalid, "vcb_bind_weights: null argument");
  VCB_CUDA(cudaSetDevice(e->device));
  const vcb_config& c = e->cfg;
  const long D = c.vit_hidden, F = c.vit_ffn, H = c.llm_hidden, FF = c.llm_ffn, V = c.llm_vocab;
  for (int i = 0; i < n; ++i) {
    const vcb_tensor& t = ts[i];
    VCB_CHECK(t.name != nullptr && t.data != nullptr, kErrInvalid, "vcb_bind_weights: null name/data");
    const std::string name(t.name);
    int rc = kOk;
    bool used = true;
    if (starts_with(name, kVT)) {
      const std::string r = name.substr(std::strlen(kVT));
      if (r == "embeddings.class_embedding") rc = put(t, e->cls, 0, D);
      else if (r == "embeddings.position_embedding.weight") rc = put(t, e->pos, 0, e->S * D);
      else if (r == "embeddings.patch_embedding.weight") {
        const long kreal = 3L * c.vit_patch * c.vit_patch;
        VCB_CHECK(numel(t) == D * kreal, kErrInvalid, "shape mismatch for patch_embedding.weight");
        VCB_CUDA(cudaMemcpy2D(e->patch_w.p, e->Kpe * 2, t.data, kreal * 2, kreal * 2, D, cudaMemcpyDefault));
      } else if (r == "pre_layrnorm.weight") rc = put(t, e->preln_w, 0, D);
      else if (r == "pre_layrnorm.bias") rc = put(t, e->preln_b, 0, D);
      else if (starts_with(r, "encoder.layers.")) {
        const size_t dot = r.find('.', 15);
        const int li = std::atoi(r.substr(15, dot - 15).c_str());
        const std::string k = r.substr(dot + 1);
        if (li >= e->vit_used) used = false;  // layers past the tapped hidden state are dead
        else {
          VitLayerW& l = e->vit[li];
          if (k == "self_attn.q_proj.weight") rc = put(t, l.qkv_w, 0, D * D);
          else if (k == "self_attn.k_proj.weight") rc = put(t, l.qkv_w, D * D, D * D);
          else if (k == "self_attn.v_proj.weight") rc = put(t, l.qkv_w, 2 * D * D, D * D);
          else if (k == "self_attn.q_proj.bias") rc = put(t, l.qkv_b, 0, D);
          else if (k == "self_attn.k_proj.bias") rc = put(t, l.qkv_b, D, D);
          else if (k == "self_attn.v_proj.bias") rc = put(t, l.qkv_b, 2 * D, D);
          else if (k == "self_attn.out_proj.weight") rc = put(t, l.out_w, 0, D * D);
          else if (k == "self_attn.out_proj.bias") rc = put(t, l.out_b, 0, D);
          else if (k == "layer_norm1.weight") rc = put(t, l.ln1_w, 0, D);
          else if (k == "layer_norm1.bias") rc = put(t, l.ln1_b, 0, D);
          else if (k == "layer_norm2.weight") rc = put(t, l.ln2_w, 0, D);
          else if (k == "layer_norm2.bias") rc = put(t, l.ln2_b, 0, D);
          else if (k == "mlp.fc1.weight") rc = put(t, l.fc1_w, 0, F * D);
          else if (k == "mlp.fc1.bias") rc = put(t, l.fc1_b, 0, F);
          else if (k == "mlp.fc2.weight") rc = put(t, l.fc2_w, 0, D * F);
          else if (k == "mlp.fc2.bias") rc = put(t, l.fc2_b, 0, D);
          else used = false;
        }
      } else used = false;  // post_layernorm.*, position_ids: dead at inference
    } else if (starts_with(name, "model.mm_projector.") || starts_with(name, "model.seg_mm_projector.")) {
      ProjW& p = starts_with(name, "model.mm_projector.") ? e->mm : e->seg;
      if (ends_with(name, ".0.weight")) rc = put(t, p.w1, 0, H * D);
      else if (ends_with(name, ".0.bias")) rc = put(t, p.b1, 0, H);
      else if (ends_with(name, ".2.weight")) rc = put(t, p.w2, 0, H * H);
      else if (ends_with(name, ".2.bias")) rc = put(t, p.b2, 0, H);
      else used = false;
    } else if (name == "model.embed_tokens.weight") rc = put(t, e->embed, 0, V * H);
    else if (name == "model.norm.weight") rc = put(t, e->final_norm, 0, H);
    else if (name == "lm_head.weight") rc = put(t, e->lm_head, 0, V * H);
    else if (starts_with(name, "model.layers.")) {
      const size_t dot = name.find('.', 13);
      const int li = std::atoi(name.substr(13, dot - 13).c_str());
      const std::string k = name.substr(dot + 1);
      VCB_CHECK(li >= 0 && li < c.llm_layers, kErrInvalid, std::string("layer index out of range: ") + name);
      LlmLayerW& l = e->llm[li];
      if (k == "self_attn.q_proj.weight") rc = put(t, l.qkv_w, 0, H * H);
      else if (k == "self_attn.k_proj.weight") rc = put(t, l.qkv_w, H * H, H * H);
      else if (k == "self_attn.v_proj.weight") rc = put(t, l.qkv_w, 2 * H * H, H * H);
      else if (k == "self_attn.o_proj.weight") rc = put(t, l.o_w, 0, H * H);
      else if (k == "mlp.down_proj.weight") rc = put(t, l.down_w, 0, H * FF);
      else if (k == "mlp.gate_proj.weight" || k == "mlp.up_proj.weight") {
        // interleave 8 gate rows | 8 up rows so the SwiGLU pair of a feature lands in one thread's registers
        VCB_CHECK(numel(t) == FF * H, kErrInvalid, std::string("shape mismatch for ") + name);
        const long off = (k == "mlp.up_proj.weight") ? 8 * H : 0;
        VCB_CUDA(cudaMemcpy2D(l.gu_w.as<__half>() + off, 16 * H * 2, t.data, 8 * H * 2, 8 * H * 2, FF / 8,
                              cudaMemcpyDefault));
      } else if (k == "input_layernorm.weight") rc = put(t, l.ln1, 0, H);
      else if (k == "post_attention_layernorm.weight") rc = put(t, l.ln2, 0, H);
      else used = false;  // rotary inv_freq buffers etc.
    } else {
      // accepted and ignored: model.depth_mm_projector.* (never called), model.mm2_projector.* (guard is
      // always False, F4), model.vcoder_lm_emb.weight (overwritten by a clone of embed_tokens every prefill, F7)
      used = false;
    }
    VCB_TRY(rc);
    if (used) {
      e->bound.insert(name);
      e->packed_valid = false;
      // new weights: neither the projected features nor the KV entries a later vcb_prefill_cached could reuse are valid
      e->feat_key = 0;
      e->seq_plan_len = 0;
    }
  }
  return kOk;
}

int engine_ensure_packed(Engine* e, cudaStream_t st) {
  if (!e->use_packed || e->packed_valid) return kOk;
  const vcb_config& c = e->cfg;
  const int H = c.llm_hidden, FF = c.llm_ffn, V = c.llm_vocab;
  for (auto& l : e->llm) {
    VCB_TRY(pack_skinny_weights(l.qkv_w.as<__half>(), l.qkv_p.as<__half>(), 3 * H, H, st));
    VCB_TRY(pack_skinny_weights(l.o_w.as<__half>(), l.o_p.as<__half>(), H, H, st));
    VCB_TRY(pack_skinny_weights(l.gu_w.as<__half>(), l.gu_p.as<__half>(), 2 * FF, H, st));
    VCB_TRY(pack_skinny_weights(l.down_w.as<__half>(), l.down_p.as<__half>(), H, FF, st));
  }
  VCB_TRY(pack_skinny_weights(e->lm_head.as<__half>(), e->lm_head_p.as<__half>(), V, H, st));
  e->packed_valid = true;
  return kOk;
}

int engine_check_bound(Engine* e, bool need_vision, bool need_llm) {
  const vcb_config& c = e->cfg;
  std::vector<std::string> need;
  if (need_vision) {
    const std::string vt(kVT);
    for (const char* s : {"embeddings.class_embedding", "embeddings.position_embedding.weight",
                          "embeddings.patch_embedding.weight", "pre_layrnorm.weight", "pre_layrnorm.bias"})
      need.push_back(vt + s);
    for (int i = 0; i < e->vit_used; ++i) {
      const std::string p = vt + "encoder.layers." + std::to_string(i) + ".";
      for (const char* s : {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                            "self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias",
                            "self_attn.out_proj.weight", "self_attn.out_proj.bias", "layer_norm1.weight",
                            "layer_norm1.bias", "layer_norm2.weight", "layer_norm2.bias", "mlp.fc1.weight",
                            "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"})
        need.push_back(p + s);
    }
  }
  if (need_llm) {
    for (const char* pj : {"model.mm_projector.", "model.seg_mm_projector."}) {
      if (c.variant == VCB_VARIANT_LLAVA && std::strcmp(pj, "model.seg_mm_projector.") == 0) continue;  // llava_arch.py:36-38
      for (const char* s : {"0.weight", "0.bias", "2.weight", "2.bias"}) need.push_back(std::string(pj) + s);
    }
    need.push_back("model.embed_tokens.weight");
    need.push_back("model.norm.weight");
    need.push_back("lm_head.weight");
    for (int i = 0; i < c.llm_layers; ++i) {
      const std::string p = "model.layers." + std::to_string(i) + ".";
      for (const char* s : {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                            "self_attn.o_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight",
                            "mlp.down_proj.weight", "input_layernorm.weight", "post_attention_layernorm.weight"})
        need.push_back(p + s);
    }
  }
  for (const auto& s : need) {
    if (!e->bound.count(s)) {
      set_thread_error("weight not bound: " + s);
      return kErrState;
    }
  }
  return kOk;
}

// ------------------------------------------------------------------------------------------------------------
// vision tower: n images in ONE batch through the shared CLIP weights, result = residual stream after
// encoder layer `vit_used` (hidden_states[select_layer]) for all S tokens, in e->vh
// ------------------------------------------------------------------------------------------------------------
int engine_tower(Engine* e, const __half* px, int n, cudaStream_t st) {
  const vcb_config& c = e->cfg;
  VCB_CHECK(n >= 1 && n <= e->vit_max_images, kErrInvalid, "tower: too many images for this handle (3*max_batch)");
  const int D = c.vit_hidden, F = c.vit_ffn, S = e->S, P = e->P;
  const long rows = static_cast<long>(n) * S;
  __half* h = e->vh.as<__half>();
  __half* hn = e->vn.as<__half>();
  // patch embedding: im2col + GEMM whose epilogue adds the position embedding and scatters past the CLS rows
  VCB_TRY(im2col_patches(px, e->im2col.as<__half>(), n, c.vit_image_size, c.vit_patch, e->Kpe, st));
  {
    GemmEpilogue ep;
    ep.C = h; ep.ldc = D;
    ep.residual = e->pos.as<__half>(); ep.ldr = D;
    ep.out_group = P; ep.out_skip = 1; ep.out_off = 1;
    ep.res_mod = P; ep.res_off = 1;
    VCB_TRY(gemm_f16(e->im2col.as<__half>(), e->Kpe, e->patch_w.as<__half>(), e->Kpe, n * P, D, e->Kpe, ep, 0, st));
  }
  VCB_TRY(write_cls_rows(e->cls.as<__half>(), e->pos.as<__half>(), h, n, S, D, st));
  VCB_TRY(layernorm_f16(h, e->preln_w.as<__half>(), e->preln_b.as<__half>(), h, rows, D, c.vit_ln_eps, st));
  for (int li = 0; li < e->vit_used; ++li) {
    const VitLayerW& l = e->vit[li];
    VCB_TRY(layernorm_f16(h, l.ln1_w.as<__half>(), l.ln1_b.as<__half>(), hn, rows, D, c.vit_ln_eps, st));
    {
      GemmEpilogue ep;
      ep.C = e->vqkv.as<__half>(); ep.ldc = 3 * D; ep.bias = l.qkv_b.as<__half>();
      VCB_TRY(gemm_f16(hn, D, l.qkv_w.as<__half>(), D, rows, 3 * D, D, ep, 0, st));
    }
    {
      AttnPrefillParams a;
      a.q = e->vqkv.as<__half>(); a.k = a.q + D; a.v = a.q + 2 * D;
      a.q_stride = 3 * D; a.kv_stride = 3 * D;
      a.out = e->vattn.as<__half>(); a.out_stride = D;
      a.batch = n; a.n_heads = c.vit_heads; a.seq_len = S; a.head_dim = 64;
      a.scale = 0.125f;
      VCB_TRY(attn_vit(a, st));
    }
    {
      GemmEpilogue ep;
      ep.C = h; ep.ldc = D; ep.bias = l.out_b.as<__half>(); ep.residual = h; ep.ldr = D;
      VCB_TRY(gemm_f16(e->vattn.as<__half>(), D, l.out_w.as<__half>(), D, rows, D, D, ep, 0, st));
    }
    VCB_TRY(layernorm_f16(h, l.ln2_w.as<__half>(), l.ln2_b.as<__half>(), hn, rows, D, c.vit_ln_eps, st));
    {
      GemmEpilogue ep;
      ep.C = e->vmid.as<__half>(); ep.ldc = F; ep.bias = l.fc1_b.as<__half>(); ep.act = kActQuickGelu;
      VCB_TRY(gemm_f16(hn, D, l.fc1_w.as<__half>(), D, rows, F, D, ep, 0, st));
    }
    {
      GemmEpilogue ep;
      ep.C = h; ep.ldc = D; ep.bias = l.fc2_b.as<__half>(); ep.residual = h; ep.ldr = D;
      VCB_TRY(gemm_f16(e->vmid.as<__half>(), F, l.fc2_w.as<__half>(), F, rows, D, F, ep, 0, st));
    }
  }
  return kOk;
}

// mlp2x_gelu on rows [row_begin, row_end) of e->vh -> e->feats (CLS rows ride along; the splice skips them)
int engine_project(Engine* e, int row_begin, int row_end, const ProjW& w, cudaStream_t st) {
  const vcb_config& c = e->cfg;
  const int D = c.vit_hidden, H = c.llm_hidden;
  const int rows = row_end - row_begin;
  if (rows <= 0) return kOk;
  __half* mid = e->pmid.as<__half>() + static_cast<long>(row_begin) * H;
  {
    GemmEpilogue ep;
    ep.C = mid; ep.ldc = H; ep.bias = w.b1.as<__half>(); ep.act = kActGeluErf;
    VCB_TRY(gemm_f16(e->vh.as<__half>() + static_cast<long>(row_begin) * D, D, w.w1.as<__half>(), D, rows, H, D, ep, 0, st));
  }
  {
    GemmEpilogue ep;
    ep.C = e->feats.as<__half>() + static_cast<long>(row_begin) * H; ep.ldc = H; ep.bias = w.b2.as<__half>();
    VCB_TRY(gemm_f16(mid, H, w.w2.as<__half>(), H, rows, H, H, ep, 0, st));
  }
  return kOk;
}

// ------------------------------------------------------------------------------------------------------------
// splice planner (host)
// ------------------------------------------------------------------------------------------------------------
int splice_plan_host(int variant, const int64_t* ids, int B, int T, int n_patches, int has_seg,
                     const uint8_t* nonzero, std::vector<int32_t>* plan, int* out_len, bool* flag_needed) {
  // `nonzero`: VCoder-DS per-sample mean(depth) != 0 (is_depth_zero, vcoder_ds_llava_arch.py:161), VCoder-IT per-sample
  // mean(seg) != 0 (is_seg_zero, vcoder_it_llava_arch.py:148); null = that modality was not passed
  const bool ds = variant == VCB_VARIANT_VCODER_DS, it = variant == VCB_VARIANT_VCODER_IT;
  const bool vc = variant == VCB_VARIANT_VCODER;
  const int tokens = n_patches + 1;
  std::vector<std::vector<int32_t>> rows(B);
  int cur_img = 0, cur_seg = 0, cur_depth = 0;
  if (flag_needed) *flag_needed = false;
  auto push_text = [&](std::vector<int32_t>& r, const int64_t* p, int n) -> int {
    for (int i = 0; i < n; ++i) {
      if (p[i] < 0) {
        set_thread_error("special token id would index the embedding table (reference: IndexError)");
        return kErrSpecialIndex;
      }
      r.push_back(static_cast<int32_t>(p[i]));
    }
    return kOk;
  };
  auto push_feat = [&](std::vector<int32_t>& r, int kind, int idx) {
    const long base = (static_cast<long>(kind) * B + idx) * tokens + 1;  // +1: skip the CLS row
    for (int p = 0; p < n_patches; ++p) r.push_back(static_cast<int32_t>(-(base + p) - 2));
  };
  for (int b = 0; b < B; ++b) {
    const int64_t* row = ids + static_cast<long>(b) * T;
    std::vector<int32_t>& r = rows[b];
    int n_img = 0, n_seg = 0;
    for (int t = 0; t < T; ++t) { n_img += row[t] == VCB_IMAGE_TOKEN_INDEX; n_seg += row[t] == VCB_SEG_TOKEN_INDEX; }
    // ds :181, vcoder :187, llava_arch.py:118 / vcoder_it_llava_arch.py:169
    const bool dummy = ds ? (n_img == 0 && n_seg == 0) : vc ? (n_img == 0 || n_seg == 0) : (n_img == 0);
    if (dummy) {
      VCB_TRY(push_text(r, row, T));
      ++cur_img; ++cur_seg; ++cur_depth;
      continue;
    }
    int pos = 0;  // start of the not-yet-consumed suffix
    auto find = [&](int64_t tok) { for (int t = pos; t < T; ++t) if (row[t] == tok) return t; return -1; };
    for (int t; (t = find(VCB_IMAGE_TOKEN_INDEX)) >= 0;) {
      VCB_TRY(push_text(r, row + pos, t - pos));
      VCB_CHECK(cur_img < B, kErrInvalid, "more <image> tokens than images");
      push_feat(r, 0, cur_img++);
      pos = t + 1;
    }
    if ((ds || vc) && has_seg) {
      for (int t; (t = find(VCB_SEG_TOKEN_INDEX)) >= 0;) {
        if (vc) VCB_TRY(push_text(r, row + pos, t - pos));  // DS drops these ids
        VCB_CHECK(cur_seg < B, kErrInvalid, "more <seg> tokens than seg images");
        push_feat(r, 1, cur_seg++);
        pos = t + 1;
      }
    } else if (it) {
      VCB_CHECK(cur_seg < B, kErrInvalid, "more seg rows consumed than seg images");
      if (has_seg && flag_needed) *flag_needed = true;
      const bool seg_nz = has_seg && nonzero != nullptr && nonzero[cur_seg] != 0;
      if (seg_nz) {
        if (find(VCB_SEG_TOKEN_INDEX) < 0) ++cur_seg;  // :212-215, an empty slice of the seg features
        for (int t; (t = find(VCB_SEG_TOKEN_INDEX)) >= 0;) {
          VCB_CHECK(cur_seg < B, kErrInvalid, "more <seg> tokens than seg images");
          push_feat(r, 1, cur_seg++);  // the ids before <seg> are dropped, as in DS (:216-226)
          pos = t + 1;
        }
      } else {
        ++cur_seg;  // :227-228
      }
    }
    if (ds) {
      const bool has_depth_tok = find(VCB_DEPTH_TOKEN_INDEX) >= 0;
      if (has_depth_tok && flag_needed) *flag_needed = true;
      const bool nz = nonzero != nullptr && nonzero[cur_depth] != 0;
      if (nz) {
        for (int t; (t = find(VCB_DEPTH_TOKEN_INDEX)) >= 0;) {
          VCB_TRY(push_text(r, row + pos, t - pos));
          VCB_CHECK(cur_depth < B, kErrInvalid, "more <depth> tokens than depth images");
          push_feat(r, 2, cur_depth++);
          pos = t + 1;
        }
      } else {
        ++cur_depth;
      }
    }
    VCB_TRY(push_text(r, row + pos, T - pos));
  }
  const size_t len = rows[0].size();
  for (int b = 1; b < B; ++b) {
    if (rows[b].size() != len) {
      set_thread_error("ragged spliced batch (the reference raises NameError: _new_labels)");
      return kErrRagged;
    }
  }
  plan->clear();
  for (int b = 0; b < B; ++b) plan->insert(plan->end(), rows[b].begin(), rows[b].end());
  *out_len = static_cast<int>(len);
  return kOk;
}

// ------------------------------------------------------------------------------------------------------------
// prefill
// ------------------------------------------------------------------------------------------------------------
namespace {

// next token of every sequence from e->logits: greedy argmax, or HF sample() when vcb_set_sampling enabled it
int select_next(Engine* e, int B, const int32_t* step_dev, cudaStream_t st) {
  const vcb_config& c = e->cfg;
  if (e->sampling)
    return sample_rows(e->logits.as<__half>(), B, c.llm_vocab, e->sample_params.p, step_dev, 0,
                       e->next_ids.as<int32_t>(), e->trace.as<int32_t>(), c.max_seq_len, nullptr, st);
  return argmax_rows(e->logits.as<__half>(), B, c.llm_vocab, e->next_ids.as<int32_t>(), e->trace.as<int32_t>(),
                     c.max_seq_len, step_dev, st);
}

__global__ void depth_sum_kernel(const __half* px, long per_image, float* sums) {
  const int b = blockIdx.y;
  const __half* p = px + b * per_image;
  float s = 0.f;
  for (long i = blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += static_cast<long>(gridDim.x) * blockDim.x)
    s += __half2float(p[i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(sums + b, s);
}

// f3: order-independent 64-bit fingerprint of a pixel buffer (sum of position-keyed splitmix64 hashes of its 8-byte
// words): identifies "the same images as the previous turn" without a host copy of the pixels
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__global__ void fingerprint_kernel(const uint64_t* __restrict__ p, long n_words, unsigned long long* out) {
  unsigned long long acc = 0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_words;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    acc += mix64(p[i] ^ (0x9e3779b97f4a7c15ULL * static_cast<uint64_t>(i + 1)));
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out, acc);
}

int ensure_pages(Engine* e, int B, int total_tokens, cudaStream_t st) {
  const int PT = e->cfg.kv_page_tokens;
  const int need = (total_tokens + PT - 1) / PT;
  VCB_CHECK(need <= e->max_pages_per_seq, kErrNoMem, "KV cache: sequence exceeds max_seq_len");
  if (need <= e->pages_per_seq_alloc) return kOk;
  // pages are handed out sequence-major from the pool; a page id is never shared between sequences
  for (int b = 0; b < B; ++b)
    for (int pg = 0; pg < e->max_pages_per_seq; ++pg)
      e->bt_host[b * e->max_pages_per_seq + pg] = pg < need ? b * e->max_pages_per_seq + pg : 0;
  e->pages_per_seq_alloc = need;
  VCB_TRY(copy_any(e->block_table.p, e->bt_host, static_cast<size_t>(B) * e->max_pages_per_seq * 4, st));
  return kOk;
}

}  // namespace

int prefix_match_host(const int32_t* cached, int cached_len, const int32_t* plan, int plan_len, int B, int features_same,
                      int limit) {
  int best = std::min(std::min(cached_len, plan_len), limit);
  for (int b = 0; b < B && best > 0; ++b) {
    const int32_t* c = cached + static_cast<long>(b) * cached_len;
    const int32_t* n = plan + static_cast<long>(b) * plan_len;
    int i = 0;
    while (i < best && c[i] == n[i] && (features_same || n[i] > -2)) ++i;
    best = i;
  }
  return std::max(best, 0);
}

int engine_prefill(Engine* e, const int64_t* ids, int B, int T, const void* images, const void* segs,
                   const void* depths, bool reuse, void* next_ids_out, void* last_logits_out, void* embeds_out,
                   int* spliced_len, int* reused_len, int* towers_skipped, cudaStream_t st, const PrefillExtra* extra) {
  const vcb_config& c = e->cfg;
  VCB_CHECK(ids != nullptr && B >= 1 && B <= c.max_batch && T >= 1, kErrInvalid, "vcb_prefill: bad ids / batch");
  VCB_CUDA(cudaSetDevice(e->device));
  const bool ds = c.variant == VCB_VARIANT_VCODER_DS;
  const bool multimodal = images != nullptr;  // reference: `images is None` -> plain LLaMA path (arch.py:130)
  VCB_TRY(engine_check_bound(e, multimodal, true));
  VCB_TRY(engine_ensure_packed(e, st));
  VCB_CHECK(ds || depths == nullptr, kErrInvalid, "only VCoder-DS takes depth images");
  VCB_CHECK(c.variant != VCB_VARIANT_LLAVA || segs == nullptr, kErrInvalid, "LLaVA takes no seg images");
  const int H = c.llm_hidden, FF = c.llm_ffn, V = c.llm_vocab, S = e->S;
  const long img_halves = 3L * c.vit_image_size * c.vit_image_size;
  if (reused_len) *reused_len = 0;
  if (towers_skipped) *towers_skipped = 0;

  // ---- f3: what the KV cache holds right now = the previous prompt's plan + the generated tokens already fed back
  //      (token i of the trace was consumed by decode step i+1, so the last generated token is not cached yet)
  std::vector<int32_t> cached;
  int cached_len = 0;
  if (reuse && e->prefilled && e->batch == B && e->seq_plan_len > 0 && embeds_out == nullptr) {
    const int fed = e->steps_done - 1;
    cached_len = e->seq_plan_len + fed;
    cached.resize(static_cast<size_t>(B) * cached_len);
    std::vector<int32_t> gen(static_cast<size_t>(B) * std::max(fed, 1));
    if (fed > 0) {
      VCB_CUDA(cudaMemcpy2DAsync(gen.data(), static_cast<size_t>(fed) * 4, e->trace.p, static_cast<size_t>(c.max_seq_len) * 4,
                                 static_cast<size_t>(fed) * 4, B, cudaMemcpyDeviceToHost, st));
      VCB_CUDA(cudaStreamSynchronize(st));
    }
    for (int b = 0; b < B; ++b) {
      int32_t* dst = cached.data() + static_cast<long>(b) * cached_len;
      std::memcpy(dst, e->seq_plan.data() + static_cast<long>(b) * e->seq_plan_len, static_cast<size_t>(e->seq_plan_len) * 4);
      if (fed > 0) std::memcpy(dst + e->seq_plan_len, gen.data() + static_cast<long>(b) * fed, static_cast<size_t>(fed) * 4);
    }
  }
  e->prefilled = false;

  int Tp = T;
  std::vector<int32_t> plan;
  bool features_same = false;
  if (multimodal) {
    // ---- all modalities in one tower batch: [rgb | seg | depth]
    const int kinds = 1 + (segs ? 1 : 0) + ((ds && depths) ? 1 : 0);
    VCB_CHECK(depths == nullptr || segs != nullptr, kErrInvalid, "depths without segs is not reachable in the reference");
    __half* pix = e->pix.as<__half>();
    VCB_TRY(copy_any(pix, images, static_cast<size_t>(B) * img_halves * 2, st));
    if (segs) VCB_TRY(copy_any(pix + B * img_halves, segs, static_cast<size_t>(B) * img_halves * 2, st));
    if (ds && depths) VCB_TRY(copy_any(pix + 2L * B * img_halves, depths, static_cast<size_t>(B) * img_halves * 2, st));
    uint64_t key = 0;
    const size_t pix_bytes = static_cast<size_t>(kinds) * B * img_halves * 2;
    if (reuse && pix_bytes % 8 == 0) {
      VCB_CUDA(cudaMemsetAsync(e->fp_dev.p, 0, 8, st));
      fingerprint_kernel<<<4 * device_sm_count(), 256, 0, st>>>(e->pix.as<uint64_t>(), static_cast<long>(pix_bytes / 8),
                                                                e->fp_dev.as<unsigned long long>());
      VCB_CUDA(cudaGetLastError());
      VCB_TRY(copy_any(e->fp_host, e->fp_dev.p, 8, st));
      VCB_CUDA(cudaStreamSynchronize(st));
      key = *e->fp_host ^ (0x51ed270b27b4f3cfULL * static_cast<uint64_t>(kinds * 64 + B));
      if (key == 0) key = 1;
    }
    features_same = key != 0 && key == e->feat_key && e->feat_batch == B && e->feat_kinds == kinds;
    if (!features_same) {
      e->feat_key = 0;  // `feats` is about to be overwritten
      VCB_TRY(engine_tower(e, pix, kinds * B, st));
      VCB_TRY(engine_project(e, 0, B * S, e->mm, st));                // encode_images: mm_projector
      VCB_TRY(engine_project(e, B * S, kinds * B * S, e->seg, st));   // seg AND depth: seg_mm_projector (F3)
      e->feat_key = key;
      e->feat_batch = B;
      e->feat_kinds = kinds;
    } else if (towers_skipped) {
      *towers_skipped = 1;
    }
    // ---- splice plan; the is_depth_zero (DS) / is_seg_zero (IT) flags cost one host sync, paid only when the plan can
    //      depend on them (DS: a <depth> token is reachable after the seg splice; IT: seg images were passed)
    const bool it = c.variant == VCB_VARIANT_VCODER_IT;
    const bool have_flag_modality = (ds && depths) || (it && segs);
    bool need_flag = false;
    std::vector<uint8_t> nz(B, 1);
    int rc = splice_plan_host(c.variant, ids, B, T, e->P, segs != nullptr, have_flag_modality ? nz.data() : nullptr, &plan,
                              &Tp, &need_flag);
    if (need_flag && have_flag_modality) {
      const __half* flag_px = pix + (ds ? 2L : 1L) * B * img_halves;
      VCB_CUDA(cudaMemsetAsync(e->depth_sums.p, 0, B * 4, st));
      depth_sum_kernel<<<dim3(32, B), 256, 0, st>>>(flag_px, img_halves, e->depth_sums.as<float>());
      VCB_CUDA(cudaGetLastError());
      VCB_TRY(copy_any(e->depth_host, e->depth_sums.p, B * 4, st));
      VCB_CUDA(cudaStreamSynchronize(st));
      for (int b = 0; b < B; ++b) {
        const __half m = __float2half_rn(e->depth_host[b] / static_cast<float>(img_halves));
        nz[b] = (__half2float(m) != 0.f) ? 1 : 0;  // torch.mean(x) == 0 on an fp16 tensor (ds arch.py:161, it arch.py:148)
      }
      rc = splice_plan_host(c.variant, ids, B, T, e->P, segs != nullptr, nz.data(), &plan, &Tp, nullptr);
    }
    VCB_TRY(rc);
  } else {
    plan.reserve(static_cast<size_t>(B) * T);
    for (long i = 0; i < static_cast<long>(B) * T; ++i) {
      VCB_CHECK(ids[i] >= 0 && ids[i] < V, kErrSpecialIndex, "token id out of range for the embedding table");
      plan.push_back(static_cast<int32_t>(ids[i]));
    }
  }
  for (int32_t v : plan) VCB_CHECK(v < V, kErrSpecialIndex, "token id out of range for the embedding table");
  VCB_CHECK(Tp >= 1 && Tp <= c.max_seq_len, kErrNoMem, "spliced prompt exceeds max_seq_len");
  if (spliced_len) *spliced_len = Tp;

  // ---- f3: positions [0, L0) keep their KV entries; rows [L0, Tp) of every sequence run through the decoder.  The last
  //      position always runs (its hidden state feeds lm_head)
  const int L0 = cached_len > 0 ? prefix_match_host(cached.data(), cached_len, plan.data(), Tp, B, features_same ? 1 : 0, Tp - 1) : 0;
  if (reused_len) *reused_len = L0;
  const int Ts = Tp - L0;
  const long rows = static_cast<long>(B) * Ts;
  for (int b = 0; b < B; ++b)
    std::memcpy(e->plan_host + static_cast<long>(b) * Ts, plan.data() + static_cast<long>(b) * Tp + L0, static_cast<size_t>(Ts) * 4);
  VCB_TRY(copy_any(e->plan_dev.p, e->plan_host, rows * 4, st));
  __half* x = e->x.as<__half>();
  VCB_TRY(splice_gather(e->plan_dev.as<int32_t>(), e->embed.as<__half>(), e->feats.as<__half>(), x, rows, H, st));
  if (embeds_out) VCB_TRY(copy_any(embeds_out, x, rows * H * 2, st));
  if (extra && extra->embeds_only) return kOk;  // prepare_inputs_labels_for_multimodal seam: the handle stays un-prefilled
  if (extra && extra->all_logits) {
    VCB_CHECK(extra->all_logits_halves >= rows * V, kErrNoMem, "vcb_forward_logits: logits buffer smaller than B * T' * V");
    cudaPointerAttributes pa;
    VCB_CUDA(cudaPointerGetAttributes(&pa, extra->all_logits));
    VCB_CHECK(pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged, kErrInvalid,
              "vcb_forward_logits: the logits buffer must be device memory");
  }

  // ---- KV pages for the prompt (the page ids of a sequence never change, so a reused prefix stays where it is)
  if (L0 == 0) e->pages_per_seq_alloc = 0;
  VCB_TRY(ensure_pages(e, B, Tp, st));
  const size_t layer_pool = static_cast<size_t>(e->n_pages) * c.llm_heads * c.kv_page_tokens * 128;
  __half* xn = e->xn.as<__half>();
  __half* qkv = e->qkv.as<__half>();
  // VCB_PREFILL_ROPE=kernel: RoPE + KV append as a separate pass over the QKV buffer (rope_kv_kernel); default: fused
  // into the QKV GEMM's epilogue (bit-identical: the epilogue rotates the fp16-rounded values)
  static int rope_fused_env = -1;
  if (rope_fused_env < 0) { const char* ev = getenv("VCB_PREFILL_ROPE"); rope_fused_env = (ev && ev[0] == 'k') ? 0 : 1; }
  const bool fuse_prefill_rope = rope_fused_env != 0;
  for (int li = 0; li < c.llm_layers; ++li) {
    const LlmLayerW& l = e->llm[li];
    __half* kp = e->k_pool.as<__half>() + li * layer_pool;
    __half* vp = e->v_pool.as<__half>() + li * layer_pool;
    VCB_TRY(rmsnorm_f16(x, l.ln1.as<__half>(), xn, rows, H, c.llm_rms_eps, st));
    {
      GemmEpilogue ep;
      ep.C = qkv; ep.ldc = 3 * H;
      if (fuse_prefill_rope) {  // RoPE + KV append in the epilogue of the QKV GEMM (q -> qkv buffer, k / v -> page pool)
        ep.act = kActRopeKv;
        ep.rk_cos_sin = e->rope.as<float2>(); ep.rk_k_pool = kp; ep.rk_v_pool = vp;
        ep.rk_block_table = e->block_table.as<int32_t>(); ep.rk_max_pages = e->max_pages_per_seq;
        ep.rk_T = Ts; ep.rk_pos0 = L0; ep.rk_heads = c.llm_heads; ep.rk_H = H; ep.rk_page_tokens = c.kv_page_tokens;
      }
      VCB_TRY(gemm_f16(xn, H, l.qkv_w.as<__half>(), H, rows, 3 * H, H, ep, 0, st));
    }
    if (!fuse_prefill_rope)
      VCB_TRY(rope_kv_write(qkv, B, Ts, H, c.llm_heads, e->rope.as<float2>(), L0, nullptr, kp, vp,
                            e->block_table.as<int32_t>(), e->max_pages_per_seq, c.kv_page_tokens, st));
    {
      AttnPrefillParams a;
      a.q = qkv; a.q_stride = 3 * H;
      a.k_pool = kp; a.v_pool = vp; a.block_table = e->block_table.as<int32_t>();
      a.max_pages = e->max_pages_per_seq; a.page_tokens = c.kv_page_tokens; a.pool_pages = e->n_pages;
      a.out = e->attn.as<__half>(); a.out_stride = H;
      a.batch = B; a.n_heads = c.llm_heads; a.seq_len = Ts; a.q_pos0 = L0; a.head_dim = 128;
      a.scale = 0.08838834764831845f;  // 1/sqrt(128)
      VCB_TRY(attn_llama_prefill(a, st));
    }
    {
      GemmEpilogue ep;
      ep.C = x; ep.ldc = H; ep.residual = x; ep.ldr = H;
      VCB_TRY(gemm_f16(e->attn.as<__half>(), H, l.o_w.as<__half>(), H, rows, H, H, ep, 0, st));
    }
    VCB_TRY(rmsnorm_f16(x, l.ln2.as<__half>(), xn, rows, H, c.llm_rms_eps, st));
    {
      GemmEpilogue ep;
      ep.C = e->mid.as<__half>(); ep.ldc = FF; ep.act = kActSwiGLU;
      VCB_TRY(gemm_f16(xn, H, l.gu_w.as<__half>(), H, rows, 2 * FF, H, ep, 0, st));
    }
    {
      GemmEpilogue ep;
      ep.C = x; ep.ldc = H; ep.residual = x; ep.ldr = H;
      VCB_TRY(gemm_f16(e->mid.as<__half>(), FF, l.down_w.as<__half>(), FF, rows, H, FF, ep, 0, st));
    }
  }
  // ---- forward() seam: logits of every position (final RMSNorm over all rows, then the tcgen05 GEMM against lm_head; a
  //      vocabulary or width the tile shapes do not divide goes through the weight-streaming GEMM 32 rows at a time)
  if (extra && extra->all_logits) {
    __half* out = static_cast<__half*>(extra->all_logits);
    if (V % 32 == 0 && H % 64 == 0) {
      VCB_TRY(rmsnorm_f16(x, e->final_norm.as<__half>(), xn, rows, H, c.llm_rms_eps, st));
      GemmEpilogue ep;
      ep.C = out; ep.ldc = V;
      VCB_TRY(gemm_f16(xn, H, e->lm_head.as<__half>(), H, rows, V, H, ep, 0, st));
    } else {
      for (long r0 = 0; r0 < rows; r0 += 32) {
        SkinnyParams s;
        s.norm_w = e->final_norm.as<__half>(); s.norm_eps = c.llm_rms_eps;
        s.X = x + r0 * H; s.ldx = H; s.W = e->lm_head.as<__half>(); s.Wp = e->lm_head_p.as<__half>(); s.Y = out + r0 * V; s.ldy = V;
        s.B = static_cast<int>(std::min<long>(32, rows - r0)); s.N = V; s.K = H;
        VCB_TRY(skinny_gemm(s, st));
      }
    }
  }
  // ---- logits for the LAST position only (the reference computes all T' rows and greedy reads [:, -1])
  __half* xd = e->xd.as<__half>();
  VCB_CUDA(cudaMemcpy2DAsync(xd, H * 2, x + static_cast<long>(Ts - 1) * H, static_cast<size_t>(Ts) * H * 2, H * 2, B,
                             cudaMemcpyDeviceToDevice, st));
  {
    SkinnyParams s;  // final RMSNorm fused into the lm_head GEMM
    s.norm_w = e->final_norm.as<__half>(); s.norm_eps = c.llm_rms_eps;
    s.X = xd; s.ldx = H; s.W = e->lm_head.as<__half>(); s.Wp = e->lm_head_p.as<__half>(); s.Y = e->logits.as<__half>(); s.ldy = V;
    s.B = B; s.N = V; s.K = H;
    VCB_TRY(skinny_gemm(s, st));
  }
  int32_t* counters = e->counters.as<int32_t>();
  const int32_t init[3] = {Tp, 0, Tp + 1};
  VCB_CUDA(cudaMemcpyAsync(counters, init, 12, cudaMemcpyHostToDevice, st));
  VCB_TRY(select_next(e, B, counters + 1, st));
  const int32_t after[3] = {Tp, 1, Tp + 1};
  VCB_CUDA(cudaMemcpyAsync(counters, after, 12, cudaMemcpyHostToDevice, st));
  if (last_logits_out) VCB_TRY(copy_any(last_logits_out, e->logits.p, static_cast<size_t>(B) * V * 2, st));
  if (next_ids_out) VCB_TRY(copy_any(next_ids_out, e->next_ids.p, B * 4, st));
  e->seq_plan.swap(plan);
  e->seq_plan_len = Tp;
  e->batch = B;
  e->cache_len = Tp;
  e->steps_done = 1;
  e->prefilled = true;
  return kOk;
}

// ------------------------------------------------------------------------------------------------------------
// decode: one greedy step = embed(next) -> L layers (skinny GEMMs + paged flash-decoding) -> lm_head -> argmax.
// All positions come from device counters, so one captured CUDA graph replays for every step.
// ------------------------------------------------------------------------------------------------------------
namespace {

int decode_step_launch(Engine* e, int max_ctx, cudaStream_t st) {
  const vcb_config& c = e->cfg;
  const int B = e->batch, H = c.llm_hidden, FF = c.llm_ffn, V = c.llm_vocab;
  int32_t* counters = e->counters.as<int32_t>();
  __half* x = e->xd.as<__half>();
  __half* xn = e->xnd.as<__half>();
  __half* qkv = e->qkvd.as<__half>();
  const size_t layer_pool = static_cast<size_t>(e->n_pages) * c.llm_heads * c.kv_page_tokens * 128;
  static int fuse_env = -1;  // VCB_FUSE_NORM=1 fuses RMSNorm into the skinny GEMMs (kept for A/B; default: separate kernels)
  if (fuse_env < 0) { const char* ev = getenv("VCB_FUSE_NORM"); fuse_env = ev ? atoi(ev) : 0; }  // measured slower on B200 (3.2 vs 5.2 TB/s): off
  const bool fuse_norm = fuse_env != 0;
  static int rope_env = -1;  // RoPE + KV append fused into the decode attention kernel (VCB_FUSE_ROPE=0: separate kernel)
  if (rope_env < 0) { const char* ev = getenv("VCB_FUSE_ROPE"); rope_env = ev ? atoi(ev) : 1; }
  const bool fuse_rope = rope_env != 0;
  // Chained RMSNorm (VCB_DECODE_NORM_CHAIN=1, off by default): no norm kernel inside the step.  The GEMM that produces the
  // residual stream (o-proj, down-proj; the embedding gather for layer 0) also emits fp16(x * w_next_norm) and per-CTA
  // partial sum(x^2); the GEMM that follows (qkv, gate|up, lm_head) scales its accumulators by rsqrt(mean(x^2) + eps).
  // Parity-green, but measured SLOWER than the 65 separate rmsnorm kernels (4.30 vs 4.05 ms per step): those overlap with
  // their neighbours through programmatic dependent launch, the extra work inside the GEMMs sits on the critical path.
  static int chain_env = -1;
  if (chain_env < 0) { const char* ev = getenv("VCB_DECODE_NORM_CHAIN"); chain_env = ev ? atoi(ev) : kNormChainDefault; }
  const bool chain = chain_env != 0 && !fuse_norm && H % 64 == 0 && FF % 64 == 0;  // the TMA kernel carries the chain
  // norm_parts0 = [rstd of the embedding rows (32) | rstd after o-proj (32) | rstd after down-proj (32) | 2 tickets]
  float* rstd0 = e->norm_parts0.as<float>();
  float* rstd_a = rstd0 + 32;
  float* rstd_b = rstd0 + 64;
  int32_t* tickets = reinterpret_cast<int32_t*>(rstd0 + 96);
  float* parts_a = e->norm_parts_a.as<float>();
  float* parts_b = e->norm_parts_b.as<float>();
  if (chain)
    VCB_TRY(embed_gather_norm(e->next_ids.as<int32_t>(), e->embed.as<__half>(), x, e->llm[0].ln1.as<__half>(), xn, rstd0,
                              c.llm_rms_eps, B, H, st));
  else
    VCB_TRY(embed_gather(e->next_ids.as<int32_t>(), e->embed.as<__half>(), x, B, H, st));
  for (int li = 0; li < c.llm_layers; ++li) {
    const LlmLayerW& l = e->llm[li];
    __half* kp = e->k_pool.as<__half>() + li * layer_pool;
    __half* vp = e->v_pool.as<__half>() + li * layer_pool;
    if (!fuse_norm && !chain) VCB_TRY(rmsnorm_f16(x, l.ln1.as<__half>(), xn, B, H, c.llm_rms_eps, st));
    {
      SkinnyParams s;
      if (fuse_norm) { s.norm_w = l.ln1.as<__half>(); s.norm_eps = c.llm_rms_eps; }
      if (chain) s.cn_rstd = li == 0 ? rstd0 : rstd_b;
      s.X = fuse_norm ? x : xn; s.ldx = H; s.W = l.qkv_w.as<__half>(); s.Wp = l.qkv_p.as<__half>(); s.Y = qkv; s.ldy = 3 * H; s.B = B; s.N = 3 * H; s.K = H;
      VCB_TRY(skinny_gemm(s, st));
    }
    if (!fuse_rope)
      VCB_TRY(rope_kv_write(qkv, B, 1, H, c.llm_heads, e->rope.as<float2>(), 0, counters, kp, vp,
                            e->block_table.as<int32_t>(), e->max_pages_per_seq, c.kv_page_tokens, st));
    {
      AttnDecodeParams a;
      if (fuse_rope) { a.cos_sin = e->rope.as<float2>(); a.k_pool_w = kp; a.v_pool_w = vp; }
      a.q = qkv; a.q_stride = 3 * H; a.k_pool = kp; a.v_pool = vp;
      a.block_table = e->block_table.as<int32_t>(); a.max_pages = e->max_pages_per_seq; a.page_tokens = c.kv_page_tokens;
      a.ctx_len = counters + 2;  // cache_len + 1 (the current token attends to itself)
      a.max_ctx = max_ctx;
      a.part_o = e->part_o.as<float>(); a.part_ml = e->part_ml.as<float>(); a.sync_counters = e->attn_sync.as<int32_t>();
      a.out = e->attnd.as<__half>(); a.batch = B; a.n_heads = c.llm_heads; a.scale = 0.08838834764831845f;
      VCB_TRY(attn_llama_decode(a, st));
    }
    {
      SkinnyParams s;
      s.X = e->attnd.as<__half>(); s.ldx = H; s.W = l.o_w.as<__half>(); s.Wp = l.o_p.as<__half>(); s.Y = x; s.ldy = H; s.residual = x; s.ldr = H;
      s.B = B; s.N = H; s.K = H;
      if (chain) { s.pn_w = l.ln2.as<__half>(); s.pn_out = xn; s.pn_ld = H; s.pn_parts = parts_a; s.pn_rstd = rstd_a; s.pn_ticket = tickets; s.pn_eps = c.llm_rms_eps; }
      VCB_TRY(skinny_gemm(s, st));
    }
    if (!fuse_norm && !chain) VCB_TRY(rmsnorm_f16(x, l.ln2.as<__half>(), xn, B, H, c.llm_rms_eps, st));
    {
      SkinnyParams s;
      if (fuse_norm) { s.norm_w = l.ln2.as<__half>(); s.norm_eps = c.llm_rms_eps; }
      if (chain) s.cn_rstd = rstd_a;
      s.X = fuse_norm ? x : xn; s.ldx = H; s.W = l.gu_w.as<__half>(); s.Wp = l.gu_p.as<__half>(); s.Y = e->midd.as<__half>(); s.ldy = FF; s.B = B; s.N = 2 * FF;
      s.K = H; s.swiglu = 1;
      VCB_TRY(skinny_gemm(s, st));
    }
    {
      SkinnyParams s;
      s.X = e->midd.as<__half>(); s.ldx = FF; s.W = l.down_w.as<__half>(); s.Wp = l.down_p.as<__half>(); s.Y = x; s.ldy = H; s.residual = x; s.ldr = H;
      s.B = B; s.N = H; s.K = FF;
      if (chain) {
        s.pn_w = (li + 1 < c.llm_layers ? e->llm[li + 1].ln1 : e->final_norm).as<__half>();
        s.pn_out = xn; s.pn_ld = H; s.pn_parts = parts_b; s.pn_rstd = rstd_b; s.pn_ticket = tickets + 1; s.pn_eps = c.llm_rms_eps;
      }
      VCB_TRY(skinny_gemm(s, st));
    }
  }
  if (!fuse_norm && !chain) VCB_TRY(rmsnorm_f16(x, e->final_norm.as<__half>(), xn, B, H, c.llm_rms_eps, st));
  {
    SkinnyParams s;
    if (fuse_norm) { s.norm_w = e->final_norm.as<__half>(); s.norm_eps = c.llm_rms_eps; }
    if (chain) s.cn_rstd = rstd_b;
    s.X = fuse_norm ? x : xn; s.ldx = H; s.W = e->lm_head.as<__half>(); s.Wp = e->lm_head_p.as<__half>(); s.Y = e->logits.as<__half>(); s.ldy = V; s.B = B; s.N = V; s.K = H;
    VCB_TRY(skinny_gemm(s, st));
  }
  VCB_TRY(select_next(e, B, counters + 1, st));
  VCB_TRY(advance_counters(counters, st));
  return kOk;
}

}  // namespace

namespace {

// Makes e->graph a replayable capture of one decode step for (batch, splits, greedy|sampling).  Building it runs ONE
// eager step first (sets per-kernel attributes outside capture and validates the launch sequence): *ran_eager = 1 then.
int ensure_graph(Engine* e, int max_ctx, cudaStream_t st, int* ran_eager) {
  const int B = e->batch, splits = attn_decode_splits(max_ctx), mode = e->sampling ? 1 : 0;
  *ran_eager = 0;
  if (e->graph != nullptr && e->graph_batch == B && e->graph_splits == splits && e->graph_sampling == mode &&
      e->graph_trace_epoch == trace_epoch())
    return kOk;
  VCB_TRY(decode_step_launch(e, max_ctx, st));
  *ran_eager = 1;
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  cudaGraph_t g = nullptr;
  const uint64_t c0 = launch_count();
  VCB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = decode_step_launch(e, max_ctx, st);
  cudaError_t ce = cudaStreamEndCapture(st, &g);
  e->graph_kernels = static_cast<int>(launch_count() - c0);
  launch_count_add(-static_cast<int64_t>(e->graph_kernels));  // captured, not executed
  if (rc != kOk) { if (g) cudaGraphDestroy(g); return rc; }
  VCB_CUDA(ce);
  ce = cudaGraphInstantiate(&e->graph, g, 0);
  cudaGraphDestroy(g);
  VCB_CUDA(ce);
  e->graph_batch = B;
  e->graph_splits = splits;
  e->graph_sampling = mode;
  e->graph_trace_epoch = trace_epoch();
  return kOk;
}

}  // namespace

int engine_decode(Engine* e, int n_steps, void* out_ids, void* logits_trace, cudaStream_t st) {
  const vcb_config& c = e->cfg;
  VCB_CHECK(e->prefilled, kErrState, "vcb_decode called before vcb_prefill");
  VCB_CHECK(n_steps >= 0, kErrInvalid, "vcb_decode: negative step count");
  VCB_CUDA(cudaSetDevice(e->device));
  const int B = e->batch, V = c.llm_vocab;
  const int first_token = e->steps_done;  // index (since prefill) of the first token this call produces
  VCB_CHECK(e->cache_len + n_steps <= c.max_seq_len, kErrNoMem, "vcb_decode: would exceed max_seq_len");
  VCB_TRY(ensure_pages(e, B, e->cache_len + n_steps, st));
  const int max_ctx = e->cache_len + n_steps;
  int done = 0;
  const bool use_graph = logits_trace == nullptr && n_steps >= 3 && !prof_enabled();
  if (use_graph) {
    VCB_TRY(ensure_graph(e, max_ctx, st, &done));
    for (; done < n_steps; ++done) {
      VCB_CUDA(cudaGraphLaunch(e->graph, st));
      launch_count_add(e->graph_kernels);
    }
  } else {
    for (; done < n_steps; ++done) {
      VCB_TRY(decode_step_launch(e, max_ctx, st));
      if (logits_trace)
        VCB_TRY(copy_any(static_cast<__half*>(logits_trace) + static_cast<long>(done) * B * V, e->logits.p,
                         static_cast<size_t>(B) * V * 2, st));
    }
  }
  e->cache_len += n_steps;
  e->steps_done += n_steps;
  if (out_ids && n_steps > 0) {
    VCB_CUDA(cudaMemcpy2DAsync(out_ids, static_cast<size_t>(n_steps) * 4, e->trace.as<int32_t>() + first_token,
                               static_cast<size_t>(c.max_seq_len) * 4, static_cast<size_t>(n_steps) * 4, B,
                               cudaMemcpyDefault, st));
  }
  return kOk;
}

// ------------------------------------------------------------------------------------------------------------
// f1: sampling parameters and the streamed decode loop (serve/cli.py:117-132, serve/chat.py:131-159)
// ------------------------------------------------------------------------------------------------------------
int engine_set_sampling(Engine* e, int do_sample, float temperature, int top_k, float top_p, uint64_t seed,
                        cudaStream_t st) {
  VCB_CUDA(cudaSetDevice(e->device));
  if (!do_sample) {
    e->sampling = false;
    return kOk;
  }
  // the same argument checks HF's warpers make (logits_process.py: Temperature / TopK / TopP __init__)
  VCB_CHECK(temperature > 0.f && std::isfinite(temperature), kErrInvalid,
            "`temperature` has to be a strictly positive float");
  VCB_CHECK(top_k >= 0, kErrInvalid, "`top_k` has to be a non-negative integer (0 disables the filter)");
  VCB_CHECK(top_p >= 0.f && top_p <= 1.f, kErrInvalid, "`top_p` has to be a float >= 0 and <= 1");
  VCB_CHECK(e->cfg.llm_vocab <= 32768, kErrInvalid, "sampling: vocabulary larger than 32768 is not supported");
  unsigned char blk[32] = {0};
  pack_sample_params(temperature, top_k, top_p, seed, blk);
  // ordered on the stream: a step already queued keeps the parameters it was launched with
  VCB_CUDA(cudaMemcpyAsync(e->sample_params.p, blk, 32, cudaMemcpyHostToDevice, st));
  VCB_CUDA(cudaStreamSynchronize(st));  // blk is a stack buffer
  e->sampling = true;
  return kOk;
}

int engine_decode_stream(Engine* e, int max_steps, int lookahead, engine_token_cb cb, void* user, int* n_accepted,
                         cudaStream_t st) {
  const vcb_config& c = e->cfg;
  VCB_CHECK(e->prefilled, kErrState, "vcb_decode_stream called before vcb_prefill");
  VCB_CHECK(max_steps >= 0 && cb != nullptr, kErrInvalid, "vcb_decode_stream: bad arguments");
  VCB_CUDA(cudaSetDevice(e->device));
  if (n_accepted) *n_accepted = 0;
  if (max_steps == 0) return kOk;
  const int B = e->batch;
  const int R = Engine::kStreamRing;
  lookahead = std::max(1, std::min(lookahead, R));
  VCB_CHECK(e->cache_len + max_steps <= c.max_seq_len, kErrNoMem, "vcb_decode_stream: would exceed max_seq_len");
  VCB_TRY(ensure_pages(e, B, e->cache_len + max_steps, st));
  const int max_ctx = e->cache_len + max_steps;
  const int first_token = e->steps_done;
  const bool use_graph = !prof_enabled();
  int issued = 0, consumed = 0, accepted = 0;
  bool stop = false;
  int rc = kOk;
  auto issue_one = [&]() -> int {
    int ran = 0;
    if (use_graph) VCB_TRY(ensure_graph(e, max_ctx, st, &ran));
    if (!ran) {
      if (use_graph) {
        VCB_CUDA(cudaGraphLaunch(e->graph, st));
        launch_count_add(e->graph_kernels);
      } else {
        VCB_TRY(decode_step_launch(e, max_ctx, st));
      }
    }
    const int slot = issued % R;
    VCB_CUDA(cudaMemcpy2DAsync(e->stream_host + static_cast<size_t>(slot) * B, 4,
                               e->trace.as<int32_t>() + first_token + issued, static_cast<size_t>(c.max_seq_len) * 4, 4,
                               B, cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaEventRecord(e->stream_ev[slot], st));
    ++issued;
    return kOk;
  };
  while (!stop && consumed < max_steps) {
    while (issued < max_steps && issued - consumed < lookahead) {
      rc = issue_one();
      if (rc != kOk) break;
    }
    if (rc != kOk) break;
    const int slot = consumed % R;
    if (cudaEventSynchronize(e->stream_ev[slot]) != cudaSuccess) {
      set_thread_error("cudaEventSynchronize failed in vcb_decode_stream");
      rc = kErrCuda;
      break;
    }
    stop = cb(user, consumed, e->stream_host + static_cast<size_t>(slot) * B, B) != 0;
    ++consumed;
    accepted = consumed;
  }
  // steps queued past the accepted one are discarded: wait for them, then roll the device state back
  cudaError_t se = cudaStreamSynchronize(st);
  if (rc == kOk && se != cudaSuccess) {
    set_thread_error(std::string("vcb_decode_stream: ") + cudaGetErrorString(se));
    rc = kErrCuda;
  }
  if (rc != kOk) {
    e->prefilled = false;  // device counters are ahead of the host mirror: force a new prefill
    return rc;
  }
  e->cache_len += accepted;
  e->steps_done += accepted;
  if (issued > accepted) {
    const int32_t ctr[3] = {e->cache_len, e->steps_done, e->cache_len + 1};
    VCB_CUDA(cudaMemcpyAsync(e->counters.p, ctr, 12, cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpy2DAsync(e->next_ids.p, 4, e->trace.as<int32_t>() + e->steps_done - 1,
                               static_cast<size_t>(c.max_seq_len) * 4, 4, B, cudaMemcpyDeviceToDevice, st));
    VCB_CUDA(cudaStreamSynchronize(st));  // ctr is a stack buffer
  }
  if (n_accepted) *n_accepted = accepted;
  return kOk;
}

// The decode step's weight-streaming GEMMs alone (per layer qkv, o, gate|up, down; then lm_head) on the bound weights,
// back to back with PDL exactly as inside the decode graph, `rounds` times between two CUDA events.  bench.py divides the
// algorithmic weight bytes by this for the roofline of the dominant kernel without per-launch event overhead.
int engine_bench_weight_stream(Engine* e, int B, int rounds, float* ms_total, int* launches, cudaStream_t st) {
  const vcb_config& c = e->cfg;
  VCB_CHECK(B >= 1 && B <= c.max_batch && rounds >= 1 && ms_total != nullptr, kErrInvalid, "bench_weight_stream: bad arguments");
  VCB_CUDA(cudaSetDevice(e->device));
  VCB_TRY(engine_check_bound(e, false, true));
  const int H = c.llm_hidden, FF = c.llm_ffn, V = c.llm_vocab;
  cudaEvent_t a = nullptr, b = nullptr;
  VCB_CUDA(cudaEventCreate(&a));
  VCB_CUDA(cudaEventCreate(&b));
  int n = 0, rc = kOk;
  VCB_TRY(engine_ensure_packed(e, st));
  auto run = [&](const __half* X, long ldx, const DevBuf& W, const DevBuf& Wp, __half* Y, long ldy, int N, int K, int swiglu) {
    if (rc != kOk) return;
    SkinnyParams s;
    s.X = X; s.ldx = ldx; s.W = W.as<__half>(); s.Wp = Wp.as<__half>(); s.Y = Y; s.ldy = ldy; s.B = B; s.N = N; s.K = K; s.swiglu = swiglu;
    rc = skinny_gemm(s, st);
    ++n;
  };
  for (int r = -1; r < rounds && rc == kOk; ++r) {  // round -1 = warm-up
    if (r == 0) { n = 0; cudaEventRecord(a, st); }
    for (int li = 0; li < c.llm_layers; ++li) {
      const LlmLayerW& l = e->llm[li];
      run(e->xnd.as<__half>(), H, l.qkv_w, l.qkv_p, e->qkvd.as<__half>(), 3 * H, 3 * H, H, 0);
      run(e->attnd.as<__half>(), H, l.o_w, l.o_p, e->xd.as<__half>(), H, H, H, 0);
      run(e->xnd.as<__half>(), H, l.gu_w, l.gu_p, e->midd.as<__half>(), FF, 2 * FF, H, 1);
      run(e->midd.as<__half>(), FF, l.down_w, l.down_p, e->xd.as<__half>(), H, H, FF, 0);
    }
    run(e->xnd.as<__half>(), H, e->lm_head, e->lm_head_p, e->logits.as<__half>(), V, V, H, 0);
  }
  cudaEventRecord(b, st);
  cudaError_t se = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (se == cudaSuccess) cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  e->prefilled = false;  // the decode workspaces were overwritten
  VCB_TRY(rc);
  VCB_CUDA(se);
  *ms_total = ms;
  if (launches) *launches = n;
  return kOk;
}

int engine_reset(Engine* e) {
  e->prefilled = false;
  e->seq_plan.clear();
  e->seq_plan_len = 0;
  e->feat_key = 0;
  e->batch = 0;
  e->cache_len = 0;
  e->steps_done = 0;
  e->pages_per_seq_alloc = 0;
  return kOk;
}

}  // namespace vcb
