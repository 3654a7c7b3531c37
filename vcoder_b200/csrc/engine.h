// Engine: owns packed weights, workspaces, the paged KV cache and the decode CUDA graph for one model on one GPU.
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "common.h"
#include "kernels.h"
#include "vcoder_b200.h"

namespace vcb {

// Host-side splice planner (pure C++, no CUDA): restates prepare_inputs_labels_for_multimodal
// (vcoder_ds_llava_arch.py:175-312 / vcoder_llava_arch.py:182-294).  plan[b*T_out + t]:
//   >= 0 token id (row of embed_tokens), -1 zero pad, <= -2 feature row -(v+2) of the [kinds*B*tokens] feature buffer.
int splice_plan_host(int variant, const int64_t* ids, int B, int T, int n_patches, int has_seg,
                     const uint8_t* depth_nonzero /* null = depths is None */, std::vector<int32_t>* plan,
                     int* out_len, bool* depth_flag_needed);

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct VitLayerW {
  DevBuf qkv_w, qkv_b, out_w, out_b, ln1_w, ln1_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
};
struct ProjW {
  DevBuf w1, b1, w2, b2;
};
struct LlmLayerW {
  DevBuf qkv_w, o_w, gu_w, down_w, ln1, ln2;
  DevBuf qkv_p, o_p, gu_p, down_p;  // tile-major copies for the decode weight stream (pack_skinny_weights), if enabled
};

struct Engine {
  vcb_config cfg{};
  int device = 0;
  // derived
  int P = 0;        // patches per image
  int S = 0;        // tokens per image (P + 1)
  int Kpe = 0;      // padded im2col K
  int vit_used = 0; // encoder layers actually evaluated
  int max_pages_per_seq = 0, n_pages = 0;
  int vit_max_images = 0;
  long max_rows = 0;  // LLM prefill rows (max_batch * max_seq_len)

  // weights
  DevBuf patch_w, cls, pos, preln_w, preln_b;
  std::vector<VitLayerW> vit;
  ProjW mm, seg;
  DevBuf embed, final_norm, lm_head, lm_head_p;
  bool use_packed = false;    // decode GEMMs stream the tile-major weight copies (VCB_SKINNY_PACKED)
  bool packed_valid = false;  // the copies reflect the bound weights
  std::vector<LlmLayerW> llm;
  std::set<std::string> bound;

  // workspaces
  DevBuf pix, im2col, vh, vn, vqkv, vattn, vmid;  // vision
  DevBuf pmid, feats;                             // projector
  DevBuf x, xn, qkv, attn, mid;                   // LLM prefill
  DevBuf xd, xnd, qkvd, attnd, midd;              // LLM decode (B rows)
  DevBuf logits, next_ids, trace, part_o, part_ml, attn_sync, plan_dev, depth_sums;
  DevBuf norm_parts0, norm_parts_a, norm_parts_b;  // chained RMSNorm of the decode step: partial sum(x^2), [producer CTAs][32]
  DevBuf k_pool, v_pool, block_table, rope, counters;  // counters: [0] cache_len, [1] step
  int32_t* plan_host = nullptr;                        // pinned
  int32_t* bt_host = nullptr;                          // pinned
  float* depth_host = nullptr;                         // pinned
  std::vector<void*> allocs;

  // state
  int batch = 0;
  int cache_len = 0;  // host mirror
  int steps_done = 0; // tokens produced since the last prefill
  int pages_per_seq_alloc = 0;
  bool prefilled = false;
  cudaGraphExec_t graph = nullptr;
  cudaStream_t own_stream = nullptr;   // used when the caller hands us the legacy default stream
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int graph_batch = 0, graph_splits = 0, graph_kernels = 0, graph_sampling = 0, graph_trace_epoch = 0;
  // f1: sampling + per-token streaming
  DevBuf sample_params;      // SampleDeviceParams (sample.cu), read by the sampling kernel inside the captured graph
  bool sampling = false;     // do_sample: the next-token kernel is sample_rows instead of argmax_rows
  static constexpr int kStreamRing = 16;
  cudaEvent_t stream_ev[kStreamRing] = {};
  int32_t* stream_host = nullptr;  // pinned [kStreamRing][max_batch]
  // f3: multi-turn prefix cache (serve/chat.py:41-159 re-sends the whole conversation and all images every turn)
  std::vector<int32_t> seq_plan;  // host mirror of the last prefill's spliced plan [batch][seq_plan_len]
  int seq_plan_len = 0;
  uint64_t feat_key = 0;          // fingerprint of the pixels whose projected features sit in `feats` (0 = none)
  int feat_batch = 0, feat_kinds = 0;
  DevBuf fp_dev;                  // 8-byte accumulator of fingerprint_kernel
  uint64_t* fp_host = nullptr;    // pinned

  ~Engine();
};

int engine_create(const vcb_config* cfg, int device, Engine** out);
int engine_bind(Engine* e, const vcb_tensor* t, int n);
int engine_check_bound(Engine* e, bool need_vision, bool need_llm);
int engine_ensure_packed(Engine* e, cudaStream_t st);
// images: device pointer to n images [n,3,S,S] fp16 already resident; result rows in e->vh ([n*S, D])
int engine_tower(Engine* e, const __half* pixels_dev, int n_images, cudaStream_t st);
int engine_project(Engine* e, int row_begin, int row_end, const ProjW& w, cudaStream_t st);
// reuse = false: every call runs the towers and the whole prompt (vcb_prefill).  reuse = true (vcb_prefill_cached): the
// towers are skipped when the pixels' fingerprint equals the previous call's, and only the part of the spliced prompt
// behind the longest prefix it shares with the sequence already in the KV cache runs through the decoder.
// `extra` (optional) serves the forward() / prepare_inputs_labels_for_multimodal seams: embeds_only stops after the
// splice (no decoder, no KV; the handle is left un-prefilled); all_logits receives final-norm + lm_head of EVERY spliced
// position, [B, T', V] fp16, and must hold all_logits_halves >= B * T' * V halves (device memory).
struct PrefillExtra {
  bool embeds_only = false;
  void* all_logits = nullptr;
  long all_logits_halves = 0;
};
int engine_prefill(Engine* e, const int64_t* ids_host, int B, int T, const void* images, const void* segs,
                   const void* depths, bool reuse, void* next_ids_out, void* last_logits_out, void* embeds_out,
                   int* spliced_len, int* reused_len, int* towers_skipped, cudaStream_t st,
                   const PrefillExtra* extra = nullptr);
// Host-only part of the prefix cache: longest prefix (<= limit) shared by every row of `cached` [B][cached_len] and
// `plan` [B][plan_len]; visual rows (entries <= -2) only count when `features_same`.
int prefix_match_host(const int32_t* cached, int cached_len, const int32_t* plan, int plan_len, int B, int features_same,
                      int limit);
int engine_decode(Engine* e, int n_steps, void* out_ids, void* logits_trace, cudaStream_t st);
int engine_reset(Engine* e);
// do_sample / temperature / top_k / top_p / seed of the following prefill + decode calls (HF GenerationMixin.sample)
int engine_set_sampling(Engine* e, int do_sample, float temperature, int top_k, float top_p, uint64_t seed,
                        cudaStream_t st);
// Decode with the host in the loop: up to `lookahead` steps stay queued on the GPU while cb() sees the tokens of step i
// (TextStreamer.put / KeywordsStoppingCriteria / eos checks of serve/cli.py:117-132).  cb != 0 stops: the steps already
// queued are discarded and the device counters are rolled back, so a later decode continues after the accepted token.
typedef int (*engine_token_cb)(void* user, int step, const int32_t* ids, int B);
int engine_decode_stream(Engine* e, int max_steps, int lookahead, engine_token_cb cb, void* user, int* n_accepted,
                         cudaStream_t st);
int engine_bench_weight_stream(Engine* e, int B, int rounds, float* ms_total, int* launches, cudaStream_t st);
int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t st);

// Work always runs on a capturable (non-legacy) stream.  When the caller passes NULL / the legacy default stream,
// the engine's own stream is ordered after the caller's prior work on entry and the caller's stream waits for the
// engine's work on exit, so the usual stream semantics are preserved.
struct StreamGuard {
  StreamGuard(Engine* e, cudaStream_t caller);
  ~StreamGuard();
  cudaStream_t st;
  Engine* e_;
  cudaStream_t caller_;
  bool swapped_;
};

}  // namespace vcb
