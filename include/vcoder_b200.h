/* libvcoder_b200 — C ABI of the B200-native VCoder inference hot path.
 *
 * The reference (SHI-Labs/VCoder) has no FFI of its own: its hot path is reached through Python
 * (`model.generate(input_ids, images=, segs=, depths=)`, vcoder_llava/serve/cli.py:121-132,
 * vcoder_llava/eval/model_depth_loader.py:146-168) and bottoms out in HF transformers -> ATen/cuBLAS.
 * This header is the boundary a maintainer binds with ctypes (INTEGRATION.md shows the stub); every entry
 * point cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VCB_ERR_* code; vcb_last_error() returns a
 *     thread-local message for the last failing call on this thread.
 *   - plain pointers and sizes only.  "dev" pointers are CUDA device pointers (e.g. torch .data_ptr());
 *     "host" pointers are ordinary host memory; "any" accepts either (resolved with cudaPointerGetAttributes).
 *   - fp16 data is IEEE binary16 stored as uint16_t in this header (no CUDA types leak into the ABI).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with VCB_ERR_NO_DEVICE.
 *   - a handle is single-stream and not re-entrant; callers serialise (the Python shim holds a lock).
 */
#ifndef VCODER_B200_H_
#define VCODER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCB_OK 0
#define VCB_ERR_INVALID (-1)
#define VCB_ERR_CUDA (-2)
#define VCB_ERR_STATE (-3)
#define VCB_ERR_NOMEM (-4)
#define VCB_ERR_RAGGED (-5)        /* spliced lengths differ inside a batch: the reference raises NameError
                                      (vcoder_ds_llava_arch.py:292-300 uses _new_labels, unset when labels=None) */
#define VCB_ERR_SPECIAL_INDEX (-6) /* a negative special id would reach the embedding table: reference IndexError
                                      (vcoder_llava_arch.py:187 dummy branch) */
#define VCB_ERR_NO_DEVICE (-7)

/* Special token ids, vcoder_llava/constants.py:3-12 */
#define VCB_IMAGE_TOKEN_INDEX (-200)
#define VCB_SEG_TOKEN_INDEX (-300)
#define VCB_DEPTH_TOKEN_INDEX (-400)

#define VCB_VARIANT_VCODER 0    /* VCoderLlavaLlamaForCausalLM   (vcoder_llava_llama.py:41)    */
#define VCB_VARIANT_VCODER_DS 1 /* VCoderDSLlavaLlamaForCausalLM (vcoder_ds_llava_llama.py:41) */
#define VCB_VARIANT_LLAVA 2     /* LlavaLlamaForCausalLM: RGB only (llava_llama.py:41, llava_arch.py:85-199) */
#define VCB_VARIANT_VCODER_IT 3 /* VCoderITLlavaLlamaForCausalLM: RGB + optional seg (vcoder_it_llava_llama.py:41,
                                   vcoder_it_llava_arch.py:128-277) */

typedef uint16_t vcb_half;
typedef struct vcb_engine* vcb_handle;

/* Model geometry: the HF config.json keys the reference reads (vcoder_ds_llava_arch.py:34-49,
 * CLIPVisionConfig, LlamaConfig) plus capacity limits for the paged KV cache. */
typedef struct vcb_config {
  int32_t struct_size; /* sizeof(vcb_config), for ABI evolution */
  int32_t variant;     /* VCB_VARIANT_* */
  /* CLIP vision tower (clip_encoder.py:7-77) */
  int32_t vit_image_size; /* 336 */
  int32_t vit_patch;      /* 14 */
  int32_t vit_hidden;     /* 1024 */
  int32_t vit_heads;      /* 16 (head_dim must be 64) */
  int32_t vit_ffn;        /* 4096 */
  int32_t vit_layers;     /* 24: layers present in the checkpoint */
  int32_t vit_select_layer; /* mm_vision_select_layer, -2: hidden_states[-2] => vit_layers-1 layers are run */
  float vit_ln_eps;       /* 1e-5 */
  /* LLaMA / Vicuna decoder */
  int32_t llm_hidden; /* 4096 / 5120 (head_dim must be 128) */
  int32_t llm_layers; /* 32 / 40 */
  int32_t llm_heads;  /* 32 / 40 */
  int32_t llm_ffn;    /* 11008 / 13824 */
  int32_t llm_vocab;  /* 32000 */
  float llm_rms_eps;  /* 1e-5 */
  float llm_rope_theta; /* 10000 */
  /* capacity */
  int32_t max_batch;      /* sequences per prefill */
  int32_t max_seq_len;    /* spliced prompt + new tokens, per sequence */
  int32_t kv_page_tokens; /* tokens per KV page (64) */
} vcb_config;

/* A named weight, exactly as it appears in the reference model's state_dict(), e.g.
 * "model.layers.0.self_attn.q_proj.weight", "model.mm_projector.0.weight",
 * "model.vision_tower.vision_tower.vision_model.encoder.layers.3.mlp.fc1.bias", "lm_head.weight". */
typedef struct vcb_tensor {
  const char* name;
  const void* data; /* any: host or device, contiguous fp16 */
  int32_t ndim;
  int64_t shape[4];
} vcb_tensor;

/* ---- library ------------------------------------------------------------------------------- */
const char* vcb_version(void);
const char* vcb_last_error(void);
/* number of visible CUDA devices (0 on a CPU-only machine; never an error) */
int vcb_device_count(void);

/* ---- operator-level entry points (device pointers; used by the parity tests and by ncu runs) */

/* C[M,N] = act(A[M,K] · W[N,K]^T + bias) (+ residual).  Replaces nn.Linear / F.linear on the path.
 * act: 0 none, 1 quick_gelu (CLIP), 2 gelu erf (mm_projector), 3 SwiGLU with 8/8-interleaved gate|up rows
 * (output then has N/2 columns).  block_n: 0 auto, 128 / 256 (tile width of the 1-CTA kernel), 512 (the
 * CTA-pair kernel: 256 x 256 tile on two SMs with tcgen05 cta_group::2; needs N % 256 == 0). */
int vcb_op_gemm_f16(const vcb_half* A_dev, int32_t lda, const vcb_half* W_dev, int32_t ldw, int32_t M, int32_t N,
                    int32_t K, const vcb_half* bias_dev, const vcb_half* residual_dev, int32_t ldr, int32_t act,
                    vcb_half* C_dev, int32_t ldc, int32_t block_n, void* stream);

/* Y[B,N] = X[B,K] · W[N,K]^T (+ residual), B <= 32: the decode / lm_head weight-streaming kernel.
 * swiglu != 0: W rows 8/8 interleaved gate|up, Y has N/2 columns. */
int vcb_op_skinny_gemm_f16(const vcb_half* X_dev, int32_t ldx, const vcb_half* W_dev, int32_t B, int32_t N, int32_t K,
                           const vcb_half* residual_dev, int32_t ldr, int32_t swiglu, vcb_half* Y_dev, int32_t ldy,
                           void* stream);

/* The same GEMM through the round-1 kernel (one CTA per 16-row tile of W; needs K % 64 == 0).  vcb_op_skinny_gemm_f16 runs the
 * balanced multi-tile kernel (every CTA streams ceil(tiles / resident CTAs) consecutive tiles): identical arithmetic in
 * identical order, so the two must agree bit for bit — that is what this entry exists to test. */
int vcb_op_skinny_gemm_onetile_f16(const vcb_half* X_dev, int32_t ldx, const vcb_half* W_dev, int32_t B, int32_t N, int32_t K,
                                   const vcb_half* residual_dev, int32_t ldr, int32_t swiglu, vcb_half* Y_dev, int32_t ldy,
                                   void* stream);

/* The decode weight stream's own HBM layout.  vcb_op_pack_skinny_weights copies a row-major W[N,K] (N % 16 == 0,
 * K % 64 == 0) into the tile-major form [N/16][K/64][16 rows][128 B] (16-byte chunks pre-swizzled with row & 7), i.e. the
 * shared-memory image of every 16-row x 64-k block, so that a CTA's 16 x K slab is one contiguous run and a pipeline stage
 * is one linear 8 KiB bulk copy.  vcb_op_skinny_gemm_packed_f16 is vcb_op_skinny_gemm_f16 reading that copy; same
 * arithmetic in the same order, hence bit-identical results.  The engine keeps such a copy of every matrix the decode step
 * streams when VCB_SKINNY_PACKED=1 (replaces the same nn.Linear calls as vcb_op_skinny_gemm_f16). */
int vcb_op_pack_skinny_weights(const vcb_half* W_dev, int32_t N, int32_t K, vcb_half* Wp_dev, void* stream);
int vcb_op_skinny_gemm_packed_f16(const vcb_half* X_dev, int32_t ldx, const vcb_half* Wp_dev, int32_t B, int32_t N, int32_t K,
                                  const vcb_half* residual_dev, int32_t ldr, int32_t swiglu, vcb_half* Y_dev, int32_t ldy,
                                  void* stream);

/* Same kernel with LlamaRMSNorm fused onto X (decode path): Y = rsqrt(mean(x^2)+eps) * ((x o norm_w) · W^T) (+ residual).
 * Replaces `self.input_layernorm` / `post_attention_layernorm` / `model.norm` followed by the Linear
 * (transformers/models/llama/modeling_llama.py:53-70, 171-184). */
int vcb_op_skinny_rmsnorm_gemm_f16(const vcb_half* X_dev, int32_t ldx, const vcb_half* norm_w_dev, float eps,
                                   const vcb_half* W_dev, int32_t B, int32_t N, int32_t K, const vcb_half* residual_dev,
                                   int32_t ldr, int32_t swiglu, vcb_half* Y_dev, int32_t ldy, void* stream);

/* nn.LayerNorm over the last dim (CLIP pre_layrnorm / layer_norm1/2), rows x D fp16. */
int vcb_op_layernorm_f16(const vcb_half* x_dev, const vcb_half* w_dev, const vcb_half* b_dev, vcb_half* y_dev,
                         int64_t rows, int32_t D, float eps, void* stream);
/* LlamaRMSNorm (transformers/models/llama/modeling_llama.py:53-70), rows x H fp16. */
int vcb_op_rmsnorm_f16(const vcb_half* x_dev, const vcb_half* w_dev, vcb_half* y_dev, int64_t rows, int32_t H,
                       float eps, void* stream);

/* CLIP self-attention on a packed qkv buffer [B*S, 3*heads*64] -> out [B*S, heads*64];
 * softmax(q k^T * scale) v, no mask (transformers/models/clip/modeling_clip.py:261-279). */
int vcb_op_attention_vit_f16(const vcb_half* qkv_dev, int32_t B, int32_t S, int32_t heads, float scale,
                             vcb_half* out_dev, void* stream);

/* LLaMA causal self-attention at prefill over the paged KV cache (transformers/models/llama/modeling_llama.py:199-222
 * eager attention with the causal mask, reached from vcoder_ds_llava_llama.py:81): query row i of sequence b sits at position
 * q_pos0 + i and attends to cached keys [0, q_pos0 + i].  q_dev [B*Tq, q_stride] fp16 (head h at columns h*128..), already
 * rotated; k_pool_dev / v_pool_dev [n_pages][heads][64 tokens][128] fp16 already hold positions [0, q_pos0 + Tq);
 * block_table_dev int32 [B][max_pages] = page id of tokens [64 p, 64 p + 64) of sequence b; out_dev [B*Tq, out_stride]. */
int vcb_op_attention_prefill_f16(const vcb_half* q_dev, int64_t q_stride, const vcb_half* k_pool_dev,
                                 const vcb_half* v_pool_dev, const int32_t* block_table_dev, int32_t max_pages, int32_t B,
                                 int32_t heads, int32_t Tq, int32_t q_pos0, float scale, vcb_half* out_dev,
                                 int64_t out_stride, void* stream);

/* One decode step's attention over the paged KV cache: every cached position [0, ctx_len) is attended (the reference
 * substitutes an all-ones mask, vcoder_ds_llava_arch.py:130-133).  qkv_dev [B, 3*heads*128] fp16 is the QKV row of the current
 * token.  fuse_rope = 0: q is already rotated and the pools already hold position ctx_len - 1.  fuse_rope != 0 (what the
 * engine's decode step runs): q and k are un-rotated; the kernel rotates them for position ctx_len - 1 (rotate-half,
 * theta = rope_theta; modeling_llama.py:124-168) and appends the current k / v rows to the pools itself.  max_ctx >= ctx_len
 * sizes the flash-decoding split grid exactly as a captured decode graph does (splits beyond ctx_len are empty).
 * out_dev [B, heads*128]. */
int vcb_op_attention_decode_f16(const vcb_half* qkv_dev, vcb_half* k_pool_dev, vcb_half* v_pool_dev,
                                const int32_t* block_table_dev, int32_t max_pages, int32_t B, int32_t heads, int32_t ctx_len,
                                int32_t max_ctx, int32_t fuse_rope, float rope_theta, float scale, vcb_half* out_dev,
                                void* stream);

/* HF GenerationMixin.sample() on a batch of fp16 logit rows [B,V] (V <= 32768): scores / temperature ->
 * TopKLogitsWarper(top_k; 0 = off; ties with the k-th score are kept) -> TopPLogitsWarper(top_p) -> softmax -> one draw per
 * row (transformers/generation/logits_process.py, generation/utils.py sample(); reached from serve/cli.py:121-132 with
 * do_sample=True).  The draw is the inverse CDF of the filtered distribution, in descending-score order (index ascending
 * among equal scores), at u = vcb_philox_uniform(seed, step, row).  kept_dev (optional) receives the number of tokens
 * that survived the filters per row. */
int vcb_op_sample_f16(const vcb_half* logits_dev, int32_t B, int32_t V, float temperature, int32_t top_k, float top_p,
                      uint64_t seed, int32_t step, int32_t* next_ids_dev, int32_t* kept_dev, void* stream);

/* process_images + CLIPImageProcessor.preprocess for ONE image (vcoder_llava/mm_utils.py:14-40; transformers 4.31
 * image_processing_clip.py: resize(shortest_edge = out_size, PIL BICUBIC) -> center_crop(out_size) -> rescale(1/255) ->
 * normalize(mean, std)) followed by the callers' `.to(device, dtype=torch.float16)` (serve/cli.py:74-88).
 * img_any: uint8 [height, width, 3] RGB, host or device.  pad_rgb: NULL, or the 3 background bytes of expand2square
 * (`image_aspect_ratio == 'pad'`: tuple(int(x*255) for x in image_mean)).  out_dev: fp16 [3, out_size, out_size].
 * The resize is bit-exact with Pillow's ImagingResample (integer arithmetic).  Synchronous. */
int vcb_op_preprocess_u8(const uint8_t* img_any, int32_t height, int32_t width, const uint8_t* pad_rgb, int32_t out_size,
                         const float* mean3, const float* std3, vcb_half* out_dev, void* stream);

/* bitsandbytes NF4 weight format as a load-time transform (reference: `load_pretrained_model(load_4bit=True)`,
 * vcoder_llava/model/builder.py:28-37 -> BitsAndBytesConfig(nf4, double quant, compute dtype fp16)): out = the fp16 weights
 * a bnb.nn.Linear4bit layer multiplies by, dequantize(quantize(w)).  n % 64 == 0 (blocks of 64 weights, absmax per block,
 * nearest of the 16 NF4 levels); double_quant != 0: the absmax vector minus its mean goes through the 8-bit dynamic code in
 * blocks of 256 (code8_host: the 256 sorted code values of create_dynamic_map, see vcoder_b200/nf4.py).  idx_out_dev
 * (optional, n bytes) receives the 4-bit level indices.  w_dev may equal out_dev.  The 4-bit storage itself is not kept: the
 * engine streams the de-quantised fp16 weights (same arithmetic, none of the memory saving).  PARITY UNPINNED: bitsandbytes
 * is not available offline; the oracle restates the published algorithm (oracle/bnb_nf4_oracle.py). */
int vcb_op_nf4_roundtrip_f16(const vcb_half* w_dev, int64_t n, int32_t double_quant, const float* code8_host, vcb_half* out_dev,
                             uint8_t* idx_out_dev, void* stream);

/* ---- host-only logic ------------------------------------------------------------------------ */

/* The splice plan of prepare_inputs_labels_for_multimodal (vcoder_ds_llava_arch.py:175-312 /
 * vcoder_llava_arch.py:182-294 / llava_arch.py:117-199 / vcoder_it_llava_arch.py:165-277) for ids_host [B,T].
 * depth_nonzero_host: VCoder-DS: NULL when depths is None, else B bytes (`mean(depth) != 0`); VCoder-IT: the same for the
 * seg maps (`mean(seg) != 0`, is_seg_zero), NULL when segs is None; ignored by the other variants.  plan_out (capacity plan_cap int32) receives B*out_len entries:
 *   >= 0 token id, -1 zero pad, <= -2 row -(v+2) of the feature buffer laid out [kind(rgb,seg,depth)][B][1+P] (row 0 of
 *   each image is the CLS token and is never referenced).  Needs no GPU. */
int vcb_splice_plan(int32_t variant, const int64_t* ids_host, int32_t B, int32_t T, int32_t n_patches, int32_t has_seg,
                    const uint8_t* depth_nonzero_host, int32_t* plan_out, int64_t plan_cap, int32_t* out_len);

/* The uniform in [0,1) the sampler uses for (seed, step = index of the token since the prompt, row = sequence in the
 * batch): top 24 bits of the first word of Philox4x32-10 with key = seed, counter = (step, row, 0, 0).  Needs no GPU. */
float vcb_philox_uniform(uint64_t seed, uint32_t step, uint32_t row);

/* ---- engine --------------------------------------------------------------------------------- */

/* Replaces load_pretrained_model's model construction (vcoder_llava/model/builder.py:93-143): allocates packed
 * weight storage, workspaces and the paged KV cache on `device`. */
int vcb_create(const vcb_config* cfg, int32_t device, vcb_handle* out);
int vcb_destroy(vcb_handle h);

/* Replaces from_pretrained's state-dict load. Tensors are COPIED (and packed: q|k|v concatenated, gate/up
 * interleaved) into handle-owned memory, so the caller may free its copies afterwards.  Unknown or dead names
 * (depth_mm_projector.*, mm2_projector.*, vcoder_lm_emb.weight, post_layernorm.*, encoder layers past the tapped
 * one) are accepted and ignored, as the reference never reads them at inference. */
int vcb_bind_weights(vcb_handle h, const vcb_tensor* tensors, int32_t n);

/* CLIPVisionTower.forward (clip_encoder.py:39-51): pixels_any [n,3,S,S] fp16 -> out_any [n,P,D] fp16
 * = hidden_states[select_layer][:, 1:].  Any n >= 1: batches beyond the handle's vision workspace (3*max_batch images)
 * run in chunks of that size. */
int vcb_tower_forward(vcb_handle h, const vcb_half* pixels_any, int32_t n, vcb_half* out_any, void* stream);

/* encode_images (modality 0: mm_projector), encode_seg_images (1) / encode_depth_images (2): both seg_mm_projector
 * (vcoder_ds_llava_arch.py:106-119).  pixels_any [n,3,S,S] -> out_any [n,P,H]. */
int vcb_encode(vcb_handle h, int32_t modality, const vcb_half* pixels_any, int32_t n, vcb_half* out_any, void* stream);

/* VCoder*LlavaLlamaForCausalLM.forward at prefill (vcoder_ds_llava_llama.py:57-118): towers + projectors + splice +
 * decoder stack + lm_head on the last position + greedy argmax.  ids_host [B,T] int64 with the -200/-300/-400
 * specials; images/segs/depths_any [B,3,S,S] fp16 (segs/depths may be NULL; images NULL = text-only).
 * Optional outputs (any, may be NULL): next_ids [B] int32, last_logits [B,V] fp16, embeds [B,T',H] fp16 (the spliced
 * inputs_embeds), spliced_len (host int32). Resets the KV cache. */
int vcb_prefill(vcb_handle h, const int64_t* ids_host, int32_t B, int32_t T, const vcb_half* images_any,
                const vcb_half* segs_any, const vcb_half* depths_any, int32_t* next_ids_any, vcb_half* last_logits_any,
                vcb_half* embeds_any, int32_t* spliced_len_host, void* stream);

/* The forward() seam: VCoderDSLlavaLlamaForCausalLM.forward(input_ids, images=, segs=, depths=) with labels = None and
 * past_key_values = None (language_model/vcoder_ds_llava_llama.py:57-118; CausalLMOutputWithPast.logits).  Same inputs
 * as vcb_prefill; logits_dev (DEVICE memory, logits_halves >= B * T' * V fp16 elements; T' from vcb_splice_plan)
 * receives lm_head(norm(hidden)) of EVERY spliced position, row-major [B, T', V].  The KV cache is filled and the first
 * greedy token selected exactly as by vcb_prefill, so vcb_decode may follow.  VCB_ERR_NOMEM when the buffer is small. */
int vcb_forward_logits(vcb_handle h, const int64_t* ids_host, int32_t B, int32_t T, const vcb_half* images_any,
                       const vcb_half* segs_any, const vcb_half* depths_any, vcb_half* logits_dev, int64_t logits_halves,
                       int32_t* spliced_len_host, void* stream);

/* The prepare_inputs_labels_for_multimodal seam (vcoder_ds_llava_arch.py:126-314; its `new_input_embeds` result):
 * towers + projectors + splice only.  embeds_any [B, T', H] fp16 (host or device).  No decoder pass, no KV: the handle
 * is left un-prefilled (vcb_decode fails until the next vcb_prefill). */
int vcb_splice_embeds(vcb_handle h, const int64_t* ids_host, int32_t B, int32_t T, const vcb_half* images_any,
                      const vcb_half* segs_any, const vcb_half* depths_any, vcb_half* embeds_any,
                      int32_t* spliced_len_host, void* stream);

/* vcb_prefill for multi-turn callers.  serve/chat.py:41-159 (Chat.generate_stream) and the
 * serve/cli.py:80-135 loop re-tokenise the WHOLE conversation and re-send every image on every turn, and
 * prepare_inputs_for_generation re-threads images/segs/depths (vcoder_ds_llava_llama.py:137-139), so the reference
 * re-runs the three towers and re-prefills the 1152 visual tokens per turn.  Same arguments and results as vcb_prefill
 * (no `embeds` output), but:
 *   - the pixels are fingerprinted on the device; when they equal the previous call's (same B, same modalities) the
 *     towers and projectors are skipped (*towers_skipped = 1) and the projected features kept in the handle are used;
 *   - positions [0, *reused_len) — the longest prefix, common to every sequence of the batch, that the new spliced
 *     prompt shares with what the KV cache already holds (previous spliced prompt + the generated tokens already fed
 *     back) — keep their KV entries; only positions [*reused_len, T') run through the decoder.  Visual positions are
 *     only reused when the features are the same.  The last position always runs.
 * Row arithmetic is position-wise identical to a full prefill; KV entries of previously GENERATED tokens were produced
 * by the decode kernels, so logits agree with a full re-prefill to fp16 round-off, not bit for bit. */
int vcb_prefill_cached(vcb_handle h, const int64_t* ids_host, int32_t B, int32_t T, const vcb_half* images_any,
                       const vcb_half* segs_any, const vcb_half* depths_any, int32_t* next_ids_any,
                       vcb_half* last_logits_any, int32_t* spliced_len_host, int32_t* reused_len_host,
                       int32_t* towers_skipped_host, void* stream);

/* Host-only part of the above: the longest prefix (<= limit) shared by every row of cached [B, cached_len] and
 * plan [B, plan_len] (vcb_splice_plan entries); visual rows (<= -2) only match when features_same != 0.  Needs no GPU. */
int vcb_prefix_match(const int32_t* cached_host, int32_t cached_len, const int32_t* plan_host, int32_t plan_len, int32_t B,
                     int32_t features_same, int32_t limit, int32_t* out_len);

/* n_steps greedy decode steps (prepare_inputs_for_generation + forward + argmax, vcoder_ds_llava_llama.py:120-142;
 * every cached position is attended, padding included: vcoder_ds_llava_arch.py:130-133).  out_ids_any [B,n_steps]
 * int32 receives the tokens produced by THIS call; logits_trace_any [n_steps,B,V] fp16 optional (disables the CUDA
 * graph).  Asynchronous on `stream` unless an output is pageable host memory. */
int vcb_decode(vcb_handle h, int32_t n_steps, int32_t* out_ids_any, vcb_half* logits_trace_any, void* stream);

/* model.generate(input_ids, images=, segs=, depths=, do_sample=False, max_new_tokens=N) for HOST buffers: prefill +
 * N-1 decode steps, copies the N new ids per sequence to out_ids_host [B,N] int64 and synchronises.  This is the
 * end-to-end entry the callers (serve/cli.py:121-132, eval/model_depth_loader.py:146-168) map onto. */
int vcb_generate(vcb_handle h, const int64_t* ids_host, int32_t B, int32_t T, const vcb_half* images_any,
                 const vcb_half* segs_any, const vcb_half* depths_any, int32_t max_new_tokens, int64_t* out_ids_host,
                 void* stream);

/* generate(do_sample=, temperature=, top_k=, top_p=) for the following vcb_prefill / vcb_decode* / vcb_generate calls on
 * this handle (serve/cli.py:121-132 passes do_sample=True, temperature; serve/chat.py:141-151 adds top_p; top_k = 50 is
 * the GenerationConfig default).  do_sample = 0 restores greedy argmax.  Invalid values fail with VCB_ERR_INVALID and
 * the message HF's warpers raise (ValueError in the shim). */
int vcb_set_sampling(vcb_handle h, int32_t do_sample, float temperature, int32_t top_k, float top_p, uint64_t seed,
                     void* stream);

/* Per-token streaming decode for TextStreamer / KeywordsStoppingCriteria / eos handling (serve/cli.py:117-132,
 * mm_utils.py:128-151, serve/chat.py:131-159): up to max_steps tokens; `lookahead` (1..16) decode steps stay queued on
 * the GPU while `cb(user, step, ids_host[B], B)` runs for the token of step `step` (0-based within this call).  A
 * non-zero return stops after that token: the steps already queued are discarded and the KV length / RNG step are
 * rolled back, so a later vcb_decode* continues right after the accepted token.  *n_accepted = tokens delivered.
 * Synchronous. */
typedef int (*vcb_token_callback)(void* user, int32_t step, const int32_t* ids_host, int32_t B);
int vcb_decode_stream(vcb_handle h, int32_t max_steps, int32_t lookahead, vcb_token_callback cb, void* user,
                      int32_t* n_accepted, void* stream);

/* Drops the KV cache contents (pages return to the pool). */
int vcb_reset(vcb_handle h);

/* ---- accounting (bench.py: `gpu_launches`, live per-kernel timings for the roofline) ------------------ */

/* Kernels launched by this library in this process so far (graph replays count their kernel nodes). */
uint64_t vcb_launch_count(void);

/* Per-category CUDA-event timing of every launch (disables the decode CUDA graph while on).
 * Categories: 0 tcgen05 GEMM, 1 skinny GEMM, 2 ViT attention, 3 LLaMA prefill attention, 4 decode attention,
 * 5 LayerNorm/RMSNorm, 6 RoPE + KV append, 7 gather/argmax/misc. */
#define VCB_PROFILE_CATEGORIES 8
int vcb_profile_enable(int32_t on);
int vcb_profile_read(float* ms_by_category, uint64_t* launches_by_category, int32_t n_categories);

/* Hardware micro-benchmarks that size the attention kernels (tools/microbench.py; never called by the engine).
 * which = 0: TMEM -> register read rate with `warps` warps per CTA and `param` tcgen05.ld.32x32b.x32 in flight per warp;
 * *result = bytes per clock per SM.  which = 1: MUFU ex2.approx rate (param != 0: one FFMA between the ex2);
 * *result = ex2 per clock per SM.  which = 2: the packed ex2.approx.f16x2 (exponentials per clock per SM).
 * which = 3: the exponential / pack / store loop of the ViT attention softmax in isolation (param bit 0: no row sums,
 * bit 1: no shared-memory stores, bit 2: software-skewed); *result = exponentials per clock per SM. */
int vcb_microbench(int32_t which, int32_t warps, int32_t param, int32_t iters, float* result);

/* Tuning hook (tools/bench_decode.py): pipeline shape of the weight-streaming kernel for GEMMs with this (N, K): `warps`
 * per CTA (2, 4, 8; 0 = built-in choice) x `depth` stages of 8 KiB per warp (1..4; 0 = built-in choice).  Results do not
 * depend on it bit for bit only when `warps` is unchanged (the split of K across warps fixes the summation order).
 * Captured decode graphs keep the shape they were captured with. */
int vcb_tune_skinny(int32_t N, int32_t K, int32_t warps, int32_t depth);

/* Device-side timeline of the decode-step kernels (tools/trace_decode.py; evidence for DESIGN.md section 3.1).  Between
 * vcb_trace_begin and vcb_trace_read every CTA of the weight-streaming GEMM, the decode attention kernels and the small
 * RMSNorm launches stamps %globaltimer (ns) at entry, after its grid-dependency wait and at its end; works inside the
 * captured decode graph (a replay overwrites the slots, so the LAST step is what is read).  launches_out [max_launches][4]
 * = {kind (0 GEMM, 1 attention split, 2 attention combine, 3 rmsnorm), CTAs, first slot, aux (N / splits / H)};
 * stamps_out [cap] uint64, three per CTA.  vcb_trace_read synchronises the device and closes the trace. */
int vcb_trace_begin(int32_t max_ctas);
int vcb_trace_read(int32_t* launches_out, int32_t max_launches, int32_t* n_launches, uint64_t* stamps_out, int64_t cap);

/* The decode step's weight-streaming GEMMs alone (every layer's qkv, o, gate|up, down, then lm_head) on the bound
 * weights with batch B, back to back as inside the decode graph, `rounds` times between two CUDA events on the
 * launching stream (one untimed warm-up round first).  *ms_total = elapsed ms, *launches = kernels timed.  Invalidates
 * the current KV state (call vcb_prefill again before decoding). */
int vcb_bench_weight_stream(vcb_handle h, int32_t B, int32_t rounds, float* ms_total, int32_t* launches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VCODER_B200_H_ */
