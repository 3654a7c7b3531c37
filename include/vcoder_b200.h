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
 * (output then has N/2 columns).  block_n: 0 auto, 128, 256. */
int vcb_op_gemm_f16(const vcb_half* A_dev, int32_t lda, const vcb_half* W_dev, int32_t ldw, int32_t M, int32_t N,
                    int32_t K, const vcb_half* bias_dev, const vcb_half* residual_dev, int32_t ldr, int32_t act,
                    vcb_half* C_dev, int32_t ldc, int32_t block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VCODER_B200_H_ */
