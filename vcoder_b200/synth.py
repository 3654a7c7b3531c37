"""Synthetic model geometry, seeded random-init weights and synthetic prompts (no checkpoints, tokenizer files or
datasets exist offline).  Weight names are exactly the keys of the reference model's ``state_dict()`` so the same
dict feeds (a) the reference itself (``load_state_dict`` in tools/make_golden.py), (b) the CPU oracle and (c)
``vcb_bind_weights``.

Geometry presets:
  tiny     – fast CI / golden fixtures (CLIP hidden 128, 4 layers, 56 px; LLaMA hidden 256, 2 layers, vocab 512)
  small    – second golden geometry (CLIP hidden 256, 6 layers, 112 px = 65 tokens; LLaMA hidden 512, 3 layers)
  vit_l    – real CLIP ViT-L/14@336 tower with a 2-layer LLaMA-7B-shaped decoder (mid-size parity on the GPU box)
  7b / 13b – VCoder(-DS) LLaVA-1.5-7B / 13B (BASELINE.json configs 2-4)
"""
from __future__ import annotations

import dataclasses
import hashlib
import math
from typing import Dict

import torch

IMAGE_TOKEN_INDEX = -200  # reference vcoder_llava/constants.py:9
SEG_TOKEN_INDEX = -300  # constants.py:10
DEPTH_TOKEN_INDEX = -400  # constants.py:11

VARIANT_VCODER = 0     # VCoderLlavaLlamaForCausalLM    (RGB + seg)
VARIANT_VCODER_DS = 1  # VCoderDSLlavaLlamaForCausalLM  (RGB + seg + depth)
VARIANT_LLAVA = 2      # LlavaLlamaForCausalLM          (RGB only; language_model/llava_llama.py, model/llava_arch.py)
VARIANT_VCODER_IT = 3  # VCoderITLlavaLlamaForCausalLM  (RGB + optional seg; model/vcoder_it_llava_arch.py)


@dataclasses.dataclass(frozen=True)
class ModelGeometry:
    name: str
    variant: int = VARIANT_VCODER_DS
    # CLIP vision tower
    vit_image_size: int = 336
    vit_patch: int = 14
    vit_hidden: int = 1024
    vit_heads: int = 16
    vit_ffn: int = 4096
    vit_layers: int = 24
    vit_select_layer: int = -2
    vit_ln_eps: float = 1e-5
    # LLaMA
    llm_hidden: int = 4096
    llm_layers: int = 32
    llm_heads: int = 32
    llm_ffn: int = 11008
    llm_vocab: int = 32000
    llm_rms_eps: float = 1e-5
    llm_rope_theta: float = 10000.0
    # init
    init_gain: float = 1.0      # LLaMA-side std multiplier
    vit_init_gain: float = 1.0  # CLIP-side std multiplier

    @property
    def n_patches(self) -> int:
        return (self.vit_image_size // self.vit_patch) ** 2

    @property
    def vit_layers_used(self) -> int:
        # hidden_states[select_layer] with len(hidden_states) == vit_layers + 1 (clip_encoder.py:29-31)
        idx = self.vit_select_layer if self.vit_select_layer >= 0 else self.vit_layers + 1 + self.vit_select_layer
        return idx

    def with_variant(self, variant: int) -> "ModelGeometry":
        return dataclasses.replace(self, variant=variant)


PRESETS: Dict[str, ModelGeometry] = {
    "tiny": ModelGeometry(
        name="tiny", vit_image_size=56, vit_hidden=128, vit_heads=2, vit_ffn=256, vit_layers=4,
        llm_hidden=256, llm_layers=2, llm_heads=2, llm_ffn=512, llm_vocab=512, init_gain=3.0),
    # second pinning geometry of the oracle (tests/golden/small_ds.npz): 64 patches + CLS = 65 tokens (S % 64 == 1 like the
    # real tower's 577), head counts > 2, more layers than `tiny`, FFN not a power of two
    "small": ModelGeometry(
        name="small", vit_image_size=112, vit_hidden=256, vit_heads=4, vit_ffn=640, vit_layers=6,
        llm_hidden=512, llm_layers=3, llm_heads=4, llm_ffn=1408, llm_vocab=512, init_gain=3.0),
    "vit_l": ModelGeometry(name="vit_l", llm_layers=2),
    "7b": ModelGeometry(name="7b"),
    "13b": ModelGeometry(name="13b", llm_hidden=5120, llm_layers=40, llm_heads=40, llm_ffn=13824),
}

_VT = "model.vision_tower.vision_tower.vision_model."


def weight_specs(g: ModelGeometry):
    """[(name, shape, kind, std)]: kind in {'normal','ln_weight','bias'}. Names = reference state_dict keys."""
    d, f, L = g.vit_hidden, g.vit_ffn, g.vit_layers
    k = g.vit_init_gain
    specs = []
    # CLIPVisionEmbeddings (transformers/models/clip/modeling_clip.py:138-160); std as HF _init_weights
    specs.append((_VT + "embeddings.class_embedding", (d,), "normal", d ** -0.5))
    specs.append((_VT + "embeddings.patch_embedding.weight", (d, 3, g.vit_patch, g.vit_patch), "normal", 0.02))
    specs.append((_VT + "embeddings.position_embedding.weight", (g.n_patches + 1, d), "normal", 0.02))
    specs.append((_VT + "pre_layrnorm.weight", (d,), "ln_weight", 0.1))
    specs.append((_VT + "pre_layrnorm.bias", (d,), "bias", 0.02))
    in_std = (d ** -0.5) * ((2 * L) ** -0.5) * k
    out_std = (d ** -0.5) * k
    fc_std = ((2 * d) ** -0.5) * k
    for i in range(L):
        p = f"{_VT}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj"):
            specs.append((p + f"self_attn.{n}.weight", (d, d), "normal", in_std * 4))
            specs.append((p + f"self_attn.{n}.bias", (d,), "bias", 0.02))
        specs.append((p + "self_attn.out_proj.weight", (d, d), "normal", out_std))
        specs.append((p + "self_attn.out_proj.bias", (d,), "bias", 0.02))
        specs.append((p + "layer_norm1.weight", (d,), "ln_weight", 0.1))
        specs.append((p + "layer_norm1.bias", (d,), "bias", 0.02))
        specs.append((p + "mlp.fc1.weight", (f, d), "normal", fc_std))
        specs.append((p + "mlp.fc1.bias", (f,), "bias", 0.02))
        specs.append((p + "mlp.fc2.weight", (d, f), "normal", in_std))
        specs.append((p + "mlp.fc2.bias", (d,), "bias", 0.02))
        specs.append((p + "layer_norm2.weight", (d,), "ln_weight", 0.1))
        specs.append((p + "layer_norm2.bias", (d,), "bias", 0.02))
    specs.append((_VT + "post_layernorm.weight", (d,), "ln_weight", 0.1))  # dead at inference
    specs.append((_VT + "post_layernorm.bias", (d,), "bias", 0.02))
    # projectors: mlp2x_gelu (multimodal_projector/builder.py:39-46); depth/mm2 exist in checkpoints but are never
    # called at inference — they are generated so that bind must accept and ignore them.
    H = g.llm_hidden
    if g.variant == VARIANT_LLAVA:
        projs = ["mm_projector"]  # llava_arch.py:36-38
    elif g.variant == VARIANT_VCODER_IT:
        projs = ["mm_projector", "seg_mm_projector"]  # vcoder_it_llava_arch.py:32-38 (no mm2 / vcoder_lm_emb)
    else:
        projs = ["mm_projector", "seg_mm_projector", "mm2_projector"]
        if g.variant == VARIANT_VCODER_DS:
            projs.insert(2, "depth_mm_projector")
    for pj in projs:
        specs.append((f"model.{pj}.0.weight", (H, d), "normal", (d ** -0.5)))
        specs.append((f"model.{pj}.0.bias", (H,), "bias", 0.02))
        specs.append((f"model.{pj}.2.weight", (H, H), "normal", (H ** -0.5)))
        specs.append((f"model.{pj}.2.bias", (H,), "bias", 0.02))
    # LLaMA (HF init: N(0, 0.02^2); RMSNorm weight 1 — perturbed here so the scale path is exercised)
    s = 0.02 * g.init_gain
    specs.append(("model.embed_tokens.weight", (g.llm_vocab, H), "normal", s))
    # vcoder_lm_emb is overwritten with a clone of embed_tokens on every prefill (vcoder_ds_llava_arch.py:173,
    # SURVEY F7): whatever the checkpoint holds is dead.  Generated with different values to prove it is ignored.
    if g.variant in (VARIANT_VCODER, VARIANT_VCODER_DS):
        specs.append(("model.vcoder_lm_emb.weight", (g.llm_vocab, H), "normal", s))
    for i in range(g.llm_layers):
        p = f"model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            specs.append((p + f"self_attn.{n}.weight", (H, H), "normal", s))
        specs.append((p + "mlp.gate_proj.weight", (g.llm_ffn, H), "normal", s))
        specs.append((p + "mlp.up_proj.weight", (g.llm_ffn, H), "normal", s))
        specs.append((p + "mlp.down_proj.weight", (H, g.llm_ffn), "normal", s))
        specs.append((p + "input_layernorm.weight", (H,), "ln_weight", 0.1))
        specs.append((p + "post_attention_layernorm.weight", (H,), "ln_weight", 0.1))
    specs.append(("model.norm.weight", (H,), "ln_weight", 0.1))
    specs.append(("lm_head.weight", (g.llm_vocab, H), "normal", s))
    return specs


def _seed_for(name: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:7], "little")


def make_state_dict(g: ModelGeometry, seed: int = 0, device: str = "cpu", only_prefix: str | None = None
                    ) -> Dict[str, torch.Tensor]:
    """fp16 tensors. On 'cpu' the values are reproducible bit-for-bit (per-tensor seeded CPU generator); on
    'cuda' they are generated on the device (same distribution, different values) for the full-size benches."""
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind, std in weight_specs(g):
        if only_prefix is not None and not name.startswith(only_prefix):
            continue
        gen = torch.Generator(device=device)
        gen.manual_seed(_seed_for(name, seed))
        t = torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        if kind == "ln_weight":
            t = 1.0 + std * t
        else:
            t = std * t
        out[name] = t.to(torch.float16)
    return out


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].cpu().contiguous().view(torch.int16).numpy().tobytes())
    return h.hexdigest()


def make_prompt_ids(g: ModelGeometry, batch: int, seed: int = 0, n_before: int = 34, n_after: int = 20,
                    specials=None) -> torch.Tensor:
    """[B, T] int64: [1] + text + specials + text.  Default specials follow the reference tokenizers:
    VCoder-DS `_tokenizer_depth_token` emits [IMAGE, DEPTH, SEG] (mm_utils.py:101-103), VCoder
    `tokenizer_seg_token` emits [IMAGE, SEG]. Text ids are >= 3 so pad/eos/bos never occur."""
    if specials is None:
        specials = {VARIANT_VCODER_DS: [IMAGE_TOKEN_INDEX, DEPTH_TOKEN_INDEX, SEG_TOKEN_INDEX],
                    VARIANT_LLAVA: [IMAGE_TOKEN_INDEX]}.get(g.variant, [IMAGE_TOKEN_INDEX, SEG_TOKEN_INDEX])
    gen = torch.Generator().manual_seed(1000 + seed)
    rows = []
    for _ in range(batch):
        a = torch.randint(3, g.llm_vocab, (n_before,), generator=gen)
        b = torch.randint(3, g.llm_vocab, (n_after,), generator=gen)
        rows.append(torch.cat([torch.tensor([1]), a, torch.tensor(specials, dtype=torch.int64), b]))
    return torch.stack(rows).to(torch.int64)


def make_pixels(g: ModelGeometry, batch: int, seed: int, device: str = "cpu") -> torch.Tensor:
    """[B,3,S,S] fp16 ~ N(0,1): the scale CLIPImageProcessor-normalised images have (mm_utils.py:28-40)."""
    gen = torch.Generator(device=device).manual_seed(77 + seed)
    return torch.randn(batch, 3, g.vit_image_size, g.vit_image_size, generator=gen, device=device).to(torch.float16)
