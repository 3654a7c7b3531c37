"""CPU ORACLE (TEST INFRASTRUCTURE ONLY) — where HF's fp16 GPU path rounds, and what the 4.31 <-> 5.5 skew is worth.

`oracle/vcoder_oracle.py` evaluates the path in fp32 from the fp16 checkpoint values: it is the "exact" answer that every
fp16 implementation scatters around.  The reference's real deployment (`model/builder.py:39` torch_dtype=float16 on CUDA) is
HF eager fp16: every op output is ROUNDED to fp16 (fp32 accumulation inside a GEMM / softmax / norm kernel, one rounding on
the way out).  This module restates the same arithmetic with those roundings made explicit — `r(x) = x.half().float()` after
each op HF runs as a separate fp16 kernel — in two styles:

  style "4.31" (the pinned `transformers==4.31.0`, reference pyproject.toml:23):
      CLIP   modeling_clip.py:261-345  q = q_proj(h) * scale BEFORE the bmm; softmax on the fp16 scores
      LLaMA  modeling_llama.py:124-168 cos/sin cache stored in fp16; :199-222 scores = (q @ k^T) / sqrt(dh), softmax fp32 -> fp16
  style "5.5" (the transformers this container has; `tools/make_golden.py` ran the reference on it, in fp32 on the CPU):
      CLIP   eager_attention_forward: (q @ k^T) * scale AFTER the matmul, softmax(dtype=float32).to(fp16)
      LLaMA  rotary: cos/sin computed in fp32, cast to fp16 at use; scores = (q @ k^T) * dh**-0.5

SURVEY.md section 8(c) lists exactly these as the version deltas "to keep inside the budget".  `tests/test_oracle_rounding_host.py`
MEASURES them: end to end at the tiny geometry the two styles differ by 3.2e-4 rel-RMS on the logits (max 3.9e-3) while
either sits 1.8e-3 from the fp32 reference fixture; over two LLaMA-7B-width layers 6.0e-4 against 5.6e-3.  (Scale 0.125 is a
power of two, so the CLIP difference is nil; x / sqrt(128) and x * 128**-0.5 differ by one fp16 ulp on some scores.)  So the
version skew is 10-20 % of the fp16 noise of the reference's OWN GPU path, and that noise is what the GPU suite's
tolerances (rel-RMS 2e-3 / 2.5e-3 at real width, |dlogit| 1e-2 / 2e-2) are sized against — the fused
kernels round less often than an op-by-op fp16 path and are held closer to fp32 than HF-fp16 itself gets.

Parity status: this file emulates third-party kernels' rounding points from their source, not their bits (cuBLAS summation
order is not reproduced) — it bounds noise, it is not a bit-exact oracle.  Nothing under vcoder_b200/ imports it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

import vcoder_oracle as orc

_VT = orc._VT
_f = orc._f


def r(x: torch.Tensor) -> torch.Tensor:
    """One fp16 kernel boundary: the value HF's next op reads."""
    return x.half().float()


def _lin(x, sd, name, bias=True):
    return r(F.linear(x, _f(sd[name + ".weight"]), _f(sd[name + ".bias"]) if bias and (name + ".bias") in sd else None))


def clip_layer(h, sd, prefix, heads, eps, style):
    B, S, D = h.shape
    dh = D // heads
    scale = dh ** -0.5
    a = r(F.layer_norm(h, (D,), _f(sd[prefix + "layer_norm1.weight"]), _f(sd[prefix + "layer_norm1.bias"]), eps))
    q, k, v = (_lin(a, sd, prefix + f"self_attn.{n}_proj").view(B, S, heads, dh).transpose(1, 2) for n in "qkv")
    if style == "4.31":
        scores = r(r(q * scale) @ k.transpose(-1, -2))      # modeling_clip.py (4.31) :288, :300
    else:
        scores = r(r(q @ k.transpose(-1, -2)) * scale)      # eager_attention_forward (5.x)
    att = r(torch.softmax(scores, dim=-1))                  # fp16 softmax = fp32 inside the kernel, rounded once
    o = r(att @ v).transpose(1, 2).reshape(B, S, D)
    h = r(h + _lin(o, sd, prefix + "self_attn.out_proj"))
    m = r(F.layer_norm(h, (D,), _f(sd[prefix + "layer_norm2.weight"]), _f(sd[prefix + "layer_norm2.bias"]), eps))
    m = _lin(m, sd, prefix + "mlp.fc1")
    m = r(m * r(torch.sigmoid(r(1.702 * m))))               # QuickGELUActivation, three fp16 elementwise kernels
    return r(h + _lin(m, sd, prefix + "mlp.fc2"))


def clip_tower_forward(pixels, sd, g, style):
    D = g.vit_hidden
    x = r(_f(pixels))
    w = _f(sd[_VT + "embeddings.patch_embedding.weight"])
    B, p = x.shape[0], g.vit_patch
    n = g.vit_image_size // p
    patches = x.reshape(B, 3, n, p, n, p).permute(0, 2, 4, 1, 3, 5).reshape(B, n * n, 3 * p * p)
    pe = r(patches @ w.reshape(w.shape[0], -1).t())
    cls = _f(sd[_VT + "embeddings.class_embedding"]).expand(B, 1, -1)
    h = r(torch.cat([cls, pe], dim=1) + _f(sd[_VT + "embeddings.position_embedding.weight"])[None])
    h = r(F.layer_norm(h, (D,), _f(sd[_VT + "pre_layrnorm.weight"]), _f(sd[_VT + "pre_layrnorm.bias"]), g.vit_ln_eps))
    for i in range(g.vit_layers_used):
        h = clip_layer(h, sd, f"{_VT}encoder.layers.{i}.", g.vit_heads, g.vit_ln_eps, style)
    return h[:, 1:]


def projector(x, sd, name):
    h = _lin(x, sd, f"model.{name}.0")
    return _lin(r(F.gelu(h)), sd, f"model.{name}.2")


def rms_norm(x, w, eps):
    var = x.pow(2).mean(-1, keepdim=True)                   # LlamaRMSNorm: fp32 statistics (modeling_llama.py:53-70)
    return r(w * r(x * torch.rsqrt(var + eps)))             # .to(input_dtype), then the fp16 multiply by the weight


def llama_layer(x, sd, i, g, pos0, kc, vc, style):
    B, T, H = x.shape
    heads, dh = g.llm_heads, g.llm_hidden // g.llm_heads
    p = f"model.layers.{i}."
    a = rms_norm(x, _f(sd[p + "input_layernorm.weight"]), g.llm_rms_eps)
    q, k, v = (_lin(a, sd, p + f"self_attn.{n}_proj", bias=False).view(B, T, heads, dh).transpose(1, 2) for n in "qkv")
    cos, sin = orc.rope_cos_sin(torch.arange(pos0, pos0 + T), dh, g.llm_rope_theta)
    cos, sin = r(cos)[None, None], r(sin)[None, None]       # both versions hand fp16 cos/sin to apply_rotary_pos_emb
    q = r(r(q * cos) + r(orc.rotate_half(q) * sin))
    k = r(r(k * cos) + r(orc.rotate_half(k) * sin))
    if kc[i] is not None:
        k, v = torch.cat([kc[i], k], dim=2), torch.cat([vc[i], v], dim=2)
    kc[i], vc[i] = k, v
    S = k.shape[2]
    if style == "4.31":
        scores = r(r(q @ k.transpose(-1, -2)) / math.sqrt(dh))   # modeling_llama.py (4.31) :199
    else:
        scores = r(r(q @ k.transpose(-1, -2)) * dh ** -0.5)      # eager_attention_forward (5.x)
    qpos = torch.arange(pos0, pos0 + T)[:, None]
    scores = scores.masked_fill(torch.arange(S)[None, :] > qpos, float("-inf"))
    att = r(torch.softmax(scores, dim=-1))                  # softmax(dtype=float32).to(fp16)
    o = r(att @ v).transpose(1, 2).reshape(B, T, H)
    x = r(x + _lin(o, sd, p + "self_attn.o_proj", bias=False))
    m = rms_norm(x, _f(sd[p + "post_attention_layernorm.weight"]), g.llm_rms_eps)
    gate, up = _lin(m, sd, p + "mlp.gate_proj", bias=False), _lin(m, sd, p + "mlp.up_proj", bias=False)
    return r(x + _lin(r(r(F.silu(gate)) * up), sd, p + "mlp.down_proj", bias=False))


def llama_forward(x, sd, g, style, n_layers: Optional[int] = None):
    """x [B,T,H] embeddings -> logits [B,T,V] as `LlamaForCausalLM` in fp16 would produce them (then .float())."""
    n = g.llm_layers if n_layers is None else n_layers
    kc, vc = [None] * n, [None] * n
    x = r(x)
    for i in range(n):
        x = llama_layer(x, sd, i, g, 0, kc, vc, style)
    x = rms_norm(x, _f(sd["model.norm.weight"]), g.llm_rms_eps)
    return r(F.linear(x, _f(sd["lm_head.weight"])))


def prefill_logits(ids, sd, g, images, segs, depths, style):
    """The whole multimodal prefill (towers -> projectors -> splice -> decoder -> lm_head) with HF-fp16 rounding points;
    the splice itself is index arithmetic and is shared with the fp32 oracle."""
    variant_ds = g.variant == orc.V_DS
    img_f = projector(clip_tower_forward(images, sd, g, style), sd, "mm_projector")
    seg_f = projector(clip_tower_forward(segs, sd, g, style), sd, "seg_mm_projector") if segs is not None else None
    dep_f, nz = None, None
    if variant_ds and depths is not None:
        nz = [bool(_f(d).mean() != 0) for d in depths]
        dep_f = projector(clip_tower_forward(depths, sd, g, style), sd, "seg_mm_projector")
    x = orc.splice_embeds(ids, sd, img_f, seg_f, dep_f, variant_ds, nz, g.variant, None)
    return llama_forward(x, sd, g, style)
