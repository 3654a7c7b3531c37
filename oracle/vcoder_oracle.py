"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A from-scratch CPU restatement (plain functional fp32 torch-on-CPU tensor ops, no nn.Module, no `transformers`
import) of the VCoder / VCoder-DS inference hot path.  It exists to CHECK libvcoder_b200; it is never shipped,
never imported by the product package `vcoder_b200/`, and only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may use it.

Why torch-on-CPU and not numpy: the reference's CPU path *is* ATen CPU kernels (HF transformers modules), so this
is both the most faithful arithmetic restatement and the fair multi-threaded CPU baseline; numpy's elementwise ops
are single-threaded and would under-state the reference.

Pinning status: the reference repository ships NO tests, fixtures or golden vectors, and its
arithmetic lives in the un-vendored third-party dependency `transformers==4.31.0` (reference pyproject.toml:23).
This oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container
(`tools/make_golden.py` imports /root/reference's `vcoder_llava` on top of the installed transformers 5.5.0, loads
the same seeded weights and dumps per-stage tensors + greedy ids to tests/golden/*.npz); `tests/test_oracle_golden.py`
checks every function below against those fixtures.

Every function cites the reference file:line (or the HF module the reference dispatches to) that it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200  # vcoder_llava/constants.py:9
SEG_TOKEN_INDEX = -300  # vcoder_llava/constants.py:10
DEPTH_TOKEN_INDEX = -400  # vcoder_llava/constants.py:11

_VT = "model.vision_tower.vision_tower.vision_model."


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32)


# ---------------------------------------------------------------------------------------------------------
# a1: CLIPVisionTower.forward + feature_select  (vcoder_llava/model/multimodal_encoder/clip_encoder.py:29-51)
#     arithmetic: transformers/models/clip/modeling_clip.py (embeddings :138-219, attention :261-345,
#     MLP :348-360, encoder layer :363-385) — third-party, restated here.
# ---------------------------------------------------------------------------------------------------------
def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    # transformers.activations.QuickGELUActivation: x * sigmoid(1.702 x)
    return x * torch.sigmoid(1.702 * x)


def clip_embeddings(pixels: torch.Tensor, sd: Dict[str, torch.Tensor], g) -> torch.Tensor:
    """[B,3,S,S] -> [B, 1+P, D]: conv(k=s=patch, no bias) -> flatten -> cat CLS -> + position embedding."""
    x = _f(pixels)
    w = _f(sd[_VT + "embeddings.patch_embedding.weight"])
    B = x.shape[0]
    p = g.vit_patch
    n = g.vit_image_size // p
    # conv with stride == kernel is a GEMM over (c, ky, kx)-ordered patches
    patches = x.reshape(B, 3, n, p, n, p).permute(0, 2, 4, 1, 3, 5).reshape(B, n * n, 3 * p * p)
    pe = patches @ w.reshape(w.shape[0], -1).t()
    cls = _f(sd[_VT + "embeddings.class_embedding"]).expand(B, 1, -1)
    emb = torch.cat([cls, pe], dim=1) + _f(sd[_VT + "embeddings.position_embedding.weight"])[None]
    return emb


def clip_layer(h: torch.Tensor, sd, prefix: str, heads: int, eps: float) -> torch.Tensor:
    B, S, D = h.shape
    dh = D // heads
    a = F.layer_norm(h, (D,), _f(sd[prefix + "layer_norm1.weight"]), _f(sd[prefix + "layer_norm1.bias"]), eps)
    q = F.linear(a, _f(sd[prefix + "self_attn.q_proj.weight"]), _f(sd[prefix + "self_attn.q_proj.bias"]))
    k = F.linear(a, _f(sd[prefix + "self_attn.k_proj.weight"]), _f(sd[prefix + "self_attn.k_proj.bias"]))
    v = F.linear(a, _f(sd[prefix + "self_attn.v_proj.weight"]), _f(sd[prefix + "self_attn.v_proj.bias"]))
    q = q.view(B, S, heads, dh).transpose(1, 2)
    k = k.view(B, S, heads, dh).transpose(1, 2)
    v = v.view(B, S, heads, dh).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (dh ** -0.5), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, D)
    h = h + F.linear(o, _f(sd[prefix + "self_attn.out_proj.weight"]), _f(sd[prefix + "self_attn.out_proj.bias"]))
    m = F.layer_norm(h, (D,), _f(sd[prefix + "layer_norm2.weight"]), _f(sd[prefix + "layer_norm2.bias"]), eps)
    m = quick_gelu(F.linear(m, _f(sd[prefix + "mlp.fc1.weight"]), _f(sd[prefix + "mlp.fc1.bias"])))
    h = h + F.linear(m, _f(sd[prefix + "mlp.fc2.weight"]), _f(sd[prefix + "mlp.fc2.bias"]))
    return h


def clip_hidden_states(pixels, sd, g, n_layers: Optional[int] = None) -> List[torch.Tensor]:
    """hidden_states tuple as HF returns it with output_hidden_states=True: [pre_layrnorm out, layer1 out, ...]."""
    D = g.vit_hidden
    h = clip_embeddings(pixels, sd, g)
    h = F.layer_norm(h, (D,), _f(sd[_VT + "pre_layrnorm.weight"]), _f(sd[_VT + "pre_layrnorm.bias"]), g.vit_ln_eps)
    hs = [h]
    n = g.vit_layers if n_layers is None else n_layers
    for i in range(n):
        h = clip_layer(h, sd, f"{_VT}encoder.layers.{i}.", g.vit_heads, g.vit_ln_eps)
        hs.append(h)
    return hs


def clip_tower_forward(pixels, sd, g) -> torch.Tensor:
    """CLIPVisionTower.forward: hidden_states[select_layer][:, 1:] ('patch' select, clip_encoder.py:29-37).
    Only the layers that feed the tapped hidden state are evaluated (the rest is dead compute)."""
    hs = clip_hidden_states(pixels, sd, g, n_layers=g.vit_layers_used)
    return hs[g.vit_layers_used][:, 1:]


# ---------------------------------------------------------------------------------------------------------
# a3: mlp2x_gelu projectors (multimodal_projector/builder.py:33-46, multimodal_adapter/builder.py:31-44)
# a2: encode_images / encode_seg_images / encode_depth_images (vcoder_ds_llava_arch.py:106-119)
# ---------------------------------------------------------------------------------------------------------
def projector(x: torch.Tensor, sd, name: str) -> torch.Tensor:
    h = F.linear(x, _f(sd[f"model.{name}.0.weight"]), _f(sd[f"model.{name}.0.bias"]))
    h = F.gelu(h)  # nn.GELU() default = exact erf
    return F.linear(h, _f(sd[f"model.{name}.2.weight"]), _f(sd[f"model.{name}.2.bias"]))


def encode_images(pixels, sd, g):
    return projector(clip_tower_forward(pixels, sd, g), sd, "mm_projector")  # vcoder_ds_llava_arch.py:116-119


def encode_seg_images(pixels, sd, g):
    return projector(clip_tower_forward(pixels, sd, g), sd, "seg_mm_projector")  # :106-109


def encode_depth_images(pixels, sd, g):
    # depth goes through seg_mm_projector, NOT depth_mm_projector (vcoder_ds_llava_arch.py:113)
    return projector(clip_tower_forward(pixels, sd, g), sd, "seg_mm_projector")


# ---------------------------------------------------------------------------------------------------------
# a4: prepare_inputs_labels_for_multimodal — the splice
#     DS:     vcoder_llava/model/vcoder_ds_llava_arch.py:175-312
#     VCoder: vcoder_llava/model/vcoder_llava_arch.py:182-294
# A plan is a list of segments per sample: ("text", [ids]) | ("image", i) | ("seg", i) | ("depth", i)
# ---------------------------------------------------------------------------------------------------------
class SpliceError(Exception):
    pass


V_VCODER, V_DS, V_LLAVA, V_IT = 0, 1, 2, 3  # == VCB_VARIANT_* (include/vcoder_b200.h)


def splice_plan(ids: Sequence[Sequence[int]], variant_ds: bool, has_seg: bool, depth_nonzero: Optional[Sequence[bool]],
                variant: Optional[int] = None, seg_nonzero: Optional[Sequence[bool]] = None
                ) -> List[List[Tuple[str, object]]]:
    """depth_nonzero: None when depths is None, else per-sample `mean(depth) != 0` (is_depth_zero, :161).
    variant (overrides variant_ds): V_VCODER vcoder_llava_arch.py:182-294 | V_DS vcoder_ds_llava_arch.py:175-312 |
    V_LLAVA llava_arch.py:117-199 | V_IT vcoder_it_llava_arch.py:165-277 (seg_nonzero: per-sample `mean(seg) != 0`,
    is_seg_zero :148; None when seg_images is None, which the reference treats as all-zero :158)."""
    v = variant if variant is not None else (V_DS if variant_ds else V_VCODER)
    plans = []
    cur_image_idx = cur_seg_idx = cur_depth_idx = 0
    for row in ids:
        cur = list(int(t) for t in row)
        n_img = sum(1 for t in cur if t == IMAGE_TOKEN_INDEX)
        n_seg = sum(1 for t in cur if t == SEG_TOKEN_INDEX)
        if v == V_DS:
            dummy = n_img == 0 and n_seg == 0  # ds:181
        elif v == V_VCODER:
            dummy = n_img == 0 or n_seg == 0  # vcoder:187
        else:
            dummy = n_img == 0  # llava_arch.py:118, vcoder_it_llava_arch.py:169
        if dummy:
            # text-only branch: embeds every id as is; a negative special id hits nn.Embedding -> IndexError (F11)
            if any(t < 0 for t in cur):
                raise SpliceError("special token id would index the embedding table (reference: IndexError)")
            half = len(cur) // 2
            plans.append([("text", cur[:half]), ("text", cur[half:])])
            cur_image_idx += 1
            cur_seg_idx += 1
            cur_depth_idx += 1
            continue
        segs: List[Tuple[str, object]] = []
        while IMAGE_TOKEN_INDEX in cur:
            i = cur.index(IMAGE_TOKEN_INDEX)
            segs.append(("text", cur[:i]))
            segs.append(("image", cur_image_idx))
            cur_image_idx += 1
            cur = cur[i + 1:]
        if v in (V_VCODER, V_DS) and has_seg:
            while SEG_TOKEN_INDEX in cur:
                i = cur.index(SEG_TOKEN_INDEX)
                if v == V_VCODER:
                    segs.append(("text", cur[:i]))  # vcoder_llava_arch.py:236
                # DS drops the ids before the seg token (vcoder_ds_llava_arch.py:233-244)
                segs.append(("seg", cur_seg_idx))
                cur_seg_idx += 1
                cur = cur[i + 1:]
        elif v == V_IT:
            seg_zero = not (has_seg and seg_nonzero is not None and bool(seg_nonzero[cur_seg_idx]))  # :148, :158
            if not seg_zero:
                if SEG_TOKEN_INDEX not in cur:
                    cur_seg_idx += 1  # :212-215 (appends an empty slice of the seg features)
                while SEG_TOKEN_INDEX in cur:
                    i = cur.index(SEG_TOKEN_INDEX)
                    segs.append(("seg", cur_seg_idx))  # ids before the seg token are dropped, as in DS (:216-226)
                    cur_seg_idx += 1
                    cur = cur[i + 1:]
            else:
                cur_seg_idx += 1  # :227-228; a remaining <seg> id then reaches embed_tokens -> IndexError
        if v == V_DS:
            nz = bool(depth_nonzero[cur_depth_idx]) if depth_nonzero is not None else False
            if nz:
                while DEPTH_TOKEN_INDEX in cur:
                    i = cur.index(DEPTH_TOKEN_INDEX)
                    segs.append(("text", cur[:i]))
                    segs.append(("depth", cur_depth_idx))
                    cur_depth_idx += 1
                    cur = cur[i + 1:]
            else:
                cur_depth_idx += 1
        if len(cur) > 0:
            segs.append(("text", cur))
        for kind, payload in segs:
            if kind == "text" and any(t < 0 for t in payload):
                raise SpliceError("special token id would index the embedding table (reference: IndexError)")
        plans.append(segs)
    return plans


def splice_embeds(ids, sd, image_feats, seg_feats, depth_feats, variant_ds: bool,
                  depth_nonzero: Optional[Sequence[bool]], variant: Optional[int] = None,
                  seg_nonzero: Optional[Sequence[bool]] = None) -> torch.Tensor:
    """-> inputs_embeds [B, T', H] fp32.  Ragged results raise (the reference crashes with NameError on the
    attention-mask rebuild when labels is None: vcoder_ds_llava_arch.py:292-300)."""
    emb = _f(sd["model.embed_tokens.weight"])  # vcoder_lm_emb is a clone of embed_tokens (:173)
    plans = splice_plan(ids, variant_ds, seg_feats is not None, depth_nonzero, variant, seg_nonzero)
    rows = []
    for segs in plans:
        parts = []
        for kind, payload in segs:
            if kind == "text":
                parts.append(emb[torch.tensor(payload, dtype=torch.long)] if len(payload) else emb[:0])
            elif kind == "image":
                parts.append(_f(image_feats[payload]))
            elif kind == "seg":
                parts.append(_f(seg_feats[payload]))
            else:
                parts.append(_f(depth_feats[payload]))
        rows.append(torch.cat(parts, dim=0))
    if any(r.shape != rows[0].shape for r in rows):
        raise SpliceError("ragged spliced batch (reference raises NameError: _new_labels)")
    return torch.stack(rows, dim=0)


# ---------------------------------------------------------------------------------------------------------
# a5: VCoder{,DS}LlavaLlamaForCausalLM.forward -> HF LlamaModel + lm_head
#     (language_model/vcoder_ds_llava_llama.py:57-118; arithmetic transformers/models/llama/modeling_llama.py:
#      RMSNorm :53-70, RoPE :124-168, MLP :171-184, eager attention :199-222)
# ---------------------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rope_cos_sin(positions: torch.Tensor, dh: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    inv_freq = 1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
    freqs = positions.float()[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


class KVCache:
    """Per-layer list of (k, v) [B, heads, T, dh] grown by concatenation (HF 4.31 tuple cache semantics)."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * n_layers
        self.v: List[Optional[torch.Tensor]] = [None] * n_layers

    @property
    def length(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[2]


def llama_layer(x: torch.Tensor, sd, i: int, g, cache: KVCache, pos0: int, return_parts: bool = False):
    B, T, H = x.shape
    heads, dh = g.llm_heads, g.llm_hidden // g.llm_heads
    p = f"model.layers.{i}."
    r = x
    a = rms_norm(x, _f(sd[p + "input_layernorm.weight"]), g.llm_rms_eps)
    q = F.linear(a, _f(sd[p + "self_attn.q_proj.weight"])).view(B, T, heads, dh).transpose(1, 2)
    k = F.linear(a, _f(sd[p + "self_attn.k_proj.weight"])).view(B, T, heads, dh).transpose(1, 2)
    v = F.linear(a, _f(sd[p + "self_attn.v_proj.weight"])).view(B, T, heads, dh).transpose(1, 2)
    # no position_ids are passed by prepare_inputs_for_generation (vcoder_ds_llava_llama.py:120-142):
    # positions = arange(past, past + T)
    cos, sin = rope_cos_sin(torch.arange(pos0, pos0 + T), dh, g.llm_rope_theta)
    q = q * cos[None, None] + rotate_half(q) * sin[None, None]
    k = k * cos[None, None] + rotate_half(k) * sin[None, None]
    if cache.k[i] is not None:
        k = torch.cat([cache.k[i], k], dim=2)
        v = torch.cat([cache.v[i], v], dim=2)
    cache.k[i], cache.v[i] = k, v
    S = k.shape[2]
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    # causal mask; padding mask is all-ones on this path (decode substitutes ones(B, past+1): arch.py:130-133)
    qpos = torch.arange(pos0, pos0 + T)[:, None]
    kpos = torch.arange(S)[None, :]
    scores = scores.masked_fill(kpos > qpos, float("-inf"))
    att = torch.softmax(scores, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, T, H)
    x = r + F.linear(o, _f(sd[p + "self_attn.o_proj.weight"]))
    r2 = x
    m = rms_norm(x, _f(sd[p + "post_attention_layernorm.weight"]), g.llm_rms_eps)
    gate = F.linear(m, _f(sd[p + "mlp.gate_proj.weight"]))
    up = F.linear(m, _f(sd[p + "mlp.up_proj.weight"]))
    x = r2 + F.linear(F.silu(gate) * up, _f(sd[p + "mlp.down_proj.weight"]))
    return x


def llama_forward(x: torch.Tensor, sd, g, cache: KVCache, last_only: bool = True,
                  collect_hidden: bool = False):
    """x: [B,T,H] input embeddings -> logits [B, V] (last position) or [B,T,V]."""
    pos0 = cache.length
    hidden = []
    for i in range(g.llm_layers):
        x = llama_layer(x, sd, i, g, cache, pos0)
        if collect_hidden:
            hidden.append(x)
    x = rms_norm(x, _f(sd["model.norm.weight"]), g.llm_rms_eps)
    if last_only:
        x = x[:, -1]
    logits = F.linear(x, _f(sd["lm_head.weight"]))
    return (logits, hidden) if collect_hidden else logits


# ---------------------------------------------------------------------------------------------------------
# a6: prepare_inputs_for_generation + HF greedy loop (vcoder_ds_llava_llama.py:120-142; transformers
#     GenerationMixin greedy: next = argmax(logits[:, -1]))
# ---------------------------------------------------------------------------------------------------------
def prefill_embeds(ids, sd, g, images, segs, depths):
    variant_ds = g.variant == V_DS
    img_f = encode_images(images, sd, g)
    seg_f = encode_seg_images(segs, sd, g) if (segs is not None and g.variant != V_LLAVA) else None
    dep_f, nz, snz = None, None, None
    if variant_ds and depths is not None:
        nz = [bool(_f(d).mean() != 0) for d in depths]  # is_depth_zero (vcoder_ds_llava_arch.py:161)
        dep_f = encode_depth_images(depths, sd, g)
    if g.variant == V_IT and segs is not None:
        snz = [bool(_f(x).mean() != 0) for x in segs]  # is_seg_zero (vcoder_it_llava_arch.py:148)
    return splice_embeds(ids, sd, img_f, seg_f, dep_f, variant_ds, nz, g.variant, snz)


def greedy_generate(ids, sd, g, images, segs, depths, max_new_tokens: int, return_logits: bool = False,
                    forced_ids: Optional[torch.Tensor] = None):
    """-> new token ids [B, N] (int64) (+ per-step logits [N][B,V]).

    forced_ids [B, N] (test infrastructure, not reference behaviour): teacher forcing — the token FED BACK at step s is
    forced_ids[:, s] instead of this function's own argmax (which is still what is returned), so the logits of a long
    run stay comparable with an implementation whose near-tie choices legitimately differ."""
    x = prefill_embeds(ids, sd, g, images, segs, depths)
    cache = KVCache(g.llm_layers)
    emb = _f(sd["model.embed_tokens.weight"])
    out, trace = [], []
    logits = llama_forward(x, sd, g, cache)
    for step in range(max_new_tokens):
        if return_logits:
            trace.append(logits)
        nxt = logits.argmax(dim=-1)
        out.append(nxt)
        if step + 1 < max_new_tokens:
            fed = nxt if forced_ids is None else forced_ids[:, step].to(torch.int64)
            logits = llama_forward(emb[fed][:, None, :], sd, g, cache)
    ids_out = torch.stack(out, dim=1)
    return (ids_out, trace) if return_logits else ids_out


# ---------------------------------------------------------------------------------------------------------
# f1: do_sample=True — HF GenerationMixin.sample() as reached from serve/cli.py:121-132 / serve/chat.py:141-151
# (third-party: transformers/generation/logits_process.py Temperature / TopK / TopP warpers, generation/utils.py).
# tests/test_oracle_golden.py pins sample_filter against the installed transformers warper classes.
# ---------------------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def philox_uniform(seed: int, step: int, row: int) -> float:
    """u in [0,1): top 24 bits of word 0 of Philox4x32-10 (Salmon et al., SC'11), key = seed (lo, hi), counter =
    (step, row, 0, 0).  torch.multinomial's own stream cannot be reproduced across devices, so the product draws
    from the SAME filtered distribution by inverse CDF at this u (include/vcoder_b200.h: vcb_philox_uniform)."""
    k0, k1 = seed & _M32, (seed >> 32) & _M32
    x0, x1, x2, x3 = step & _M32, row & _M32, 0, 0
    for _ in range(10):
        p0, p1 = 0xD2511F53 * x0, 0xCD9E8D57 * x2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & _M32, p1 >> 32, p1 & _M32
        x0, x1, x2, x3 = hi1 ^ x1 ^ k0, lo1, hi0 ^ x3 ^ k1, lo0
        k0, k1 = (k0 + 0x9E3779B9) & _M32, (k1 + 0xBB67AE85) & _M32
    return (x0 >> 8) / 16777216.0


def sample_filter(logits: torch.Tensor, temperature: float, top_k: int, top_p: float):
    """One row of logits [V] -> (order, cdf): `order` = indices of the surviving tokens by descending score (index
    ascending among equal scores — one valid outcome of torch.sort's unstable order), `cdf` = their inclusive
    cumulative probability (float64) after TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper -> softmax."""
    s = logits.detach().to(torch.float64) / float(temperature)
    V = s.numel()
    keep = torch.ones(V, dtype=torch.bool)
    if top_k and top_k > 0:
        kth = torch.topk(s, min(top_k, V)).values[-1]
        keep &= ~(s < kth)  # TopKLogitsWarper: scores < k-th largest are removed, ties stay
    idx = torch.arange(V)
    order = sorted(idx[keep].tolist(), key=lambda i: (-float(s[i]), i))
    order = torch.tensor(order, dtype=torch.int64)
    p = torch.softmax(s[order], dim=0)
    if top_p is not None and top_p < 1.0:
        # TopPLogitsWarper on the ascending sort: removed while cumsum_asc <= 1 - top_p  <=>  kept while the mass of the
        # higher-ranked tokens is < top_p; the top token always stays (min_tokens_to_keep = 1)
        before = torch.cumsum(p, 0) - p
        n = max(1, int((before < top_p).sum()))
        order, p = order[:n], p[:n]
        p = p / p.sum()
    return order, torch.cumsum(p, 0)


def sample_token(logits: torch.Tensor, temperature: float, top_k: int, top_p: float, u: float):
    """-> (token, margin): inverse-CDF draw at u; margin = distance of u from the nearest CDF edge (tests skip exact
    token equality when it is below the fp32 round-off of the device scan)."""
    order, cdf = sample_filter(logits, temperature, top_k, top_p)
    pick = int((cdf <= u).sum())
    pick = min(pick, order.numel() - 1)
    edges = torch.cat([torch.zeros(1, dtype=cdf.dtype), cdf])
    margin = float((edges - u).abs().min())
    return int(order[pick]), margin


def sample_generate(ids, sd, g, images, segs, depths, max_new_tokens: int, temperature: float, top_k: int,
                    top_p: float, seed: int, return_logits: bool = False):
    """greedy_generate with the argmax replaced by sample_token at u = philox_uniform(seed, step, row)."""
    x = prefill_embeds(ids, sd, g, images, segs, depths)
    cache = KVCache(g.llm_layers)
    emb = _f(sd["model.embed_tokens.weight"])
    out, trace, margins = [], [], []
    logits = llama_forward(x, sd, g, cache)
    for step in range(max_new_tokens):
        if return_logits:
            trace.append(logits)
        picks = [sample_token(logits[b], temperature, top_k, top_p, philox_uniform(seed, step, b))
                 for b in range(logits.shape[0])]
        nxt = torch.tensor([p[0] for p in picks], dtype=torch.int64)
        margins.append([p[1] for p in picks])
        out.append(nxt)
        if step + 1 < max_new_tokens:
            logits = llama_forward(emb[nxt][:, None, :], sd, g, cache)
    ids_out = torch.stack(out, dim=1)
    return (ids_out, trace, margins) if return_logits else ids_out
