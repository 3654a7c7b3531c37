"""Host-side mirror of the reference's model interface for the hot path — same names, argument meaning and error
behaviour as the reference classes, with every computation dispatched to libvcoder_b200.so:

    CLIPVisionTower                     vcoder_llava/model/multimodal_encoder/clip_encoder.py:7-77
    VCoderLlavaLlamaForCausalLM         vcoder_llava/model/language_model/vcoder_llava_llama.py:41-139
    VCoderDSLlavaLlamaForCausalLM       vcoder_llava/model/language_model/vcoder_ds_llava_llama.py:41-142
    encode_images / encode_seg_images / encode_depth_images   vcoder_llava/model/vcoder_ds_llava_arch.py:106-119

so that callers written against the reference (`model.generate(input_ids, images=, segs=, depths=, do_sample=False,
max_new_tokens=..., use_cache=True, streamer=, stopping_criteria=)`, `model.get_vision_tower()`, `model.config`,
`model.device`: serve/cli.py:121-132, eval/model_depth_loader.py:146-168, eval/model_seg_loader.py:127-149) keep working.
"""
from __future__ import annotations

import types
from typing import Dict, List, Optional

import torch

from . import synth
from ._lib import VcbError
from .engine import Engine
from .synth import ModelGeometry


class CLIPVisionTower:
    """Mirror of clip_encoder.py:7 — one shared, frozen tower used for RGB, seg and depth."""

    def __init__(self, engine: Engine, g: ModelGeometry, image_processor=None):
        self._eng = engine
        self._g = g
        self.is_loaded = True
        self.select_layer = g.vit_select_layer
        self.select_feature = "patch"
        self.image_processor = image_processor
        self.config = types.SimpleNamespace(hidden_size=g.vit_hidden, image_size=g.vit_image_size,
                                            patch_size=g.vit_patch, num_hidden_layers=g.vit_layers,
                                            num_attention_heads=g.vit_heads, intermediate_size=g.vit_ffn)

    def load_model(self):  # weights are bound into the engine at construction; kept for interface parity (:22-27)
        self.is_loaded = True

    @torch.no_grad()
    def forward(self, images):
        """clip_encoder.py:39-51: tensor [B,3,S,S] or a list of [3,S,S] -> hidden_states[select_layer][:, 1:]."""
        if type(images) is list:
            return [self._eng.tower_forward(im.unsqueeze(0)).to(im.dtype) for im in images]
        return self._eng.tower_forward(images).to(images.dtype)

    __call__ = forward

    @property
    def dtype(self):
        return torch.float16

    @property
    def device(self):
        return self._eng.tdev

    @property
    def hidden_size(self):
        return self._g.vit_hidden

    @property
    def num_patches(self):
        return self._g.n_patches

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)


class _VCoderBase:
    variant = synth.VARIANT_VCODER_DS
    model_type = "vcoder_ds_llava"

    def __init__(self, g: ModelGeometry, state_dict: Optional[Dict[str, torch.Tensor]] = None, max_batch: int = 8,
                 max_seq_len: int = 2048, device: int = 0, image_processor=None, load_4bit: bool = False):
        g = g.with_variant(self.variant)
        self.geometry = g
        self.load_4bit = load_4bit  # builder.py:30-37: NF4 + double quantisation, compute dtype fp16 (vcoder_b200/nf4.py)
        self._eng = Engine(g, max_batch=max_batch, max_seq_len=max_seq_len, device=device)
        self._tower = CLIPVisionTower(self._eng, g, image_processor)
        self.config = types.SimpleNamespace(
            model_type=self.model_type, hidden_size=g.llm_hidden, intermediate_size=g.llm_ffn,
            num_hidden_layers=g.llm_layers, num_attention_heads=g.llm_heads, vocab_size=g.llm_vocab,
            rms_norm_eps=g.llm_rms_eps, max_position_embeddings=max_seq_len, mm_vision_select_layer=g.vit_select_layer,
            mm_vision_select_feature="patch", mm_projector_type="mlp2x_gelu", seg_mm_projector_type="mlp2x_gelu",
            mm_hidden_size=g.vit_hidden, use_cache=True, pad_token_id=0, bos_token_id=1, eos_token_id=2)
        # GenerationConfig defaults the callers rely on implicitly: top_k = 50 applies whenever do_sample=True;
        # eos/pad come from the checkpoint (builder.load_pretrained_model fills them in), None = never stop early
        self.generation_config = types.SimpleNamespace(eos_token_id=None, pad_token_id=None, top_k=50)
        self.stream_lookahead = 4  # decode steps kept queued on the GPU while Python handles a token
        # f3: interactive callers (serve/cli.py loop, serve/chat.py:41-159) re-send the whole conversation and every image
        # each turn; with the prefix cache the towers are skipped for unchanged pixels and only the new part of the prompt
        # is prefilled.  `last_prefill` reports what the most recent interactive generate() reused.
        self.ragged_batches = True  # f3: batches of different spliced lengths are split into equal-length groups (see _generate_ragged)
        self.prefix_cache = True
        self.last_prefill = {"spliced_len": 0, "reused_len": 0, "towers_skipped": False}
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ---- nn.Module-ish surface the callers touch (builder.py:139-153) -----------------------------------------
    def load_state_dict(self, state_dict, strict: bool = False):
        if self.load_4bit:
            self._eng.bind(state_dict, nf4=True)
        else:
            self._eng.bind(state_dict)
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    def to(self, *a, **k):
        return self

    def cuda(self, *a, **k):
        return self

    def half(self):
        return self

    @property
    def device(self):
        return self._eng.tdev

    @property
    def dtype(self):
        return torch.float16

    def get_model(self):
        return self

    def get_vision_tower(self):
        return self._tower

    # ---- f3 remainder: batches whose spliced lengths differ -----------------------------------------------------------
    def _generate_ragged(self, input_ids, images, segs, depths, max_new_tokens):
        """The reference pads such a batch (vcoder_ds_llava_arch.py:278-300) and then dies with NameError (`_new_labels`
        is unset when labels is None).  Here the rows are grouped by spliced length (host planner, one row at a time) and
        every group runs through the engine as its own batch — each sequence gets exactly the tokens it would get alone;
        `ragged_batches = False` restores the reference's NameError."""
        from .engine import splice_plan
        P = self.geometry.n_patches
        ids_h = input_ids.detach().cpu().long()
        B = ids_h.shape[0]

        def nz(t, b):  # `mean(x) == 0` on the fp16 image, as is_depth_zero / is_seg_zero (arch.py:161, it arch.py:148)
            return bool(t[b].to(torch.float16).float().mean().to(torch.float16) != 0)

        groups = {}
        for b in range(B):
            flag = None
            if self.variant == synth.VARIANT_VCODER_DS and depths is not None:
                flag = [nz(depths, b)]
            elif self.variant == synth.VARIANT_VCODER_IT and segs is not None:
                flag = [nz(segs, b)]
            plan = splice_plan(self.variant, ids_h[b:b + 1].numpy(), P, segs is not None, flag)
            groups.setdefault(plan.shape[1], []).append(b)
        out = torch.zeros(B, max_new_tokens, dtype=torch.int64)
        for _, rows in sorted(groups.items()):
            idx = torch.tensor(rows)
            pick = lambda t: None if t is None else (torch.stack([t[r] for r in rows]) if isinstance(t, (list, tuple)) else t[idx.to(t.device)])
            out[idx] = self._eng.generate(ids_h[idx], pick(images), pick(segs), pick(depths), max_new_tokens=max_new_tokens).cpu()
        return out

    # ---- a2: encoders -------------------------------------------------------------------------------------------
    def encode_images(self, images):
        return self._eng.encode(0, images)

    def encode_seg_images(self, seg_images):
        return self._eng.encode(1, seg_images)

    # ---- a4/a5 as callable seams: the two methods a maintainer would monkey-patch ------------
    @torch.no_grad()
    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, past_key_values, labels, images,
                                             seg_images=None, depth_images=None):
        """Same 5-tuple as vcoder_ds_llava_arch.py:126-314 for the inference callers (labels = None, no KV yet):
        `(None, attention_mask, past_key_values, new_input_embeds, None)`; the text-only / single-token early return
        (:130-133) hands the arguments back unchanged.  Towers, projectors and the splice run on the device."""
        if labels is not None:
            raise NotImplementedError("labels (training) are outside the inference path this engine replaces")
        if images is None or input_ids.shape[1] == 1:
            if past_key_values is not None and images is not None and input_ids.shape[1] == 1 and attention_mask is not None:
                raise NotImplementedError("single-token steps run inside the engine (generate / Engine.decode)")
            return input_ids, attention_mask, past_key_values, None, labels
        if past_key_values is not None:
            raise NotImplementedError("the engine owns the KV cache (multi-turn prompts: generate() with model.prefix_cache = True)")
        if self.variant != synth.VARIANT_VCODER_DS and depth_images is not None:
            raise TypeError("only VCoder-DS takes depth images")
        embeds = self._eng.splice_embeds(input_ids, images, seg_images, depth_images).to(input_ids.device)
        if attention_mask is not None:  # :308-311: the mask grows on the LEFT by the number of spliced-in positions
            am = torch.as_tensor(attention_mask)
            pad = torch.full((am.shape[0], embeds.shape[1] - input_ids.shape[1]), True, dtype=am.dtype, device=am.device)
            attention_mask = torch.cat((pad, am), dim=1)
            assert attention_mask.shape == embeds.shape[:2]
        return None, attention_mask, past_key_values, embeds, None

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, inputs_embeds=None, labels=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, images=None, segs=None, depths=None,
                return_dict=None):
        """forward() of language_model/vcoder_ds_llava_llama.py:57-118 for a whole prompt: logits of every spliced
        position [B, T', V].  `past_key_values` of the result is None: the KV cache lives in the engine (filled by this
        call; `Engine.decode` continues from it).  Training arguments (labels, inputs_embeds) are not served."""
        if labels is not None or inputs_embeds is not None:
            raise NotImplementedError("labels / inputs_embeds (training path) are outside the inference path")
        if past_key_values is not None:
            raise NotImplementedError("the engine owns the KV cache: use generate() / Engine.decode for the next tokens")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are not materialised by the fused kernels")
        if input_ids is None:
            raise ValueError("forward() needs input_ids")
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise NotImplementedError("padded prompts: the reference itself mis-aligns the mask after the splice")
        if self.variant != synth.VARIANT_VCODER_DS and depths is not None:
            raise TypeError(f"{type(self).__name__}.forward() got an unexpected keyword argument 'depths'")
        logits = self._eng.forward_logits(input_ids, images, segs, depths).to(input_ids.device)
        if return_dict is False:
            return (logits,)
        from transformers.modeling_outputs import CausalLMOutputWithPast
        return CausalLMOutputWithPast(loss=None, logits=logits, past_key_values=None, hidden_states=None, attentions=None)

    __call__ = forward

    # ---- a5/a6 + f1: generate ----------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids=None, images=None, segs=None, depths=None, do_sample: bool = False,
                 temperature: float = 1.0, top_p: Optional[float] = None, top_k: Optional[int] = None,
                 num_beams: int = 1, max_new_tokens: int = 20, use_cache: bool = True, streamer=None,
                 stopping_criteria=None, attention_mask=None, eos_token_id=None, pad_token_id=None, seed=None,
                 inputs=None, **kwargs):
        """`GenerationMixin.generate` semantics for the two modes the callers use: greedy
        (eval/model_*_loader.py with --temperature 0) and do_sample=True with temperature / top_p (serve/cli.py:121-132,
        serve/chat.py:141-151; top_k = 50 is HF's GenerationConfig default).  Returns cat([input_ids, new_ids]) on
        input_ids' device; callers slice `[:, input_ids.shape[1]:]` (eval/model_depth_loader.py:170-174).

        Sampling runs on the device (csrc/sample.cu).  `seed` fixes the Philox stream; None draws one from torch's
        global generator, so `torch.manual_seed` makes a run reproducible as it does for the reference."""
        if input_ids is None:
            input_ids = inputs  # serve/chat.py:141 passes `inputs=`
        if input_ids is None:
            raise ValueError("generate() needs input_ids")
        if num_beams != 1:
            raise NotImplementedError("beam search is not on the reference's hot path (num_beams=1 everywhere)")
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise NotImplementedError("padded prompts: the reference itself mis-aligns the mask after the splice")
        if self.variant != synth.VARIANT_VCODER_DS and depths is not None:
            raise TypeError(f"{type(self).__name__}.generate() got an unexpected keyword argument 'depths'")
        if self.variant == synth.VARIANT_LLAVA and segs is not None:
            raise TypeError(f"{type(self).__name__}.generate() got an unexpected keyword argument 'segs'")
        gc = self.generation_config
        if eos_token_id is None:
            eos_token_id = gc.eos_token_id
        if pad_token_id is None:
            pad_token_id = gc.pad_token_id if gc.pad_token_id is not None else eos_token_id
        if do_sample:
            # the argument checks of HF's warpers (logits_process.py), same exception type
            if not (isinstance(temperature, (int, float)) and temperature > 0):
                raise ValueError(f"`temperature` (={temperature}) has to be a strictly positive float")
            tp = 1.0 if top_p is None else float(top_p)
            if not 0.0 <= tp <= 1.0:
                raise ValueError(f"`top_p` has to be a float > 0 and < 1, but is {top_p}")
            tk = gc.top_k if top_k is None else int(top_k)
            if tk < 0:
                raise ValueError(f"`top_k` has to be a non-negative integer, but is {top_k}")
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self._eng.set_sampling(True, float(temperature), tk, tp, seed)
        else:
            self._eng.set_sampling(False)
        dev = input_ids.device
        B = input_ids.shape[0]
        # capacity: generation ends at the KV-cache capacity like HF's max_length ends it, instead of failing up front.
        # Upper bound of the spliced length: every special token replaced by n_patches rows.
        n_special = int((input_ids < 0).sum(dim=1).max()) if input_ids.numel() else 0
        room = self._eng.max_seq_len - (input_ids.shape[1] + n_special * (self.geometry.n_patches - 1))
        if room < 1:
            raise ValueError(f"Input length of input_ids is {input_ids.shape[1]} ({n_special} special tokens x "
                             f"{self.geometry.n_patches} patches), but the engine holds {self._eng.max_seq_len} positions")
        max_new_tokens = min(int(max_new_tokens), room)
        interactive = streamer is not None or bool(stopping_criteria) or eos_token_id is not None
        try:
            if not interactive:
                def run(rows):  # rows = None: the whole batch
                    pick = lambda t: t if (t is None or rows is None) else t[rows.start:rows.stop]
                    ids_r, im, sg, dp = pick(input_ids), pick(images), pick(segs), pick(depths)
                    try:
                        return self._eng.generate(ids_r, im, sg, dp, max_new_tokens=max_new_tokens)
                    except VcbError as e:
                        if e.code != -5 or not self.ragged_batches or ids_r.shape[0] < 2:
                            raise
                        return self._generate_ragged(ids_r, im, sg, dp, max_new_tokens)
                mb = self._eng.max_batch
                if B <= mb:
                    new = run(None)
                else:  # the reference takes any batch: more rows than the handle holds run as consecutive slices
                    new = torch.cat([run(slice(b0, min(B, b0 + mb))).cpu() for b0 in range(0, B, mb)], dim=0)
                return torch.cat([input_ids.to(dev), new.to(dev)], dim=1)
            # token-by-token: TextStreamer / KeywordsStoppingCriteria (mm_utils.py:128-151) / eos need host ids.  The
            # GPU keeps `lookahead` steps queued while Python looks at a token (vcb_decode_stream).
            eos = None if eos_token_id is None else set(eos_token_id if isinstance(eos_token_id, (list, tuple)) else [eos_token_id])
            seq = input_ids.cpu().long()
            unfinished = torch.ones(B, dtype=torch.bool)
            if streamer is not None:
                streamer.put(seq)

            def on_token(_step, ids):
                # the body of HF's sample()/greedy_search() loop after the token is chosen
                nonlocal seq, unfinished
                nxt = torch.tensor(ids, dtype=torch.long)
                if eos is not None:
                    nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_token_id))
                seq = torch.cat([seq, nxt[:, None]], dim=1)
                if streamer is not None:
                    streamer.put(nxt)
                done = False
                if eos is not None:
                    unfinished = unfinished & torch.tensor([int(t) not in eos for t in nxt.tolist()])
                    done = not bool(unfinished.any())
                if stopping_criteria and any(bool(sc(seq, None)) for sc in stopping_criteria):  # StoppingCriteriaList.__call__
                    done = True
                return done

            pre = self._eng.prefill(input_ids, images, segs, depths, reuse=self.prefix_cache)
            self.last_prefill = {k: pre[k] for k in ("spliced_len", "reused_len", "towers_skipped")}
            stop = on_token(0, pre["next_ids"].cpu().tolist())
            if not stop and max_new_tokens > 1:
                self._eng.decode_stream(max_new_tokens - 1, on_token, lookahead=self.stream_lookahead)
            if streamer is not None:
                streamer.end()
            return seq.to(dev)
        except VcbError as e:
            # same exception types the reference surfaces
            if e.code == -6:
                raise IndexError("index out of range in self") from e
            if e.code == -5:
                raise NameError("name '_new_labels' is not defined") from e
            raise


class VCoderDSLlavaLlamaForCausalLM(_VCoderBase):
    """vcoder_ds_llava_llama.py:41 — RGB + seg + depth."""
    variant = synth.VARIANT_VCODER_DS
    model_type = "vcoder_ds_llava"

    def encode_depth_images(self, depth_images):
        return self._eng.encode(2, depth_images)  # seg_mm_projector, as the reference (vcoder_ds_llava_arch.py:113)


class VCoderLlavaLlamaForCausalLM(_VCoderBase):
    """vcoder_llava_llama.py:41 — RGB + seg."""
    variant = synth.VARIANT_VCODER
    model_type = "vcoder_llava"


class VCoderITLlavaLlamaForCausalLM(_VCoderBase):
    """vcoder_it_llava_llama.py:41 — RGB + optional seg map (all-zero seg maps are skipped, vcoder_it_llava_arch.py:148)."""
    variant = synth.VARIANT_VCODER_IT
    model_type = "vcoder_it_llava"


class LlavaLlamaForCausalLM(_VCoderBase):
    """llava_llama.py:41 — plain LLaVA-1.5: RGB only (llava_arch.py:85-199)."""
    variant = synth.VARIANT_LLAVA
    model_type = "llava"

    def encode_seg_images(self, seg_images):
        raise AttributeError("'LlavaLlamaForCausalLM' object has no attribute 'encode_seg_images'")


def variant_from_name(model_name: str) -> int:
    """Class dispatch by substring of the model name, in the order load_pretrained_model tests them (builder.py:93-108)."""
    name = model_name.lower()
    if "vcoder_it_llava" in name:
        return synth.VARIANT_VCODER_IT
    if "vcoder_ds_llava" in name:
        return synth.VARIANT_VCODER_DS
    if "vcoder_llava" in name:
        return synth.VARIANT_VCODER
    if "llava" in name:
        return synth.VARIANT_LLAVA
    raise ValueError(f"{model_name}: not a LLaVA-family model (plain language models are not on the accelerated path)")


_CLASSES = {synth.VARIANT_VCODER: VCoderLlavaLlamaForCausalLM, synth.VARIANT_VCODER_DS: VCoderDSLlavaLlamaForCausalLM,
            synth.VARIANT_LLAVA: LlavaLlamaForCausalLM, synth.VARIANT_VCODER_IT: VCoderITLlavaLlamaForCausalLM}


def build_model(model_name: str, g: ModelGeometry, state_dict=None, **kw):
    return _CLASSES[variant_from_name(model_name)](g, state_dict, **kw)
