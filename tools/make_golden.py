"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF (read-only /root/reference, `vcoder_llava`) on CPU.

Only runs in the build container (the reference cannot travel to the GPU box); the fixtures it writes are
committed.  What is executed is the reference's own code path:
    CLIPVisionTower.forward            (vcoder_llava/model/multimodal_encoder/clip_encoder.py:39-51)
    encode_images/seg_images/depth_images + prepare_inputs_labels_for_multimodal
                                       (vcoder_llava/model/vcoder_ds_llava_arch.py:106-314, vcoder_llava_arch.py:131-296)
    VCoder{,DS}LlavaLlamaForCausalLM.forward / generate
                                       (language_model/vcoder_ds_llava_llama.py:57-142, vcoder_llava_llama.py:56-139)
on top of the installed transformers (5.5.0; the reference pins 4.31.0 — see SURVEY.md §8c for the deltas), with
the seeded synthetic weights of vcoder_b200.synth loaded through load_state_dict.

Greedy ids are produced by BOTH working routes of SURVEY §8(c) and asserted identical:
  (i)  model.generate(..., use_cache=False, do_sample=False)  — the reference forward+splice every step
  (ii) prepare_inputs_labels_for_multimodal -> model.model(inputs_embeds, use_cache=True) -> lm_head -> argmax loop

usage: PYTHONPATH=/root/reference python tools/make_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True

from transformers import AutoConfig, AutoModelForCausalLM  # noqa: E402

# -- compatibility shim: 'llava' model_type clashes with transformers>=4.36 -------------
_o1 = AutoConfig.register
AutoConfig.register = staticmethod(lambda mt, cfg, exist_ok=False: _o1(mt, cfg, exist_ok=True))
_o2 = AutoModelForCausalLM.register
AutoModelForCausalLM.register = classmethod(lambda cls, cc, mc, exist_ok=False: _o2(cc, mc, exist_ok=True))

from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel  # noqa: E402

from vcoder_llava.model import (LlavaLlamaForCausalLM, VCoderDSLlavaLlamaForCausalLM,  # noqa: E402
                                VCoderITLlavaLlamaForCausalLM, VCoderLlavaLlamaForCausalLM)
from vcoder_llava.model.language_model.llava_llama import LlavaConfig  # noqa: E402
from vcoder_llava.model.language_model.vcoder_it_llava_llama import VCoderITLlavaConfig  # noqa: E402
from vcoder_llava.model.language_model.vcoder_ds_llava_llama import VCoderDSLlavaConfig  # noqa: E402
from vcoder_llava.model.language_model.vcoder_llava_llama import VCoderLlavaConfig  # noqa: E402
from vcoder_llava import mm_utils  # noqa: E402

from vcoder_b200 import synth  # noqa: E402


def build_reference_model(g, sd, tower_dir):
    vcfg = CLIPVisionConfig(hidden_size=g.vit_hidden, intermediate_size=g.vit_ffn, num_hidden_layers=g.vit_layers,
                            num_attention_heads=g.vit_heads, image_size=g.vit_image_size, patch_size=g.vit_patch,
                            hidden_act="quick_gelu", layer_norm_eps=g.vit_ln_eps, projection_dim=64)
    vcfg._attn_implementation = "eager"
    CLIPVisionModel(vcfg).save_pretrained(tower_dir)
    CLIPImageProcessor(size={"shortest_edge": g.vit_image_size}, crop_size=g.vit_image_size).save_pretrained(tower_dir)
    kw = dict(hidden_size=g.llm_hidden, intermediate_size=g.llm_ffn, num_hidden_layers=g.llm_layers,
              num_attention_heads=g.llm_heads, num_key_value_heads=g.llm_heads, vocab_size=g.llm_vocab,
              rms_norm_eps=g.llm_rms_eps, max_position_embeddings=4096, pad_token_id=0, bos_token_id=1, eos_token_id=2,
              mm_vision_tower=tower_dir, mm_vision_select_layer=g.vit_select_layer, mm_vision_select_feature="patch",
              mm_projector_type="mlp2x_gelu", mm_hidden_size=g.vit_hidden, seg_mm_projector_type="mlp2x_gelu",
              seg_mm_hidden_size=g.vit_hidden, mm_vcoder_lm_emb=True, use_mm2_proj=True, attn_implementation="eager")
    if g.variant == synth.VARIANT_VCODER_DS:
        cfg = VCoderDSLlavaConfig(depth_mm_projector_type="mlp2x_gelu", depth_mm_hidden_size=g.vit_hidden, **kw)
        model = VCoderDSLlavaLlamaForCausalLM(cfg)
    elif g.variant == synth.VARIANT_LLAVA:  # plain LLaVA-1.5 (model/llava_arch.py)
        for k in ("seg_mm_projector_type", "seg_mm_hidden_size", "mm_vcoder_lm_emb", "use_mm2_proj"):
            kw.pop(k)
        model = LlavaLlamaForCausalLM(LlavaConfig(**kw))
    elif g.variant == synth.VARIANT_VCODER_IT:  # f4: VCoder-IT (model/vcoder_it_llava_arch.py)
        for k in ("mm_vcoder_lm_emb", "use_mm2_proj"):
            kw.pop(k)
        model = VCoderITLlavaLlamaForCausalLM(VCoderITLlavaConfig(**kw))
    else:
        cfg = VCoderLlavaConfig(**kw)
        model = VCoderLlavaLlamaForCausalLM(cfg)
    model = model.eval()
    model.get_vision_tower().load_model()
    model.get_vision_tower().vision_tower.config._attn_implementation = "eager"
    # fp32 copies of the fp16-valued synthetic weights
    ref_keys = set(model.state_dict().keys())
    mine = {k: v.float() for k, v in sd.items()}
    missing = sorted(k for k in ref_keys if k not in mine and "position_ids" not in k and "inv_freq" not in k)
    extra = sorted(k for k in mine if k not in ref_keys)
    assert not missing, f"synth.weight_specs misses reference keys: {missing[:8]}"
    assert not extra, f"synth.weight_specs has non-reference keys: {extra[:8]}"
    model.load_state_dict(mine, strict=False)
    return model.float()


@torch.no_grad()
def run_variant(g, out_path, batch=2, n_new=6, seed=0, zero_seg_rows=(), ids=None):
    sd = synth.make_state_dict(g, seed=seed)
    with tempfile.TemporaryDirectory() as td:
        model = build_reference_model(g, sd, td)
        if ids is None:
            ids = synth.make_prompt_ids(g, batch, seed=seed, n_before=5, n_after=4)
        images = synth.make_pixels(g, batch, 1).float()
        segs = synth.make_pixels(g, batch, 2).float() if g.variant != synth.VARIANT_LLAVA else None
        depths = synth.make_pixels(g, batch, 3).float() if g.variant == synth.VARIANT_VCODER_DS else None
        if zero_seg_rows:  # VCoder-IT is_seg_zero path (vcoder_it_llava_arch.py:148): an all-zero seg map is skipped
            for r in zero_seg_rows:
                segs[r] = 0
        tower = model.get_vision_tower()
        out = {"ids": ids.numpy(), "weights_digest": np.array(synth.state_dict_digest(sd))}
        # a1: tower + all hidden states of the first image batch
        hs = tower.vision_tower(images, output_hidden_states=True).hidden_states
        out["tower_hidden_states"] = torch.stack([h for h in hs]).numpy()  # [L+1, B, 1+P, D]
        out["tower_features"] = tower(images).numpy()
        # a2/a3
        out["image_features"] = model.encode_images(images).numpy()
        if segs is not None:
            out["seg_features"] = model.encode_seg_images(segs).numpy()
        if depths is not None:
            out["depth_features"] = model.encode_depth_images(depths).numpy()
        # a4: splice
        mask = torch.ones_like(ids)
        if depths is not None:
            _, m2, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(ids, mask, None, None, images, segs, depths)
        elif segs is None:
            _, m2, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(ids, mask, None, None, images)
        else:
            _, m2, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(ids, mask, None, None, images, segs)
        out["inputs_embeds"] = embeds.numpy()
        out["spliced_mask"] = m2.numpy()
        # a5/a6 route (i) FIRST: the reference generate() path (forward + splice each step, no cache).
        # (Calling HF's model.model(..., output_hidden_states=True) beforehand perturbs later generate() calls under
        #  transformers 5.5.0, so per-layer hidden states below are captured with plain forward hooks instead.)
        kw = dict(images=images, do_sample=False, use_cache=False, max_new_tokens=n_new,
                  pad_token_id=0, eos_token_id=None)
        if segs is not None:
            kw["segs"] = segs
        if depths is not None:
            kw["depths"] = depths
        gen = model.generate(ids, **kw)
        gen_new = gen[:, ids.shape[1]:]
        out["generate_ids"] = gen_new.numpy()
        # route (ii): manual greedy loop with KV cache
        layer_out = []
        hooks = [l.register_forward_hook(lambda m, i, o: layer_out.append((o[0] if isinstance(o, tuple) else o).detach().clone()))
                 for l in model.model.layers]
        o = model.model(inputs_embeds=embeds, use_cache=True)
        for h in hooks:
            h.remove()
        out["prefill_layer_outputs"] = torch.stack(layer_out).numpy()  # [L, B, T', H] residual stream after each layer
        logits = model.lm_head(o.last_hidden_state)
        out["prefill_logits"] = logits.numpy()
        past = o.past_key_values
        new_ids, step_logits = [], [logits[:, -1].numpy()]
        nxt = logits[:, -1].argmax(-1)
        new_ids.append(nxt)
        for _ in range(n_new - 1):
            o = model.model(input_ids=nxt[:, None], past_key_values=past, use_cache=True)
            past = o.past_key_values
            lg = model.lm_head(o.last_hidden_state)[:, -1]
            step_logits.append(lg.numpy())
            nxt = lg.argmax(-1)
            new_ids.append(nxt)
        new_ids = torch.stack(new_ids, dim=1)
        out["greedy_ids"] = new_ids.numpy()
        out["step_logits"] = np.stack(step_logits)  # [N, B, V]
        assert torch.equal(gen_new, new_ids), (gen_new, new_ids)
        np.savez_compressed(out_path, **out)
        print(out_path, {k: getattr(v, "shape", None) for k, v in out.items()})
        print("  greedy:", new_ids.tolist())


@torch.no_grad()
def multi_turn_fixture(g, out_path, n1=6, n2=5, seed=0):
    """SURVEY §8 f3 known answers: what the reference produces on the SECOND turn of a conversation, where the callers
    (serve/chat.py:41-159, the serve/cli.py loop) re-send the whole conversation — turn-1 prompt + turn-1 answer + the new
    user message — and all images.  The reference has no cache across generate() calls, so its turn 2 is a full run."""
    sd = synth.make_state_dict(g, seed=seed)
    with tempfile.TemporaryDirectory() as td:
        model = build_reference_model(g, sd, td)
        images = synth.make_pixels(g, 2, 1).float()
        segs = synth.make_pixels(g, 2, 2).float()
        depths = synth.make_pixels(g, 2, 3).float()

        def greedy(ids, n):
            gen = model.generate(ids, images=images, segs=segs, depths=depths, do_sample=False, use_cache=False,
                                 max_new_tokens=n, pad_token_id=0, eos_token_id=None)[:, ids.shape[1]:]
            mask = torch.ones_like(ids)
            _, _, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(ids, mask, None, None, images, segs, depths)
            o = model.model(inputs_embeds=embeds, use_cache=True)
            past = o.past_key_values
            lg = model.lm_head(o.last_hidden_state)[:, -1]
            logits, new = [lg.numpy()], [lg.argmax(-1)]
            for _ in range(n - 1):
                o = model.model(input_ids=new[-1][:, None], past_key_values=past, use_cache=True)
                past = o.past_key_values
                lg = model.lm_head(o.last_hidden_state)[:, -1]
                logits.append(lg.numpy())
                new.append(lg.argmax(-1))
            new = torch.stack(new, dim=1)
            assert torch.equal(gen, new), (gen, new)
            return new, np.stack(logits), embeds.shape[1]

        ids1 = synth.make_prompt_ids(g, 2, seed=seed, n_before=5, n_after=4)
        a1, l1, t1 = greedy(ids1, n1)
        follow = torch.randint(3, g.llm_vocab, (2, 7), generator=torch.Generator().manual_seed(9))
        ids2 = torch.cat([ids1, a1, follow], dim=1)
        a2, l2, t2 = greedy(ids2, n2)
        out = {"weights_digest": np.array(synth.state_dict_digest(sd)), "ids1": ids1.numpy(), "answer1": a1.numpy(),
               "logits1": l1, "spliced_len1": np.array(t1), "ids2": ids2.numpy(), "answer2": a2.numpy(), "logits2": l2,
               "spliced_len2": np.array(t2)}
        np.savez_compressed(out_path, **out)
        print(out_path, {k: getattr(v, "shape", None) for k, v in out.items()})
        print("  turn 1:", a1.tolist(), " turn 2:", a2.tolist())


def tokenizer_order_fixture(out_path):
    """F5 known-answer: special-token order emitted by the reference tokenizer helpers (mm_utils.py:43-117)."""
    class Tok:
        bos_token_id = 1

        def __call__(self, text):
            return types.SimpleNamespace(input_ids=[1] + [10 + (ord(c) % 50) for c in text])

    tok = Tok()
    prompt = "<depth>\n<seg>\n<image>\nhi"
    ds = mm_utils.tokenizer_depth_seg_token(prompt, tok, return_tensors=None)
    sg = mm_utils.tokenizer_seg_token("<seg>\n<image>\nhi", tok, return_tensors=None)
    np.savez_compressed(out_path, depth_seg_ids=np.array(ds), seg_ids=np.array(sg))
    print(out_path, ds, sg)


def preprocess_fixture(out_path):
    """f2 known answers: the reference's own `process_images` (mm_utils.py:28-40) with and without
    `image_aspect_ratio='pad'`, driving the PIL-backed CLIP image processor (the algorithm of the pinned transformers 4.31
    CLIPImageProcessor; 5.5.0 ships it as CLIPImageProcessorPil, its default is a torchvision backend that differs by
    +-1 LSB from PIL)."""
    from PIL import Image
    from transformers.models.clip import CLIPImageProcessorPil
    rng = np.random.default_rng(5)
    out = {}
    for size in (56,):  # 336-pixel outputs are checked against Pillow itself in tests/test_preprocess_host.py
        ip = CLIPImageProcessorPil(size={"shortest_edge": size}, crop_size={"height": size, "width": size})
        shapes = [(40, 90), (97, 61), (56, 56), (150, 149)]
        for k, (h, w) in enumerate(shapes):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            # smooth half of them so the resize sees structure, not only noise
            if k % 2 == 0:
                yy, xx = np.mgrid[0:h, 0:w]
                img = np.stack([(yy * 255 // h), (xx * 255 // w), ((yy + xx) * 255 // (h + w))], -1).astype(np.uint8) ^ (img & 7)
            out[f"img_{size}_{k}"] = img
            for pad in (False, True):
                cfg = types.SimpleNamespace(image_aspect_ratio="pad" if pad else None)
                ref = mm_utils.process_images([Image.fromarray(img)], ip, cfg)[0].numpy()
                out[f"out_{size}_{k}_{'pad' if pad else 'crop'}"] = ref
    np.savez_compressed(out_path, **out)
    print(out_path, sorted(out))


if __name__ == "__main__":
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    torch.manual_seed(0)
    if "--multi-turn" in sys.argv:  # f3 fixture alone (the others stay byte-identical)
        multi_turn_fixture(synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS), os.path.join(gold, "tiny_ds_turn2.npz"))
        sys.exit(0)
    if "--small" in sys.argv:  # second pinning geometry alone (the tiny fixtures stay byte-identical)
        run_variant(synth.PRESETS["small"].with_variant(synth.VARIANT_VCODER_DS), os.path.join(gold, "small_ds.npz"), batch=1, n_new=5)
        sys.exit(0)
    only_new = "--only-new" in sys.argv  # keep the committed a-row fixtures byte-identical
    if not only_new:
        run_variant(synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS), os.path.join(gold, "tiny_ds.npz"))
        run_variant(synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER), os.path.join(gold, "tiny_vcoder.npz"))
    tokenizer_order_fixture(os.path.join(gold, "tokenizer_order.npz"))
    preprocess_fixture(os.path.join(gold, "preprocess.npz"))
    # SURVEY §8 f4 variants
    run_variant(synth.PRESETS["tiny"].with_variant(synth.VARIANT_LLAVA), os.path.join(gold, "tiny_llava.npz"))
    run_variant(synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_IT), os.path.join(gold, "tiny_it.npz"))
    # VCoder-IT with no <seg> token in the prompt: row 0 has an all-zero seg map (is_seg_zero, :148/:227), row 1 a
    # non-zero one that is encoded but spliced as an empty slice (:212-215)
    g_it = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_IT)
    run_variant(g_it, os.path.join(gold, "tiny_it_noseg.npz"), zero_seg_rows=(0,),
                ids=synth.make_prompt_ids(g_it, 2, seed=0, n_before=5, n_after=4, specials=[synth.IMAGE_TOKEN_INDEX]))
    multi_turn_fixture(synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS), os.path.join(gold, "tiny_ds_turn2.npz"))
