"""GPU tests of the reference-facing Python surface: load_pretrained_model (builder.py mirror) and the
VCoder{,DS}LlavaLlamaForCausalLM.generate(...) mirror with the callers' kwargs (serve/cli.py:121-132,
eval/model_depth_loader.py:146-168)."""
import json
import os

import numpy as np
import pytest
import torch

from vcoder_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _write_checkpoint(tmp, g, sd):
    from safetensors.torch import save_file
    tower = os.path.join(tmp, "tower")
    os.makedirs(tower)
    vt = "model.vision_tower.vision_tower."
    save_file({k[len(vt):]: v.contiguous() for k, v in sd.items() if k.startswith(vt)}, os.path.join(tower, "model.safetensors"))
    json.dump({"hidden_size": g.vit_hidden, "intermediate_size": g.vit_ffn, "num_hidden_layers": g.vit_layers,
               "num_attention_heads": g.vit_heads, "image_size": g.vit_image_size, "patch_size": g.vit_patch,
               "layer_norm_eps": g.vit_ln_eps}, open(os.path.join(tower, "config.json"), "w"))
    from transformers import CLIPImageProcessor  # the tower directory of a real checkpoint carries its processor config
    CLIPImageProcessor(size={"shortest_edge": g.vit_image_size}, crop_size=g.vit_image_size).save_pretrained(tower)
    save_file({k: v.contiguous() for k, v in sd.items() if not k.startswith(vt)}, os.path.join(tmp, "model.safetensors"))
    json.dump({"hidden_size": g.llm_hidden, "intermediate_size": g.llm_ffn, "num_hidden_layers": g.llm_layers,
               "num_attention_heads": g.llm_heads, "vocab_size": g.llm_vocab, "rms_norm_eps": g.llm_rms_eps,
               "mm_vision_tower": tower, "mm_vision_select_layer": -2, "mm_projector_type": "mlp2x_gelu",
               "max_sequence_length": 256}, open(os.path.join(tmp, "config.json"), "w"))


class _Streamer:
    def __init__(self):
        self.items, self.ended = [], False

    def put(self, x):
        self.items.append(torch.as_tensor(x).clone())

    def end(self):
        self.ended = True


def test_load_pretrained_model_and_generate(tmp_path):
    from vcoder_b200.builder import load_pretrained_model
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    _write_checkpoint(str(tmp_path), g, sd)
    z = np.load(os.path.join(GOLD, "tiny_ds.npz"))
    tok, model, ip, sip, dip, ctx = load_pretrained_model(str(tmp_path), None, "vcoder_ds_llava-tiny", tokenizer="stub", max_batch=2)
    assert tok == "stub" and ctx == 256
    assert type(model).__name__ == "VCoderDSLlavaLlamaForCausalLM"
    assert model.get_vision_tower().num_patches == g.n_patches and model.get_vision_tower().is_loaded
    ids = torch.from_numpy(z["ids"]).cuda()
    B = ids.shape[0]
    images, segs, depths = (synth.make_pixels(g, B, s).cuda() for s in (1, 2, 3))
    n_new = z["greedy_ids"].shape[1]
    out = model.generate(ids, images=images, segs=segs, depths=depths, do_sample=False, temperature=0, top_p=None,
                         num_beams=1, max_new_tokens=n_new, use_cache=True)
    assert out.device == ids.device and out.shape == (B, ids.shape[1] + n_new)
    assert torch.equal(out[:, : ids.shape[1]].cpu(), torch.from_numpy(z["ids"]))
    assert np.array_equal(out[:, ids.shape[1]:].cpu().numpy(), z["greedy_ids"])  # == the reference's own generate()
    # tower through the mirrored CLIPVisionTower.forward, list input as clip_encoder.py:40-45
    feats = model.get_vision_tower()([images[0], images[1]])
    assert len(feats) == 2 and feats[0].shape == (1, g.n_patches, g.vit_hidden)
    # streamer + stopping criteria (the serve/cli.py path), batch 1 like KeywordsStoppingCriteria asserts
    st = _Streamer()
    stop = [lambda seq, scores: seq.shape[1] >= ids.shape[1] + 3]
    o1 = model.generate(ids[:1], images=images[:1], segs=segs[:1], depths=depths[:1], do_sample=False, max_new_tokens=n_new,
                        streamer=st, stopping_criteria=stop)
    assert o1.shape[1] == ids.shape[1] + 3 and st.ended
    assert np.array_equal(o1[0, ids.shape[1]:].cpu().numpy(), z["greedy_ids"][0, :3])


def test_reference_error_behaviour():
    from vcoder_b200.model import VCoderDSLlavaLlamaForCausalLM, VCoderLlavaLlamaForCausalLM
    I, S = synth.IMAGE_TOKEN_INDEX, synth.SEG_TOKEN_INDEX
    g = synth.PRESETS["tiny"]
    sd = synth.make_state_dict(g.with_variant(synth.VARIANT_VCODER), seed=0)
    m = VCoderLlavaLlamaForCausalLM(g, sd, max_batch=2, max_seq_len=128)
    images, segs = synth.make_pixels(g, 1, 1), synth.make_pixels(g, 1, 2)
    with pytest.raises(IndexError):  # <image> without <seg> on the non-DS model
        m.generate(torch.tensor([[1, 5, I, 7]]), images=images, segs=segs, max_new_tokens=2)
    with pytest.raises(TypeError):
        m.generate(torch.tensor([[1, I, S, 7]]), images=images, segs=segs, depths=images, max_new_tokens=2)
    sd2 = synth.make_state_dict(g.with_variant(synth.VARIANT_VCODER_DS), seed=0)
    m2 = VCoderDSLlavaLlamaForCausalLM(g, sd2, max_batch=2, max_seq_len=128)
    im2, sg2 = synth.make_pixels(g, 2, 1), synth.make_pixels(g, 2, 2)
    # ragged batch (spliced lengths differ): the reference dies with NameError; here every row gets what it would get alone
    ragged = torch.tensor([[1, I, S, 7], [1, 5, 6, 7]])
    got = m2.generate(ragged, images=im2, segs=sg2, depths=im2, max_new_tokens=3)
    for b in range(2):
        alone = m2.generate(ragged[b:b + 1], images=im2[b:b + 1], segs=sg2[b:b + 1], depths=im2[b:b + 1], max_new_tokens=3)
        assert torch.equal(got[b:b + 1], alone)
    m2.ragged_batches = False
    with pytest.raises(NameError):  # the reference's behaviour (vcoder_ds_llava_arch.py:292-300)
        m2.generate(ragged, images=im2, segs=sg2, depths=im2, max_new_tokens=2)
    with pytest.raises(ValueError):  # HF: `temperature` has to be a strictly positive float
        m2.generate(torch.tensor([[1, I, S, 7]]), images=images, segs=segs, do_sample=True, temperature=0.0, max_new_tokens=2)
    with pytest.raises(NotImplementedError):
        m2.generate(torch.tensor([[1, I, S, 7]]), images=images, segs=segs, num_beams=2, max_new_tokens=2)


def test_f4_variant_classes_and_dispatch():
    """class dispatch by model name in the reference's order (builder.py:93-108) and the LLaVA / VCoder-IT
    shims reproducing the reference's own greedy ids (fixtures written by the reference)."""
    from vcoder_b200.model import build_model, variant_from_name
    assert variant_from_name("vcoder_it_llava-v1.5-7b") == synth.VARIANT_VCODER_IT
    assert variant_from_name("vcoder_ds_llava-v1.5-7b") == synth.VARIANT_VCODER_DS
    assert variant_from_name("vcoder_llava-v1.5-7b") == synth.VARIANT_VCODER
    assert variant_from_name("llava-v1.5-7b") == synth.VARIANT_LLAVA
    with pytest.raises(ValueError):
        variant_from_name("vicuna-7b")
    for name, variant, fname in (("llava-tiny", synth.VARIANT_LLAVA, "tiny_llava.npz"),
                                 ("vcoder_it_llava-tiny", synth.VARIANT_VCODER_IT, "tiny_it.npz")):
        g = synth.PRESETS["tiny"].with_variant(variant)
        sd = synth.make_state_dict(g, seed=0)
        z = np.load(os.path.join(GOLD, fname))
        m = build_model(name, g, sd, max_batch=2, max_seq_len=128)
        assert m.geometry.variant == variant
        ids = torch.from_numpy(z["ids"])
        images, segs = synth.make_pixels(g, 2, 1), synth.make_pixels(g, 2, 2)
        n_new = z["greedy_ids"].shape[1]
        kw = dict(images=images, max_new_tokens=n_new)
        if variant == synth.VARIANT_VCODER_IT:
            kw["segs"] = segs
        out = m.generate(ids, **kw)
        assert np.array_equal(out[:, ids.shape[1]:].cpu().numpy(), z["greedy_ids"])
        with pytest.raises(TypeError):
            m.generate(ids, images=images, segs=segs, depths=images, max_new_tokens=2)
        if variant == synth.VARIANT_LLAVA:
            with pytest.raises(TypeError):
                m.generate(ids, images=images, segs=segs, max_new_tokens=2)
