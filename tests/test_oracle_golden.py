"""Pin the CPU oracle (oracle/vcoder_oracle.py) against outputs of the reference itself (tests/golden/*.npz,
written by tools/make_golden.py from /root/reference's vcoder_llava running on the installed transformers).
fp32 vs fp32: tolerances are float round-off only."""
import os

import numpy as np
import pytest
import torch

from oracle import vcoder_oracle as orc
from vcoder_b200 import synth

CASES = [("tiny_ds.npz", synth.VARIANT_VCODER_DS), ("tiny_vcoder.npz", synth.VARIANT_VCODER),
         # plain LLaVA-1.5 (llava_arch.py) and VCoder-IT (vcoder_it_llava_arch.py), same generator script
         ("tiny_llava.npz", synth.VARIANT_LLAVA), ("tiny_it.npz", synth.VARIANT_VCODER_IT),
         # a second geometry (synth.PRESETS["small"]: 65 ViT tokens, 4 heads, 6 + 3 layers, FFN 640 / 1408), written by the
         # reference with `tools/make_golden.py --small`: the oracle is not only pinned at the tiny shape
         ("small_ds.npz", synth.VARIANT_VCODER_DS)]


def _load(golden_dir, fname, variant):
    z = np.load(os.path.join(golden_dir, fname))
    g = synth.PRESETS[fname.split("_")[0]].with_variant(variant)  # tiny_*.npz / small_*.npz
    sd = synth.make_state_dict(g, seed=0)
    assert synth.state_dict_digest(sd) == str(z["weights_digest"]), "seeded weight generator drifted from the fixture"
    B = z["ids"].shape[0]
    images = synth.make_pixels(g, B, 1)
    segs = synth.make_pixels(g, B, 2) if variant != synth.VARIANT_LLAVA else None
    depths = synth.make_pixels(g, B, 3) if variant == synth.VARIANT_VCODER_DS else None
    return z, g, sd, images, segs, depths


def _close(a, b, atol=2e-4, rtol=1e-4):
    """fp32-vs-fp32: round-off only, scaled by the tensor's magnitude."""
    a = a.numpy() if isinstance(a, torch.Tensor) else a
    np.testing.assert_allclose(a, b, atol=atol * max(1.0, float(np.abs(b).max())), rtol=rtol)


@pytest.mark.parametrize("fname,variant", CASES)
def test_tower_hidden_states_and_select(golden_dir, fname, variant):
    z, g, sd, images, _, _ = _load(golden_dir, fname, variant)
    hs = orc.clip_hidden_states(images, sd, g)
    assert len(hs) == g.vit_layers + 1 == z["tower_hidden_states"].shape[0]
    for i, h in enumerate(hs):
        _close(h, z["tower_hidden_states"][i])
    # hidden_states[-2][:, 1:]: output of layer L-1, CLS dropped, no post-LN
    _close(orc.clip_tower_forward(images, sd, g), z["tower_features"])
    _close(hs[g.vit_layers - 1][:, 1:], z["tower_features"])


@pytest.mark.parametrize("fname,variant", CASES)
def test_encoders(golden_dir, fname, variant):
    z, g, sd, images, segs, depths = _load(golden_dir, fname, variant)
    _close(orc.encode_images(images, sd, g), z["image_features"])
    if segs is not None:
        _close(orc.encode_seg_images(segs, sd, g), z["seg_features"])
    if depths is not None:
        # depth goes through seg_mm_projector
        _close(orc.encode_depth_images(depths, sd, g), z["depth_features"])


@pytest.mark.parametrize("fname,variant", CASES)
def test_splice(golden_dir, fname, variant):
    z, g, sd, images, segs, depths = _load(golden_dir, fname, variant)
    x = orc.prefill_embeds(torch.from_numpy(z["ids"]), sd, g, images, segs, depths)
    assert tuple(x.shape) == z["inputs_embeds"].shape
    T_ids = z["ids"].shape[1]
    n_special = {synth.VARIANT_VCODER_DS: 3, synth.VARIANT_LLAVA: 1}.get(variant, 2)
    n_vis = 1 if variant == synth.VARIANT_LLAVA else 2
    assert x.shape[1] == T_ids - n_special + n_vis * g.n_patches  # depth computed then dropped
    _close(x, z["inputs_embeds"])
    assert z["spliced_mask"].all()


@pytest.mark.parametrize("fname,variant", CASES)
def test_prefill_layers_logits_and_greedy(golden_dir, fname, variant):
    z, g, sd, images, segs, depths = _load(golden_dir, fname, variant)
    x = torch.from_numpy(z["inputs_embeds"])
    cache = orc.KVCache(g.llm_layers)
    logits, hidden = orc.llama_forward(x, sd, g, cache, last_only=False, collect_hidden=True)
    for i, h in enumerate(hidden):
        _close(h, z["prefill_layer_outputs"][i], atol=5e-4)
    _close(logits, z["prefill_logits"], atol=5e-4)
    ids_new, trace = orc.greedy_generate(torch.from_numpy(z["ids"]), sd, g, images, segs, depths,
                                         max_new_tokens=z["greedy_ids"].shape[1], return_logits=True)
    assert np.array_equal(ids_new.numpy(), z["greedy_ids"])
    assert np.array_equal(ids_new.numpy(), z["generate_ids"])  # the reference generate() route
    for s, lg in enumerate(trace):
        _close(lg, z["step_logits"][s], atol=5e-4)


def test_tokenizer_special_order(golden_dir):
    """F5 known answer: the reference tokenizer helpers emit [IMAGE, DEPTH, SEG] / [IMAGE, SEG]."""
    z = np.load(os.path.join(golden_dir, "tokenizer_order.npz"))
    ds = [t for t in z["depth_seg_ids"].tolist() if t < 0]
    sg = [t for t in z["seg_ids"].tolist() if t < 0]
    assert ds == [orc.IMAGE_TOKEN_INDEX, orc.DEPTH_TOKEN_INDEX, orc.SEG_TOKEN_INDEX]
    assert sg == [orc.IMAGE_TOKEN_INDEX, orc.SEG_TOKEN_INDEX]
    assert synth.make_prompt_ids(synth.PRESETS["tiny"], 1)[0].tolist()[35:38] == ds


def test_splice_plan_quirks():
    I, S, D = orc.IMAGE_TOKEN_INDEX, orc.SEG_TOKEN_INDEX, orc.DEPTH_TOKEN_INDEX
    # DS, tokenizer order [I, D, S]: depth id silently dropped, depth features never spliced (F5)
    p = orc.splice_plan([[1, 5, I, D, S, 9]], True, True, [True])[0]
    assert p == [("text", [1, 5]), ("image", 0), ("seg", 0), ("text", [9])]
    # DS with depth AFTER seg: text between is embedded, depth spliced
    p = orc.splice_plan([[1, I, 7, S, 8, D, 9]], True, True, [True])[0]
    assert p == [("text", [1]), ("image", 0), ("seg", 0), ("text", [8]), ("depth", 0), ("text", [9])]
    # ... unless the depth image is all zero (is_depth_zero): the DEPTH id then reaches the embedding -> IndexError
    with pytest.raises(orc.SpliceError):
        orc.splice_plan([[1, I, 7, S, 8, D, 9]], True, True, [False])
    # VCoder keeps the text between image and seg
    p = orc.splice_plan([[1, I, 7, S, 9]], False, True, None)[0]
    assert p == [("text", [1]), ("image", 0), ("text", [7]), ("seg", 0), ("text", [9])]
    # VCoder with image but no seg falls into the dummy branch and indexes the embedding with -200 (F11)
    with pytest.raises(orc.SpliceError):
        orc.splice_plan([[1, I, 9]], False, True, None)
    # DS tolerates image-only
    p = orc.splice_plan([[1, I, 9]], True, True, [True])[0]
    assert p == [("text", [1]), ("image", 0), ("text", [9])]
    # text-only prompt: split in halves, both embedded
    assert orc.splice_plan([[1, 2, 3]], True, True, None)[0] == [("text", [1]), ("text", [2, 3])]


def test_vcoder_it_zero_seg_and_missing_seg_token(golden_dir):
    """VCoder-IT (f4): prompt with <image> only; row 0's seg map is all-zero (skipped, vcoder_it_llava_arch.py:148/:227),
    row 1's is encoded but contributes an empty slice (:212-215).  Fixture written by the reference itself."""
    z = np.load(os.path.join(golden_dir, "tiny_it_noseg.npz"))
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_IT)
    sd = synth.make_state_dict(g, seed=0)
    assert synth.state_dict_digest(sd) == str(z["weights_digest"])
    images, segs = synth.make_pixels(g, 2, 1), synth.make_pixels(g, 2, 2)
    segs[0] = 0
    ids = torch.from_numpy(z["ids"])
    x = orc.prefill_embeds(ids, sd, g, images, segs, None)
    assert x.shape[1] == ids.shape[1] - 1 + g.n_patches
    _close(x, z["inputs_embeds"])
    new = orc.greedy_generate(ids, sd, g, images, segs, None, max_new_tokens=z["greedy_ids"].shape[1])
    assert np.array_equal(new.numpy(), z["greedy_ids"])
    # a <seg> token behind an all-zero seg map reaches embed_tokens: the reference's IndexError
    I, S = orc.IMAGE_TOKEN_INDEX, orc.SEG_TOKEN_INDEX
    with pytest.raises(orc.SpliceError):
        orc.splice_plan([[1, I, S, 9]], False, True, None, variant=orc.V_IT, seg_nonzero=[False])
    p = orc.splice_plan([[1, I, 7, S, 9]], False, True, None, variant=orc.V_IT, seg_nonzero=[True])[0]
    assert p == [("text", [1]), ("image", 0), ("seg", 0), ("text", [9])]  # the 7 before <seg> is dropped, as in DS
    p = orc.splice_plan([[1, I, 7, 9]], False, False, None, variant=orc.V_LLAVA)[0]
    assert p == [("text", [1]), ("image", 0), ("text", [7, 9])]


def test_multi_turn_second_turn(golden_dir):
    """f3: the reference's own second turn (whole conversation re-sent: turn-1 prompt + its answer + a new message, same
    images) — the full-recompute result the prefix cache of the product must reproduce (tests/test_gpu_prefix_cache.py)."""
    z = np.load(os.path.join(golden_dir, "tiny_ds_turn2.npz"))
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    assert synth.state_dict_digest(sd) == str(z["weights_digest"])
    z1 = np.load(os.path.join(golden_dir, "tiny_ds.npz"))
    assert np.array_equal(z["ids1"], z1["ids"]) and np.array_equal(z["answer1"], z1["greedy_ids"])  # turn 1 = the a-row fixture
    assert np.array_equal(z["ids2"][:, : z["ids1"].shape[1] + z["answer1"].shape[1]], np.concatenate([z["ids1"], z["answer1"]], axis=1))
    px = [synth.make_pixels(g, 2, s) for s in (1, 2, 3)]
    n2 = z["answer2"].shape[1]
    ids_new, trace = orc.greedy_generate(torch.from_numpy(z["ids2"]), sd, g, *px, max_new_tokens=n2, return_logits=True)
    assert np.array_equal(ids_new.numpy(), z["answer2"])
    for s, lg in enumerate(trace):
        _close(lg, z["logits2"][s], atol=5e-4)
    # F5: two blocks of visual tokens, the depth block dropped; the answer and the follow-up are plain text rows
    assert int(z["spliced_len2"]) == int(z["spliced_len1"]) + z["answer1"].shape[1] + 7
