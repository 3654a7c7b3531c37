"""What the transformers 4.31 <-> 5.5 skew is worth, MEASURED (VERDICT r1 weak #3 / SURVEY.md section 8c): the reference's
fp16 GPU path restated with explicit rounding points in both versions' styles (`oracle/hf_fp16_rounding.py`), against each
other and against the fp32 oracle.  CPU only; numbers printed with `-s` and recorded in DESIGN.md section 4."""
import dataclasses
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hf_fp16_rounding as hf  # noqa: E402
import vcoder_oracle as orc  # noqa: E402

from vcoder_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_rms(a, b):
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-12))


def test_version_skew_is_far_below_the_fp16_noise_floor_tiny_end_to_end():
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    z = np.load(os.path.join(GOLD, "tiny_ds.npz"))
    ids = torch.from_numpy(z["ids"])
    B = ids.shape[0]
    images, segs, depths = (synth.make_pixels(g, B, s) for s in (1, 2, 3))
    exact = torch.from_numpy(z["prefill_logits"])           # written by the reference itself (fp32, transformers 5.5)
    l431 = hf.prefill_logits(ids, sd, g, images, segs, depths, "4.31")
    l55 = hf.prefill_logits(ids, sd, g, images, segs, depths, "5.5")
    skew, n431, n55 = rel_rms(l431, l55), rel_rms(l431, exact), rel_rms(l55, exact)
    dmax = float((l431 - l55).abs().max())
    print(f"\ntiny e2e logits: skew(4.31 vs 5.5) rel-RMS {skew:.2e} max {dmax:.2e}; fp16 path vs fp32 reference: {n431:.2e} / {n55:.2e}")
    assert skew <= 6e-4 and dmax <= 8e-3                     # measured 3.2e-4 / 3.9e-3: the version skew ...
    assert 5e-4 <= n431 <= 4e-3 and 5e-4 <= n55 <= 4e-3      # ... measured 1.8e-3: HF-fp16's own distance to the fp32 reference
    assert skew <= 0.35 * min(n431, n55)
    # and it does not move a greedy token on this fixture
    first = torch.from_numpy(z["greedy_ids"][:, 0])
    assert torch.equal(l431[:, -1].argmax(-1), first) and torch.equal(l55[:, -1].argmax(-1), first)
    # the tower alone: scale 0.125 is a power of two, so moving it across the matmul changes nothing but underflow cases
    t431, t55 = hf.clip_tower_forward(images, sd, g, "4.31"), hf.clip_tower_forward(images, sd, g, "5.5")
    assert rel_rms(t431, t55) <= 1e-4
    assert rel_rms(t431, torch.from_numpy(z["tower_features"])) <= 2e-3


def test_version_skew_at_llama_7b_width():
    """Two decoder layers at the real width (H = 4096, 32 heads of 128, FFN 11008), T = 48: the sqrt(128) division vs the
    128**-0.5 multiplication and nothing else differs between the styles."""
    g = dataclasses.replace(synth.PRESETS["tiny"], name="tiny-vit+7b-width", llm_hidden=4096, llm_heads=32, llm_ffn=11008,
                            llm_layers=2, llm_vocab=2048)
    sd = synth.make_state_dict(g, seed=1)
    ids = torch.randint(3, g.llm_vocab, (2, 48), generator=torch.Generator().manual_seed(2))
    x = sd["model.embed_tokens.weight"].float()[ids]
    exact = orc.llama_forward(x, sd, g, orc.KVCache(g.llm_layers), last_only=False)
    l431, l55 = hf.llama_forward(x, sd, g, "4.31"), hf.llama_forward(x, sd, g, "5.5")
    skew, noise = rel_rms(l431, l55), rel_rms(l431, exact)
    print(f"\n7B-width logits: skew rel-RMS {skew:.2e} max {float((l431 - l55).abs().max()):.2e}; fp16 path vs fp32 oracle {noise:.2e} "
          f"max {float((l431 - exact).abs().max()):.2e}")
    assert skew <= 1.2e-3 and skew <= 0.35 * noise           # measured 6.0e-4 against 5.6e-3
    # the GPU suite holds the engine to rel-RMS 2.5e-3 of the fp32 oracle at real width (tests/test_gpu_longctx.py): the
    # fused kernels round less often than HF's op-by-op fp16 path does, so they sit CLOSER to fp32 than the reference's own
    # GPU deployment; a per-op fp16 implementation could not meet that bar
    assert 2.5e-3 <= noise <= 1.2e-2
