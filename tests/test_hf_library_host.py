"""CPU check of tools/bench_hf_library.py (the "HF library path" second baseline of SURVEY §8d): HF CLIPVisionModel +
LlamaForCausalLM filled with the seeded synthetic weights, the reference's glue restated around them and the C++ splice
planner for the row plan — must reproduce the fixture the reference itself wrote (ids exactly, logits to fp32 round-off)."""
import importlib.util
import os

import numpy as np
import torch

from vcoder_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hf_library_pipeline_reproduces_the_reference_fixture(golden_dir):
    spec = importlib.util.spec_from_file_location("bench_hf_library", os.path.join(ROOT, "tools", "bench_hf_library.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    z = np.load(os.path.join(golden_dir, "tiny_ds.npz"))
    tower, llm, proj = mod.build_modules(g, "cpu", torch.float32)
    ids = torch.from_numpy(z["ids"])
    px = [synth.make_pixels(g, ids.shape[0], s).float() for s in (1, 2, 3)]
    n = z["greedy_ids"].shape[1]
    new, trace = mod.generate(g, tower, llm, proj, ids, *px, n)
    assert np.array_equal(new.numpy(), z["greedy_ids"])
    for s in range(n):
        np.testing.assert_allclose(trace[s].numpy(), z["step_logits"][s], atol=2e-4, rtol=1e-4)
