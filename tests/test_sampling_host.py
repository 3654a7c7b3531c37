"""CPU tests for SURVEY §8 row f1 (do_sample=True): the sampling oracle is pinned to the installed transformers
warper classes (the third-party code the reference's `generate(do_sample=True, temperature=, top_p=)` dispatches to,
serve/cli.py:121-132), and the library's host-side Philox to the oracle's and to the Random123 known answer."""
import numpy as np
import pytest
import torch

from oracle import vcoder_oracle as orc
from vcoder_b200 import _lib


def test_philox_known_answer_and_library_agreement():
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    assert int(orc.philox_uniform(0, 0, 0) * 16777216) == 0x6627E8
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(500):
        seed, step, row = int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 20)), int(rng.integers(0, 32))
        assert orc.philox_uniform(seed, step, row) == lib.vcb_philox_uniform(seed, step, row)
    u = np.array([orc.philox_uniform(7, s, r) for s in range(400) for r in range(8)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03


@pytest.mark.parametrize("temperature,top_k,top_p", [(0.2, 50, 1.0), (0.2, 50, 0.7), (1.0, 50, 0.9), (0.7, 0, 0.8),
                                                     (1.5, 5, 0.3), (1.0, 0, 1.0), (0.9, 1000, 0.95)])
def test_sample_filter_matches_transformers_warpers(temperature, top_k, top_p):
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper
    g = torch.Generator().manual_seed(int(temperature * 100) + top_k)
    for trial in range(4):
        V = [320, 1000, 4096, 32000][trial]
        logits = (torch.randn(1, V, generator=g) * 2.5).half().float()
        if trial == 1:
            logits = torch.round(logits * 4) / 4  # heavy ties, including at the top-k boundary
        s = logits.clone()
        if temperature != 1.0:
            s = TemperatureLogitsWarper(temperature)(None, s)
        if top_k:
            s = TopKLogitsWarper(top_k)(None, s)
        if top_p < 1.0:
            s = TopPLogitsWarper(top_p)(None, s)
        probs = torch.softmax(s.double(), dim=-1)[0]
        hf_keep = set(torch.nonzero(torch.isfinite(s[0])).flatten().tolist())
        order, cdf = orc.sample_filter(logits[0], temperature, top_k, top_p)
        mine = set(order.tolist())
        if mine != hf_keep:
            # only tokens tied with the last survivor may differ (torch.sort is unstable; HF keeps an arbitrary subset)
            last = float(logits[0, order[-1]])
            diff = mine ^ hf_keep
            assert all(float(logits[0, i]) == last for i in diff), (trial, len(mine), len(hf_keep))
            continue
        p_sorted = probs[order]
        assert torch.allclose(torch.cumsum(p_sorted, 0), cdf, atol=1e-9)
        # descending score, index ascending among equals
        sc = logits[0, order]
        assert bool((sc[:-1] >= sc[1:]).all())
        eq = sc[:-1] == sc[1:]
        assert bool((order[:-1][eq] < order[1:][eq]).all())


def test_sample_token_is_inverse_cdf():
    logits = torch.tensor([0.0, 1.0, 3.0, 2.0, -1.0])
    order, cdf = orc.sample_filter(logits, 1.0, 0, 1.0)
    assert order.tolist() == [2, 3, 1, 0, 4]
    counts = np.zeros(5)
    n = 20000
    for i in range(n):
        tok, _ = orc.sample_token(logits, 1.0, 0, 1.0, (i + 0.5) / n)
        counts[tok] += 1
    assert np.allclose(counts / n, torch.softmax(logits, 0).numpy(), atol=2e-3)
    assert orc.sample_token(logits, 1.0, 1, 1.0, 0.999)[0] == 2  # top_k = 1 == greedy
    assert orc.sample_token(logits, 1.0, 0, 0.0, 0.999)[0] == 2  # top_p = 0 keeps min_tokens_to_keep = 1
