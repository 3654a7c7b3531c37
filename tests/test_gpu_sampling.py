"""GPU tests for SURVEY §8 row f1: on-device sampling (HF sample(): temperature -> top-k -> top-p -> draw) and the
per-token streamed decode, all through the C ABI, checked against the CPU oracle (oracle.sample_* — itself pinned to the
installed transformers warpers in tests/test_sampling_host.py).

Integer results (kept-token counts, chosen ids) are exact except where the oracle's float64 CDF puts the decision within
1e-5 of an edge (fp32 scan round-off) — those draws are skipped, and the test asserts they are rare."""
import os

import numpy as np
import pytest
import torch

from oracle import vcoder_oracle as orc
from vcoder_b200 import _lib, synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from vcoder_b200.engine import Engine

EDGE = 1e-5


def _tiny_engine(B=2, max_seq_len=192):
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    ids = synth.make_prompt_ids(g, B, seed=0, n_before=5, n_after=4)
    px = [synth.make_pixels(g, B, s) for s in (1, 2, 3)]
    eng = Engine(g, max_batch=B, max_seq_len=max_seq_len)
    eng.bind(sd)
    return g, sd, eng, ids, px


def _check_draw(order, cdf, tok, u, what):
    """tok must be the inverse-CDF draw at u: cdf[r-1] <= u < cdf[r] for its rank r, up to EDGE of fp32 scan round-off
    (with 32000 survivors the CDF edges are ~3e-5 apart, so an exact-rank comparison would mostly test round-off)."""
    pos = (order == tok).nonzero()
    assert pos.numel() == 1, f"{what}: token {tok} is not among the survivors"
    r = int(pos[0, 0])
    lo = float(cdf[r - 1]) if r > 0 else 0.0
    hi = float(cdf[r])
    assert lo - EDGE <= u < hi + EDGE, f"{what}: token {tok} (rank {r}) covers [{lo}, {hi}) but u = {u}"


@pytest.mark.parametrize("V,B", [(320, 4), (1000, 3), (32000, 8)])
@pytest.mark.parametrize("temperature,top_k,top_p", [(0.2, 50, 1.0), (0.2, 50, 0.7), (1.0, 50, 0.9), (0.7, 0, 0.8),
                                                     (1.0, 0, 1.0), (1.3, 7, 0.5)])
def test_sample_op_matches_oracle(V, B, temperature, top_k, top_p):
    g, sd, eng, ids, px = _tiny_engine()
    gen = torch.Generator().manual_seed(V + top_k)
    logits = (torch.randn(B, V, generator=gen) * 2.0).half()
    logits[0] = (torch.round(logits[0].float() * 2) / 2).half()  # a row full of ties
    for step in (0, 5):
        seed = 1234 + step
        nxt, kept = eng.sample_op(logits, temperature, top_k, top_p, seed, step)
        torch.cuda.synchronize()
        for b in range(B):
            order, cdf = orc.sample_filter(logits[b].float(), temperature, top_k, top_p)
            n = order.numel()
            # the number of survivors is exact unless the top-p cut falls within fp32 round-off of a CDF edge
            if int(kept[b]) != n:
                assert abs(int(kept[b]) - n) == 1 and top_p < 1.0, (b, int(kept[b]), n)
                full_order, full_cdf = orc.sample_filter(logits[b].float(), temperature, top_k, 1.0)
                edge = float(full_cdf[min(int(kept[b]), n) - 1])
                assert abs(edge - top_p) < EDGE, (b, edge, top_p)
                continue
            _check_draw(order, cdf, int(nxt[b]), orc.philox_uniform(seed, step, b), (b, step))


def test_sample_op_distribution():
    """Many steps over one row: empirical frequencies follow the filtered softmax."""
    g, sd, eng, ids, px = _tiny_engine()
    V = 320
    row = (torch.randn(1, V, generator=torch.Generator().manual_seed(3)) * 1.5).half()
    order, cdf = orc.sample_filter(row[0].float(), 0.8, 20, 0.9)
    p = torch.diff(cdf, prepend=torch.zeros(1, dtype=cdf.dtype)).numpy()
    counts = np.zeros(V)
    n = 3000
    batch = row.repeat(30, 1)
    for step in range(n // 30):
        nxt, _ = eng.sample_op(batch, 0.8, 20, 0.9, 99, step)
        for t in nxt.tolist():
            counts[t] += 1
    assert set(np.nonzero(counts)[0]) <= set(order.tolist())
    assert np.abs(counts[order.numpy()] / n - p).max() < 0.03


def test_sampled_generation_follows_oracle_and_graph_equals_eager():
    g, sd, eng, ids, px = _tiny_engine()
    T, K, P, seed, n_new = 0.7, 50, 0.9, 4321, 8
    eng.set_sampling(True, T, K, P, seed)
    pre = eng.prefill(ids, *px, want_logits=True)
    dec_ids, dec_logits = eng.decode(n_new - 1, want_logits=True)  # eager launches, logits trace
    torch.cuda.synchronize()
    got = torch.cat([pre["next_ids"][:, None], dec_ids], dim=1).cpu().long()
    lg = [pre["logits"]] + [dec_logits[s] for s in range(n_new - 1)]
    # (1) every sampled id is the oracle's draw from the DEVICE's own logits at the same (seed, step, row)
    for s in range(n_new):
        for b in range(ids.shape[0]):
            order, cdf = orc.sample_filter(lg[s][b].float().cpu(), T, K, P)
            _check_draw(order, cdf, int(got[b, s]), orc.philox_uniform(seed, s, b), (s, b))
    # (2) against the full oracle run: same ids until the first draw whose CDF margin is inside the logit error
    ref_ids, ref_logits, margins = orc.sample_generate(ids, sd, g, *px, n_new, T, K, P, seed, return_logits=True)
    for b in range(ids.shape[0]):
        for s in range(n_new):
            if int(got[b, s]) != int(ref_ids[b, s]):
                assert margins[s][b] < 5e-2, (b, s, margins[s][b])
                break
    # (3) graph path (vcb_generate) == eager path for the same seed; a different seed changes the draw
    gen = eng.generate(ids, *px, max_new_tokens=n_new)
    assert torch.equal(gen, got)
    eng.set_sampling(True, T, K, P, seed + 1)
    gen2 = eng.generate(ids, *px, max_new_tokens=n_new)
    assert not torch.equal(gen2, gen)
    # (4) back to greedy
    eng.set_sampling(False)
    a = eng.generate(ids, *px, max_new_tokens=n_new)  # greedy parity itself is covered in test_gpu_parity.py
    eng.set_sampling(True, 1.0, 0, 0.0, seed)  # top_p = 0 keeps exactly the top-ranked token == first-index argmax
    c = eng.generate(ids, *px, max_new_tokens=n_new)
    assert torch.equal(c, a)


@pytest.mark.parametrize("sampling", [False, True])
def test_decode_stream_rollback(sampling):
    g, sd, eng, ids, px = _tiny_engine()
    n_new = 24
    if sampling:
        eng.set_sampling(True, 0.9, 50, 0.95, 77)
    full = eng.generate(ids, *px, max_new_tokens=n_new)  # [B, n_new], uninterrupted graph replay
    # streamed, never stopped: identical tokens, one callback per step
    pre = eng.prefill(ids, *px)
    seen = []
    n = eng.decode_stream(n_new - 1, lambda step, t: seen.append((step, t)) and False, lookahead=4)
    assert n == n_new - 1 and [s for s, _ in seen] == list(range(n_new - 1))
    got = torch.tensor([t for _, t in seen]).T
    assert torch.equal(torch.cat([pre["next_ids"].cpu().long()[:, None], got], dim=1), full)
    # stop after 5 tokens with 8 steps in flight, then continue: the discarded steps must leave no trace
    for lookahead in (1, 8, 16):
        pre = eng.prefill(ids, *px)
        seen = []
        n = eng.decode_stream(n_new - 1, lambda step, t: (seen.append(t), step == 4)[1], lookahead=lookahead)
        assert n == 5
        rest = eng.decode(n_new - 1 - 5).cpu().long()
        seq = torch.cat([pre["next_ids"].cpu().long()[:, None], torch.tensor(seen).T, rest], dim=1)
        assert torch.equal(seq, full), lookahead
    # exception in the callback propagates and the handle stays usable after a new prefill
    pre = eng.prefill(ids, *px)
    with pytest.raises(ZeroDivisionError):
        eng.decode_stream(4, lambda step, t: 1 / 0)
    again = eng.generate(ids, *px, max_new_tokens=n_new)
    assert torch.equal(again, full)


def test_shim_sampling_streaming_eos():
    from vcoder_b200.model import VCoderDSLlavaLlamaForCausalLM
    g = synth.PRESETS["tiny"]
    sd = synth.make_state_dict(g.with_variant(synth.VARIANT_VCODER_DS), seed=0)
    m = VCoderDSLlavaLlamaForCausalLM(g, sd, max_batch=2, max_seq_len=192)
    ids = synth.make_prompt_ids(m.geometry, 2, seed=0, n_before=5, n_after=4)
    px = [synth.make_pixels(m.geometry, 2, s) for s in (1, 2, 3)]
    kw = dict(images=px[0], segs=px[1], depths=px[2], max_new_tokens=12)
    a = m.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, seed=5, **kw)
    b = m.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, seed=5, **kw)
    c = m.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, seed=6, **kw)
    assert a.shape == (2, ids.shape[1] + 12) and torch.equal(a, b) and not torch.equal(a, c)
    torch.manual_seed(11)
    d = m.generate(ids, do_sample=True, temperature=0.8, **kw)
    torch.manual_seed(11)
    e = m.generate(ids, do_sample=True, temperature=0.8, **kw)
    assert torch.equal(d, e)  # torch.manual_seed reproducibility, as for the reference
    with pytest.raises(ValueError):
        m.generate(ids, do_sample=True, temperature=0, **kw)
    with pytest.raises(ValueError):
        m.generate(ids, do_sample=True, temperature=0.5, top_p=1.5, **kw)

    class S:
        def __init__(self):
            self.got, self.ended = [], False

        def put(self, x):
            self.got.append(x.clone())

        def end(self):
            self.ended = True

    st = S()
    f = m.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, seed=5, streamer=st, **kw)
    assert torch.equal(f, a) and st.ended and len(st.got) == 1 + 12 and torch.equal(st.got[0], ids)
    assert torch.equal(torch.stack(st.got[1:], dim=1), a[:, ids.shape[1]:])
    # eos: pick the 4th sampled token of row 0 as "eos": row 0 is padded after it; generation ends when all rows finished
    new = a[:, ids.shape[1]:]
    eos = int(new[0, 3])
    h = m.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, seed=5, eos_token_id=eos, pad_token_id=0, **kw)
    hn = h[:, ids.shape[1]:]
    first = [next((i for i, t in enumerate(new[r].tolist()) if t == eos), None) for r in range(2)]
    end = 12 if any(x is None for x in first) else max(first) + 1
    assert hn.shape[1] == end
    for r in range(2):
        stop = first[r] if first[r] is not None else 12
        assert torch.equal(hn[r, : min(stop + 1, end)], new[r, : min(stop + 1, end)])
        assert bool((hn[r, stop + 1:] == 0).all())
    # a later greedy call is unaffected by the sampling state
    g1 = m.generate(ids, do_sample=False, **kw)
    g2 = m.generate(ids, **kw)
    assert torch.equal(g1, g2)
