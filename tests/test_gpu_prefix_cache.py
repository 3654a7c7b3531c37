"""GPU parity tests of the multi-turn prefix cache (`vcb_prefill_cached`).

The reference has no such cache: serve/chat.py:41-159 and the serve/cli.py loop re-run the towers and re-prefill the
whole conversation every turn.  So the checker for a cached turn is simply the full computation on the same inputs:
the CPU oracle's greedy_generate on the turn's full prompt, and a FRESH engine's uncached vcb_prefill.

  logits vs the oracle: max-abs <= 1e-2 (the engine-level fp16 tolerance of tests/test_gpu_parity.py)
  logits vs an uncached prefill on the GPU: max-abs <= 1e-2 when previously GENERATED tokens are reused (their KV
      entries came from the decode kernels, a different summation order), <= 2e-3 when only prompt positions are
      reused (position-wise the same arithmetic)
  reused_len / towers_skipped / ids: exact
"""
import pytest
import torch

from oracle import vcoder_oracle as orc
from vcoder_b200 import synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from vcoder_b200.engine import Engine


def max_abs(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max())


def _setup(B=2, variant=synth.VARIANT_VCODER_DS, max_seq_len=192):
    g = synth.PRESETS["tiny"].with_variant(variant)
    sd = synth.make_state_dict(g, seed=0)
    ids = synth.make_prompt_ids(g, B, seed=5, n_before=6, n_after=5)
    px = [synth.make_pixels(g, B, s) for s in (1, 2, 3)]
    eng = Engine(g, max_batch=B, max_seq_len=max_seq_len)
    eng.bind(sd)
    return g, sd, ids, px, eng


def _fresh_logits(g, sd, ids, px, B):
    eng = Engine(g, max_batch=B, max_seq_len=192)
    eng.bind(sd)
    pre = eng.prefill(ids, *px, want_logits=True)
    torch.cuda.synchronize()
    return pre


def test_multi_turn_reuses_visual_and_generated_tokens():
    B, n1 = 2, 5
    g, sd, ids1, px, eng = _setup(B)
    # ---- turn 1: nothing cached yet
    pre1 = eng.prefill(ids1, *px, want_logits=True, reuse=True)
    assert pre1["reused_len"] == 0 and not pre1["towers_skipped"]
    Tp1 = pre1["spliced_len"]
    assert Tp1 == ids1.shape[1] - 3 + 2 * g.n_patches  # F5: depth dropped
    dec = eng.decode(n1 - 1)
    torch.cuda.synchronize()
    answer = torch.cat([pre1["next_ids"][:, None], dec], dim=1).cpu().long()  # n1 generated tokens
    ref1 = _fresh_logits(g, sd, ids1, px, B)
    assert max_abs(pre1["logits"], ref1["logits"]) == 0.0  # reuse=True with an empty cache IS the plain prefill
    # ---- turn 2: conversation so far + the answer + a new user message (same images re-sent, as chat.py does)
    follow = torch.randint(3, g.llm_vocab, (B, 7), generator=torch.Generator().manual_seed(9))
    ids2 = torch.cat([ids1, answer, follow], dim=1)
    pre2 = eng.prefill(ids2, *px, want_logits=True, reuse=True)
    assert pre2["towers_skipped"]
    assert pre2["spliced_len"] == Tp1 + n1 + 7
    # the last generated token was never fed back, so n1-1 of the answer's KV entries exist
    assert pre2["reused_len"] == Tp1 + n1 - 1
    n2 = 4
    dec2, dec2_logits = eng.decode(n2 - 1, want_logits=True)
    torch.cuda.synchronize()
    got_ids = torch.cat([pre2["next_ids"][:, None], dec2], dim=1).cpu().long()
    got_logits = [pre2["logits"]] + [dec2_logits[s] for s in range(n2 - 1)]
    # vs the oracle on the full turn-2 prompt
    ref_ids, ref_logits = orc.greedy_generate(ids2, sd, g, *px, n2, return_logits=True)
    alive = torch.ones(B, dtype=torch.bool)
    for s in range(n2):
        err = (got_logits[s].float().cpu() - ref_logits[s]).abs().max(dim=-1).values
        assert bool((err[alive] <= 1e-2).all()), (s, err.tolist())
        top2 = ref_logits[s].topk(2, dim=-1).values
        same = got_ids[:, s] == ref_ids[:, s]
        assert not bool((alive & ~same & ((top2[:, 0] - top2[:, 1]) > 6e-2)).any()), (s, got_ids[:, s], ref_ids[:, s])
        alive &= same
    # vs an uncached prefill of the same prompt on a fresh handle
    ref2 = _fresh_logits(g, sd, ids2, px, B)
    assert max_abs(pre2["logits"], ref2["logits"]) <= 1e-2
    # ---- turn 3 = "regenerate": the same prompt again -> everything but the last position is reused
    pre3 = eng.prefill(ids2, *px, want_logits=True, reuse=True)
    assert pre3["towers_skipped"] and pre3["reused_len"] == pre2["spliced_len"] - 1
    assert max_abs(pre3["logits"], pre2["logits"]) <= 2e-3
    assert torch.equal(pre3["next_ids"], pre2["next_ids"])


def test_second_turn_against_the_reference_fixture():
    """The reference's own two-turn run (tests/golden/tiny_ds_turn2.npz, written by tools/make_golden.py --multi-turn: no
    cache, full recompute per turn) vs the product with the prefix cache on."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_ds_turn2.npz"))
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    assert synth.state_dict_digest(sd) == str(z["weights_digest"])
    px = [synth.make_pixels(g, 2, s) for s in (1, 2, 3)]
    eng = Engine(g, max_batch=2, max_seq_len=192)
    eng.bind(sd)
    ids1, ids2 = torch.from_numpy(z["ids1"]), torch.from_numpy(z["ids2"])
    n1, n2 = z["answer1"].shape[1], z["answer2"].shape[1]

    def run_turn(ids, n, ref_ids, ref_logits):
        pre = eng.prefill(ids, *px, want_logits=True, reuse=True)
        dec, dec_logits = eng.decode(n - 1, want_logits=True)
        torch.cuda.synchronize()
        got = torch.cat([pre["next_ids"][:, None], dec], dim=1).cpu().long()
        logits = [pre["logits"]] + [dec_logits[s] for s in range(n - 1)]
        alive = torch.ones(2, dtype=torch.bool)
        for s in range(n):
            rl = torch.from_numpy(ref_logits[s])
            err = (logits[s].float().cpu() - rl).abs().max(dim=-1).values
            assert bool((err[alive] <= 1e-2).all()), (s, err.tolist())
            top2 = rl.topk(2, dim=-1).values
            same = got[:, s] == torch.from_numpy(ref_ids[:, s])
            assert not bool((alive & ~same & ((top2[:, 0] - top2[:, 1]) > 6e-2)).any()), (s, got[:, s].tolist(), ref_ids[:, s].tolist())
            alive &= same
        return pre, got

    pre1, got1 = run_turn(ids1, n1, z["answer1"], z["logits1"])
    assert pre1["spliced_len"] == int(z["spliced_len1"]) and pre1["reused_len"] == 0 and not pre1["towers_skipped"]
    pre2, _ = run_turn(ids2, n2, z["answer2"], z["logits2"])
    assert pre2["spliced_len"] == int(z["spliced_len2"]) and pre2["towers_skipped"]
    # what is reusable: the turn-1 prompt plus as many of the n1-1 fed-back tokens as agree with the answer the fixture's
    # turn-2 prompt contains (all of them unless a near-tie flipped a token on the GPU)
    fed = got1[:, : n1 - 1]
    ref = torch.from_numpy(z["answer1"])[:, : n1 - 1]
    agree = min(int((fed[b] == ref[b]).long().cumprod(0).sum()) for b in range(2))
    assert pre2["reused_len"] == int(z["spliced_len1"]) + agree


def test_new_images_only_reuse_the_text_before_them():
    B = 2
    g, sd, ids, px, eng = _setup(B)
    eng.prefill(ids, *px, reuse=True)
    eng.decode(3)
    px2 = [synth.make_pixels(g, B, s) for s in (11, 12, 13)]
    pre = eng.prefill(ids, *px2, want_logits=True, reuse=True)
    assert not pre["towers_skipped"]
    assert pre["reused_len"] == 1 + 6  # bos + the 6 ids in front of <image>
    ref = _fresh_logits(g, sd, ids, px2, B)
    assert max_abs(pre["logits"], ref["logits"]) <= 2e-3
    assert torch.equal(pre["next_ids"].cpu(), ref["next_ids"].cpu())
    # only the depth image changes: with [IMAGE, DEPTH, SEG] it never reaches the sequence (F5), but the pixels differ,
    # so the towers run again and the visual positions are not trusted
    pre = eng.prefill(ids, px2[0], px2[1], px[2], want_logits=True, reuse=True)
    assert not pre["towers_skipped"] and pre["reused_len"] == 7
    assert max_abs(pre["logits"], ref["logits"]) <= 2e-3


def test_batch_rows_share_the_shortest_common_prefix():
    B = 2
    g, sd, ids, px, eng = _setup(B)
    pre1 = eng.prefill(ids, *px, reuse=True)
    Tp = pre1["spliced_len"]
    ids2 = ids.clone()
    ids2[1, -2] = (ids2[1, -2] + 1) % g.llm_vocab  # row 1 diverges two ids before the end, row 0 not at all
    pre2 = eng.prefill(ids2, *px, want_logits=True, reuse=True)
    assert pre2["towers_skipped"] and pre2["reused_len"] == Tp - 2
    ref = _fresh_logits(g, sd, ids2, px, B)
    assert max_abs(pre2["logits"], ref["logits"]) <= 2e-3
    # a different batch size never reuses the cache
    pre3 = eng.prefill(ids2[:1], *[p[:1] for p in px], want_logits=True, reuse=True)
    assert pre3["reused_len"] == 0 and not pre3["towers_skipped"]
    ref = _fresh_logits(g, sd, ids2[:1], [p[:1] for p in px], 1)
    assert max_abs(pre3["logits"], ref["logits"]) <= 2e-3


def test_cache_is_invalidated_by_other_entry_points():
    B = 2
    g, sd, ids, px, eng = _setup(B)
    eng.prefill(ids, *px, reuse=True)
    eng.encode(0, px[1])  # overwrites the handle's pixel / feature buffers
    pre = eng.prefill(ids, *px, want_logits=True, reuse=True)
    assert not pre["towers_skipped"] and pre["reused_len"] == 7
    ref = _fresh_logits(g, sd, ids, px, B)
    assert max_abs(pre["logits"], ref["logits"]) <= 2e-3
    eng.reset()
    pre = eng.prefill(ids, *px, want_logits=True, reuse=True)
    assert not pre["towers_skipped"] and pre["reused_len"] == 0
    # re-binding weights invalidates both the cached features and the cached KV entries
    eng.bind({k: v for k, v in synth.make_state_dict(g, seed=1).items() if k.startswith("model.layers.0.")})
    pre = eng.prefill(ids, *px, want_logits=True, reuse=True)
    assert not pre["towers_skipped"] and pre["reused_len"] == 0
    eng.bind(sd)
    pre = eng.prefill(ids, *px, want_logits=True, reuse=True)
    assert not pre["towers_skipped"] and pre["reused_len"] == 0
    assert max_abs(pre["logits"], ref["logits"]) <= 2e-3
    # plain vcb_prefill never reuses, but what it leaves behind can be reused by the next cached call (text only:
    # it does not fingerprint its pixels)
    eng.prefill(ids, *px)
    pre = eng.prefill(ids, *px, want_logits=True, reuse=True)
    assert not pre["towers_skipped"] and pre["reused_len"] == 7
    assert max_abs(pre["logits"], ref["logits"]) <= 2e-3


def test_streaming_rollback_then_cached_turn():
    """serve/cli.py flow: turn 1 streams and is stopped by the stopping criteria (queued lookahead steps are rolled back);
    turn 2's cached prefill must reuse exactly the accepted tokens."""
    B = 1
    g, sd, ids, px, eng = _setup(B)
    px = [p[:1] for p in px]
    ids = ids[:1]
    pre1 = eng.prefill(ids, *px, reuse=True)
    got = [int(pre1["next_ids"][0])]

    def on_token(step, toks):
        got.append(toks[0])
        return len(got) >= 4  # stop after 4 tokens in total

    n = eng.decode_stream(10, on_token, lookahead=4)
    assert n == 3 and len(got) == 4
    answer = torch.tensor([got])
    ids2 = torch.cat([ids, answer, torch.tensor([[5, 6, 7]])], dim=1)
    pre2 = eng.prefill(ids2, *px, want_logits=True, reuse=True)
    assert pre2["towers_skipped"] and pre2["reused_len"] == pre1["spliced_len"] + 3
    ref = _fresh_logits(g, sd, ids2, px, B)
    assert max_abs(pre2["logits"], ref["logits"]) <= 1e-2


def test_shim_generate_uses_the_cache_for_interactive_calls():
    from vcoder_b200.model import VCoderDSLlavaLlamaForCausalLM
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    m = VCoderDSLlavaLlamaForCausalLM(g, sd, max_batch=1, max_seq_len=192)
    ids = synth.make_prompt_ids(g, 1, seed=5, n_before=6, n_after=5)
    px = [synth.make_pixels(g, 1, s) for s in (1, 2, 3)]
    stop = [lambda seq, scores: seq.shape[1] >= ids.shape[1] + 4]
    o1 = m.generate(ids, images=px[0], segs=px[1], depths=px[2], max_new_tokens=8, stopping_criteria=stop)
    assert m.last_prefill["reused_len"] == 0 and not m.last_prefill["towers_skipped"]
    ids2 = torch.cat([o1.cpu(), torch.tensor([[9, 10]])], dim=1)
    stop2 = [lambda seq, scores: seq.shape[1] >= ids2.shape[1] + 3]
    o2 = m.generate(ids2, images=px[0], segs=px[1], depths=px[2], max_new_tokens=8, stopping_criteria=stop2)
    assert m.last_prefill["towers_skipped"] and m.last_prefill["reused_len"] == m.last_prefill["spliced_len"] - 2 - 1
    # same tokens as a model that never caches
    m.prefix_cache = False
    o3 = m.generate(ids2, images=px[0], segs=px[1], depths=px[2], max_new_tokens=8, stopping_criteria=stop2)
    assert m.last_prefill["reused_len"] == 0
    ref_ids = orc.greedy_generate(ids2, sd, g, *px, 3)
    for o in (o2, o3):
        assert o.shape == (1, ids2.shape[1] + 3)
    # greedy ids: cached == uncached == oracle unless a near-tie sits inside the fp16 noise floor
    _, ref_logits = orc.greedy_generate(ids2, sd, g, *px, 3, return_logits=True)
    top2 = ref_logits[0].topk(2, dim=-1).values
    if float(top2[0, 0] - top2[0, 1]) > 6e-2:
        assert int(o2[0, ids2.shape[1]]) == int(o3[0, ids2.shape[1]]) == int(ref_ids[0, 0])
