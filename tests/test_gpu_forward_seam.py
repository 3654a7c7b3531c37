"""GPU tests of the two seams SURVEY.md section 8(b) lists besides generate(): `forward()` (all-position logits,
language_model/vcoder_ds_llava_llama.py:57-118) and `prepare_inputs_labels_for_multimodal` (vcoder_ds_llava_arch.py:126-314),
through vcb_forward_logits / vcb_splice_embeds, against the tensors the reference itself wrote into the golden fixtures
(`prefill_logits`, `inputs_embeds`, `spliced_mask`; tools/make_golden.py) and against the oracle."""
import dataclasses
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import vcoder_oracle as orc  # noqa: E402

from vcoder_b200 import _lib, synth  # noqa: E402
from vcoder_b200.engine import Engine  # noqa: E402
from vcoder_b200.model import build_model  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGIT_TOL, RELRMS_TOL = 1e-2, 2e-3  # the suite's bars


def rel_rms(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-12))


def max_abs(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max())


CASES = [(synth.VARIANT_VCODER_DS, "vcoder_ds_llava-tiny", "tiny_ds.npz"), (synth.VARIANT_VCODER, "vcoder_llava-tiny", "tiny_vcoder.npz"),
         (synth.VARIANT_LLAVA, "llava-tiny", "tiny_llava.npz")]


@pytest.mark.parametrize("variant,name,fname", CASES)
def test_forward_and_prepare_inputs_against_reference_golden(variant, name, fname):
    g = synth.PRESETS["tiny"].with_variant(variant)
    sd = synth.make_state_dict(g, seed=0)
    z = np.load(os.path.join(GOLD, fname))
    assert synth.state_dict_digest(sd) == str(z["weights_digest"])
    ids = torch.from_numpy(z["ids"]).cuda()
    B, T = ids.shape
    images = synth.make_pixels(g, B, 1).cuda()
    segs = synth.make_pixels(g, B, 2).cuda() if variant != synth.VARIANT_LLAVA else None
    depths = synth.make_pixels(g, B, 3).cuda() if variant == synth.VARIANT_VCODER_DS else None
    model = build_model(name, g, sd, max_batch=B, max_seq_len=128)
    kw = {"images": images}
    if segs is not None:
        kw["segs"] = segs
    if depths is not None:
        kw["depths"] = depths

    # ---- prepare_inputs_labels_for_multimodal: the reference's own 5-tuple
    mask = torch.ones(B, T, dtype=torch.bool, device=ids.device)
    r_ids, r_mask, r_pkv, r_embeds, r_labels = model.prepare_inputs_labels_for_multimodal(ids, mask, None, None, images, segs, depths)
    assert r_ids is None and r_pkv is None and r_labels is None
    assert r_embeds.shape == z["inputs_embeds"].shape and r_embeds.device == ids.device
    assert rel_rms(r_embeds, torch.from_numpy(z["inputs_embeds"])) <= RELRMS_TOL
    assert r_mask.shape == z["spliced_mask"].shape and bool(r_mask.all()) and r_mask.dtype == mask.dtype
    # text-only and single-token early return (:130-133): arguments come back unchanged
    out = model.prepare_inputs_labels_for_multimodal(ids, mask, None, None, None, None, None)
    assert out[0] is ids and out[1] is mask and out[3] is None
    with pytest.raises(NotImplementedError):
        model.prepare_inputs_labels_for_multimodal(ids, mask, None, ids.clone(), images, segs, depths)

    # ---- forward(): CausalLMOutputWithPast.logits of every spliced position == what the reference's forward() wrote
    res = model(input_ids=ids, **kw)
    ref = torch.from_numpy(z["prefill_logits"])
    assert type(res).__name__ == "CausalLMOutputWithPast" and res.loss is None
    assert res.logits.shape == ref.shape and res.logits.dtype == torch.float16
    assert rel_rms(res.logits, ref) <= RELRMS_TOL
    assert max_abs(res.logits, ref) <= LOGIT_TOL
    assert torch.equal(res.logits[:, -1].float().argmax(-1).cpu(), torch.from_numpy(z["greedy_ids"][:, 0]))
    tup = model.forward(input_ids=ids, return_dict=False, **kw)
    assert isinstance(tup, tuple) and torch.equal(tup[0], res.logits)  # deterministic
    # forward() fills the KV cache and selects the first token like a prefill: decoding continues from it
    n_new = z["greedy_ids"].shape[1]
    dec_ids = model._eng.decode(n_new - 1)
    assert np.array_equal(dec_ids.cpu().numpy(), z["greedy_ids"][:, 1:])
    # training arguments are refused, not ignored
    for bad in ({"labels": ids}, {"inputs_embeds": r_embeds}, {"past_key_values": ((None, None),)}, {"output_attentions": True}):
        with pytest.raises(NotImplementedError):
            model(input_ids=ids, **kw, **bad)


def test_forward_logits_text_only_and_last_row_consistency():
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    eng = Engine(g, max_batch=3, max_seq_len=128)
    eng.bind(sd)
    t_ids = torch.randint(3, g.llm_vocab, (3, 37), generator=torch.Generator().manual_seed(5))
    got = eng.forward_logits(t_ids)
    emb = sd["model.embed_tokens.weight"].float()[t_ids]
    ref = orc.llama_forward(emb, sd, g, orc.KVCache(g.llm_layers), last_only=False)
    assert got.shape == ref.shape == (3, 37, g.llm_vocab)
    assert rel_rms(got, ref) <= RELRMS_TOL and max_abs(got, ref) <= LOGIT_TOL
    # the last row agrees with the last-position logits of vcb_prefill (another GEMM kernel: same values to fp16 round-off)
    pre = eng.prefill(t_ids, want_logits=True)
    assert max_abs(got[:, -1], pre["logits"]) <= 4e-3
    # a buffer that cannot hold B * T' * V is refused
    small = torch.empty(16, dtype=torch.float16, device="cuda")
    import ctypes as C
    slen = C.c_int32(0)
    ids_h = t_ids.contiguous()
    rc = eng.lib.vcb_forward_logits(eng.h, _lib.ptr(ids_h), 3, 37, None, None, None, _lib.ptr(small), small.numel(), C.byref(slen), None)
    assert rc == -4  # VCB_ERR_NOMEM
    host = torch.empty(3 * 37 * g.llm_vocab, dtype=torch.float16)
    rc = eng.lib.vcb_forward_logits(eng.h, _lib.ptr(ids_h), 3, 37, None, None, None, _lib.ptr(host), host.numel(), C.byref(slen), None)
    assert rc == -1  # VCB_ERR_INVALID: the logits buffer must be device memory


def test_forward_logits_vocabulary_not_a_tile_multiple():
    """V % 32 != 0 takes the weight-streaming GEMM 32 rows at a time instead of the tcgen05 tiles."""
    g = dataclasses.replace(synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS), llm_vocab=528)
    sd = synth.make_state_dict(g, seed=3)
    eng = Engine(g, max_batch=2, max_seq_len=128)
    eng.bind(sd)
    ids = torch.tensor([[1, 7, synth.IMAGE_TOKEN_INDEX, 9, synth.SEG_TOKEN_INDEX, 11], [1, 8, synth.IMAGE_TOKEN_INDEX, 10, synth.SEG_TOKEN_INDEX, 12]])
    images, segs, depths = (synth.make_pixels(g, 2, s) for s in (1, 2, 3))
    got = eng.forward_logits(ids, images, segs, depths)
    x = orc.prefill_embeds(ids, sd, g, images, segs, depths)
    ref = orc.llama_forward(x, sd, g, orc.KVCache(g.llm_layers), last_only=False)
    assert got.shape == ref.shape
    assert rel_rms(got, ref) <= RELRMS_TOL and max_abs(got, ref) <= LOGIT_TOL
    emb = eng.splice_embeds(ids, images, segs, depths)
    assert rel_rms(emb, x) <= RELRMS_TOL
    with pytest.raises(_lib.VcbError):  # the handle is un-prefilled after vcb_splice_embeds
        eng.decode(1)
