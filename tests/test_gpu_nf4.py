"""GPU tests of `load_4bit` (NF4 + double quantisation as a bind-time transform): the CUDA op against the numpy oracle, bit
for bit (integer level indices, fp16 values), and the engine end to end with NF4-rounded weights against the fp32 oracle
run on the SAME rounded weights.  Parity unpinned with respect to bitsandbytes itself (see oracle/bnb_nf4_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import bnb_nf4_oracle as nf4_orc
from oracle import vcoder_oracle as orc
from vcoder_b200 import nf4, synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from vcoder_b200.engine import Engine


@pytest.mark.parametrize("n,scale,dq", [(64, 0.02, 1), (64 * 255, 0.02, 1), (64 * 256 * 3 + 64, 1.0, 1), (4096 * 1024, 0.02, 1),
                                        (64 * 1000, 0.02, 0), (11008 * 512, 0.05, 1)])
def test_nf4_roundtrip_op_matches_the_oracle_bit_for_bit(n, scale, dq):
    g = torch.Generator().manual_seed(n)
    w = (torch.randn(n, generator=g) * scale).half()
    w[:64] = 0  # an all-zero block
    ref_q = nf4_orc.quantize_nf4(w.numpy(), double_quant=bool(dq))
    ref = nf4_orc.dequantize_nf4(ref_q)
    t = w.cuda().clone()
    idx = nf4.nf4_roundtrip_(t, double_quant=bool(dq), want_idx=True)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), ref_q["idx"])
    assert np.array_equal(t.cpu().view(torch.int16).numpy(), ref.view(np.int16))


def test_engine_with_load_4bit_weights():
    """Engine.bind(nf4=True) == binding weights that went through the oracle's NF4 round trip; logits then agree with the
    fp32 oracle evaluated on those rounded weights within the usual fp16 tolerance."""
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    sd = synth.make_state_dict(g, seed=0)
    sd_q = {k: (torch.from_numpy(nf4_orc.roundtrip_nf4(v.numpy())) if nf4.is_bnb_linear(k) else v) for k, v in sd.items()}
    ids = synth.make_prompt_ids(g, 2, seed=4, n_before=6, n_after=5)
    px = [synth.make_pixels(g, 2, s) for s in (1, 2, 3)]
    outs = []
    for weights, flag in ((sd, True), (sd_q, False)):
        eng = Engine(g, max_batch=2, max_seq_len=128)
        eng.bind(weights, nf4=flag)
        pre = eng.prefill(ids, *px, want_logits=True)
        dec = eng.decode(5)
        torch.cuda.synchronize()
        outs.append((pre["logits"].float().cpu(), dec.cpu()))
        eng.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref_ids, ref_logits = orc.greedy_generate(ids, {k: v.float() for k, v in sd_q.items()}, g, *px, 1, return_logits=True)
    assert float((outs[0][0] - ref_logits[0]).abs().max()) <= 1e-2
    # and the quantisation really changed the model
    ref_fp, _ = orc.greedy_generate(ids, {k: v.float() for k, v in sd.items()}, g, *px, 1, return_logits=True)
    _, lg_fp = orc.greedy_generate(ids, {k: v.float() for k, v in sd.items()}, g, *px, 1, return_logits=True)
    assert float((lg_fp[0] - ref_logits[0]).abs().max()) > 5e-2
