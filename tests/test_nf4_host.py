"""CPU tests of the NF4 oracle (oracle/bnb_nf4_oracle.py) and of the host side of `load_4bit` (vcoder_b200/nf4.py).

PARITY UNPINNED: bitsandbytes is not available offline and the reference holds no vectors for it, so these are properties
of the published algorithm (QLoRA appendix E: the 16 levels are quantiles of N(0,1) normalised to [-1, 1] with an exact
zero; levels and decision thresholds as in bitsandbytes' dQuantizeNF4 / dDequantizeNF4), not known answers."""
import numpy as np

from oracle import bnb_nf4_oracle as orc
from vcoder_b200 import nf4


def test_levels_and_thresholds():
    lv = orc.NF4_LEVELS
    assert lv.shape == (16,) and lv[0] == -1.0 and lv[7] == 0.0 and lv[15] == 1.0 and np.all(np.diff(lv) > 0)
    thr = orc.nf4_thresholds()
    # the constants of bitsandbytes' decision tree (csrc/kernels.cu dQuantizeNF4), to float32 precision
    want = [-0.8480964004993439, -0.6106329262256622, -0.4599952697753906, -0.33967943489551544, -0.23460740596055984,
            -0.13791173323988914, -0.045525018125772476, 0.03979014977812767, 0.1202552504837513, 0.2035212516784668,
            0.2920137718319893, 0.3893125355243683, 0.5016634166240692, 0.6427869200706482, 0.8614784181118011]
    assert np.allclose(thr, np.array(want, dtype=np.float32), rtol=0, atol=1e-7)


def test_dynamic_map_matches_host_table_and_is_a_valid_code():
    code = orc.dynamic_map_8bit()
    assert np.array_equal(code, nf4.dynamic_map_8bit())  # product host logic == oracle
    assert code.shape == (256,) and len(np.unique(code)) == 256 and code[-1] == 1.0 and 0.0 in code
    assert np.allclose(np.sort(-code[code < 0]), np.sort(code[(code > 0) & (code < 1)]))  # signed, symmetric but for the 1.0


def test_roundtrip_properties():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(64 * 300) * 0.02).astype(np.float16)
    w[64:128] = 0  # an all-zero block stays zero
    q = orc.quantize_nf4(w, double_quant=False)
    r = orc.dequantize_nf4(q)
    assert np.all(r[64:128] == 0)
    blocks = w.astype(np.float32).reshape(-1, 64)
    am = np.abs(blocks).max(1)
    # the largest-magnitude element of every block maps to level +-1 and is reproduced exactly
    j = np.abs(blocks).argmax(1)
    assert np.allclose(np.abs(r.astype(np.float32).reshape(-1, 64)[np.arange(len(j)), j]), am, rtol=1e-3)
    # nearest-level assignment: no other level is closer
    x = blocks / np.where(am > 0, am, 1)[:, None]
    d = np.abs(x[..., None] - orc.NF4_LEVELS[None, None, :])
    assert np.array_equal(d.argmin(-1).astype(np.uint8).reshape(-1)[w != 0], q["idx"][w != 0]) or \
        np.mean(d.argmin(-1).reshape(-1) != q["idx"]) < 1e-3  # exact ties aside
    # idempotent without double quantisation; with it the error of the scales stays below the 8-bit code's resolution
    assert np.array_equal(orc.roundtrip_nf4(r, double_quant=False), r)
    rel = np.abs(orc.quantize_nf4(w)["absmax"] - am) / np.maximum(am, 1e-12)
    assert rel[am > 0].max() < 0.08
    err = np.sqrt(np.mean((orc.roundtrip_nf4(w).astype(np.float32) - w.astype(np.float32)) ** 2)) / np.sqrt(np.mean(w.astype(np.float32) ** 2))
    assert 0.05 < err < 0.12  # NF4 on Gaussian weights: ~9 % relative RMS


def test_which_weights_are_quantised():
    yes = ["model.layers.0.self_attn.q_proj.weight", "model.layers.39.mlp.down_proj.weight", "model.mm_projector.2.weight",
           "model.seg_mm_projector.0.weight", "model.depth_mm_projector.0.weight", "model.mm2_projector.2.weight"]
    no = ["lm_head.weight", "model.embed_tokens.weight", "model.norm.weight", "model.layers.0.input_layernorm.weight",
          "model.mm_projector.0.bias", "model.vision_tower.vision_tower.vision_model.encoder.layers.0.mlp.fc1.weight"]
    assert all(nf4.is_bnb_linear(k) for k in yes) and not any(nf4.is_bnb_linear(k) for k in no)
