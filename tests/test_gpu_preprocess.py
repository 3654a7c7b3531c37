"""GPU tests for SURVEY §8 row f2 (image preprocessing on the device, csrc/preprocess.cu) through the C ABI
(vcb_op_preprocess_u8 via vcoder_b200.preprocess.process_images).  Byte/integer work: the bar is BIT-EXACT — against
the committed outputs of the reference's own process_images, against the numpy oracle (pinned to Pillow on the CPU), and
against Pillow + numpy run live on the box."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ip(size):
    return types.SimpleNamespace(size={"shortest_edge": size}, crop_size={"height": size, "width": size},
                                 image_mean=list(po.CLIP_MEAN), image_std=list(po.CLIP_STD))


def _cfg(pad):
    return types.SimpleNamespace(image_aspect_ratio="pad" if pad else None)


def _bits(t):
    return t.detach().cpu().contiguous().view(torch.int16).numpy()


def test_matches_reference_fixture():
    from vcoder_b200.preprocess import process_images
    z = np.load(os.path.join(GOLD, "preprocess.npz"))
    n = 0
    for key in z.files:
        if not key.startswith("out_"):
            continue
        _, size, k, mode = key.split("_")
        got = process_images([z[f"img_{size}_{k}"]], _ip(int(size)), _cfg(mode == "pad"))
        assert got.shape == (1, 3, int(size), int(size)) and got.dtype == torch.float16 and got.is_cuda
        want = torch.from_numpy(z[key]).half()  # the callers' .to(dtype=torch.float16)
        assert np.array_equal(_bits(got[0]), _bits(want)), key
        n += 1
    assert n == 8


@pytest.mark.parametrize("h,w", [(480, 640), (640, 480), (336, 336), (700, 300), (123, 457), (1000, 1000), (1365, 2048),
                                 (20, 31)])
@pytest.mark.parametrize("pad", [False, True])
def test_matches_oracle_and_pillow_at_336(h, w, pad):
    from vcoder_b200.preprocess import process_images
    rng = np.random.default_rng(h + 7 * w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = process_images([img], _ip(336), _cfg(pad))[0]
    want = torch.from_numpy(po.clip_preprocess(img, 336, pad)).half()
    assert np.array_equal(_bits(got), _bits(want))
    # Pillow live: expand2square + resize + centre crop reproduce the uint8 stage; map back through the same LUT
    from PIL import Image
    pil = Image.fromarray(img)
    if pad and h != w:
        side = max(h, w)
        canvas = Image.new("RGB", (side, side), tuple(int(x * 255) for x in po.CLIP_MEAN))
        canvas.paste(pil, (0, (w - h) // 2) if w > h else ((h - w) // 2, 0))
        pil = canvas
    nh, nw = po.resize_target(pil.size[1], pil.size[0], 336)
    r = np.asarray(pil.resize((nw, nh), Image.BICUBIC))
    top, left = (nh - 336) // 2, (nw - 336) // 2
    c = r[top:top + 336, left:left + 336]
    lut = torch.from_numpy(po.normalize_lut()).half()
    via_pil = torch.stack([lut[ch][torch.from_numpy(c[..., ch].astype(np.int64))] for ch in range(3)])
    assert np.array_equal(_bits(got), _bits(via_pil))


def test_batch_pil_input_and_tower_feed():
    """PIL images in (as serve/cli.py passes them), fp16 device tensor out, consumable by the tower without a copy."""
    from PIL import Image
    from vcoder_b200 import synth
    from vcoder_b200.engine import Engine
    from vcoder_b200.preprocess import process_images
    g = synth.PRESETS["tiny"]
    rng = np.random.default_rng(3)
    pils = [Image.fromarray(rng.integers(0, 256, (70 + 9 * i, 50 + 13 * i, 3), dtype=np.uint8)) for i in range(3)]
    px = process_images(pils, _ip(g.vit_image_size), _cfg(True))
    assert px.shape == (3, 3, g.vit_image_size, g.vit_image_size)
    for i, p in enumerate(pils):
        want = torch.from_numpy(po.clip_preprocess(np.asarray(p), g.vit_image_size, True)).half()
        assert np.array_equal(_bits(px[i]), _bits(want))
    eng = Engine(g, max_batch=1, max_seq_len=64)
    eng.bind(synth.make_state_dict(g, seed=0))
    feats = eng.tower_forward(px)
    assert feats.shape == (3, g.n_patches, g.vit_hidden) and bool(torch.isfinite(feats.float()).all())
    with pytest.raises(ValueError):
        process_images([np.zeros((4, 4), np.uint8)], _ip(56), _cfg(False))
