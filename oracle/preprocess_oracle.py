"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/vcoder_oracle.py for the rules).

numpy restatement of the image preprocessing the reference runs on the CPU before every prefill:

    process_images / expand2square                vcoder_llava/mm_utils.py:14-40
      -> CLIPImageProcessor.preprocess            third party: transformers (pinned 4.31.0) image_processing_clip.py:
         convert RGB -> resize(shortest_edge=336, PIL BICUBIC) -> center_crop(336) -> rescale(1/255) -> normalize(mean, std)
      -> `.to(device, dtype=torch.float16)`       serve/cli.py:74-88, eval/model_depth_loader.py:120-140

The resize is Pillow's ImagingResample (third party: Pillow, src/libImaging/Resample.c): a separable two-pass
convolution — horizontal first, uint8 intermediate — with the bicubic kernel (a = -0.5) widened by the down-scale factor
(antialiasing), coefficients normalised in double and rounded to 22-bit fixed point, accumulators seeded with 1 << 21 and
clipped to [0, 255] after `>> 22`.  It is pure integer arithmetic on bytes, so the bar is BIT-EXACT.

Pinning: tests/test_preprocess_host.py checks every function here against the installed Pillow (`Image.resize(...,
Image.BICUBIC)`) and the installed transformers `CLIPImageProcessor` on random images of many sizes.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # OPENAI_CLIP_MEAN (transformers/utils/constants.py)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)  # OPENAI_CLIP_STD


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for the full-image box [0, in_size).
    -> bounds [out,2] int32 (xmin, count), coeffs [out, ksize] int32, ksize."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        xmax -= xmin
        w = np.array([_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)], dtype=np.float64)
        ww = 0.0
        for v in w:  # same left-to-right double accumulation as the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            v = float(w[x]) * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit convolution pass of ImagingResampleHorizontal/Vertical_8bpc along `axis` of an HWC uint8 image."""
    in_size = img.shape[axis]
    bounds, kk, _ = bicubic_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)  # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h), Image.BICUBIC)` on an HWC uint8 RGB array: horizontal pass (skipped when the
    width is unchanged), then vertical pass (skipped when the height is unchanged) — ImagingResample."""
    h, w = img.shape[:2]
    cur = img
    if out_w != w:
        cur = _pass(cur, out_w, axis=1)
    if out_h != h:
        cur = _pass(cur, out_h, axis=0)
    return cur


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """mm_utils.py:14-26: paste centred on a square canvas of the background colour."""
    h, w = img.shape[:2]
    if w == h:
        return img
    side = max(w, h)
    out = np.empty((side, side, 3), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    if w > h:
        top = (w - h) // 2
        out[top:top + h, :] = img
    else:
        left = (h - w) // 2
        out[:, left:left + w] = img
    return out


def resize_target(h: int, w: int, size: int) -> Tuple[int, int]:
    """get_resize_output_image_size(shortest_edge=size, default_to_square=False) -> (new_h, new_w): the short side becomes
    `size`, the long side int(size * long / short) (transformers/image_transforms.py)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def normalize_lut(mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD) -> np.ndarray:
    """[3,256] float32: rescale (uint8 * (1/255) in float64 -> float32) then normalize ((x - mean) / std in float32), as
    image_transforms.rescale / normalize do it; the fp16 cast is the caller's `.to(dtype=torch.float16)`."""
    x = (np.arange(256, dtype=np.uint8).astype(np.float64) * (1 / 255)).astype(np.float32)
    m = np.asarray(mean, dtype=np.float32)
    s = np.asarray(std, dtype=np.float32)
    return ((x[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def clip_preprocess(img: np.ndarray, size: int = 336, pad: bool = False, mean: Sequence[float] = CLIP_MEAN,
                    std: Sequence[float] = CLIP_STD) -> np.ndarray:
    """process_images for ONE HWC uint8 RGB image -> [3, size, size] float32 (cast to fp16 by the caller).
    pad=True is `image_aspect_ratio == 'pad'` (mm_utils.py:31-35): background = int(mean * 255)."""
    if pad:
        img = expand2square(img, tuple(int(x * 255) for x in mean))
    h, w = img.shape[:2]
    nh, nw = resize_target(h, w, size)
    r = resize_bicubic_u8(img, nw, nh)
    top, left = (nh - size) // 2, (nw - size) // 2  # center_crop (image_transforms.center_crop)
    c = r[top:top + size, left:left + size]
    lut = normalize_lut(mean, std)
    out = np.empty((3, size, size), dtype=np.float32)
    for ch in range(3):
        out[ch] = lut[ch][c[..., ch]]
    return out
