"""GPU image preprocessing — drop-in for the reference's `process_images` (vcoder_llava/mm_utils.py:28-40).

Same signature and result as the reference (`process_images(images, image_processor, model_cfg)` -> pixel_values
[n, 3, S, S]), except that the tensor already lives on the GPU in fp16 — exactly what every caller does next
(`.to(model.device, dtype=torch.float16)`, serve/cli.py:74-88, eval/model_depth_loader.py:120-140).  The resize is
bit-exact with Pillow's bicubic (libvcoder_b200: csrc/preprocess.cu); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib


def _as_rgb_u8(image) -> np.ndarray:
    if hasattr(image, "convert"):  # PIL.Image: do_convert_rgb (image_processing_clip.py)
        image = image.convert("RGB")
    a = np.asarray(image)
    if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
        raise ValueError(f"expected an RGB uint8 HWC image, got {a.dtype} {a.shape}")
    return np.ascontiguousarray(a)


def _edge(v, *keys) -> int:
    if isinstance(v, dict):
        for k in keys:
            if k in v:
                return int(v[k])
        raise ValueError(f"cannot read a size from {v}")
    if isinstance(v, (tuple, list)):
        return int(v[0])
    return int(v)


def process_images(images: Sequence, image_processor, model_cfg, device: int = 0) -> torch.Tensor:
    """mm_utils.py:28-40.  `image_processor` needs image_mean / image_std / size / crop_size (a CLIPImageProcessor or any
    namespace carrying them); `model_cfg.image_aspect_ratio == 'pad'` selects expand2square with the mean colour."""
    if not torch.cuda.is_available():
        raise RuntimeError("vcoder_b200.preprocess needs a CUDA device (no CPU fallback exists)")
    lib = _lib.load()
    size = _edge(image_processor.size, "shortest_edge", "height")
    crop = _edge(getattr(image_processor, "crop_size", size), "height", "shortest_edge")
    if crop != size:
        raise NotImplementedError("crop_size != size['shortest_edge'] does not occur for the CLIP towers the reference uses")
    mean = np.asarray(image_processor.image_mean, dtype=np.float32)
    std = np.asarray(image_processor.image_std, dtype=np.float32)
    pad = getattr(model_cfg, "image_aspect_ratio", None) == "pad"
    bg = np.asarray([int(x * 255) for x in image_processor.image_mean], dtype=np.uint8) if pad else None
    dev = torch.device("cuda", device)
    out = torch.empty(len(images), 3, size, size, dtype=torch.float16, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        for i, im in enumerate(images):
            a = _as_rgb_u8(im)
            _lib.check(lib.vcb_op_preprocess_u8(_lib.ptr(a), a.shape[0], a.shape[1], _lib.ptr(bg), size, _lib.ptr(mean),
                                                _lib.ptr(std), C.c_void_p(out[i].data_ptr()), st))
    return out
