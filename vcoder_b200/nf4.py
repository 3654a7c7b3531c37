"""Host side of `load_4bit=True` (reference: vcoder_llava/model/builder.py:28-37): which weights bitsandbytes would turn
into NF4 `Linear4bit` layers, the 8-bit "dynamic" code table of its double quantisation, and the call into
`vcb_op_nf4_roundtrip_f16` that replaces such a weight by what the 4-bit layer multiplies with.

HF's quantizer replaces every `nn.Linear` of the loaded model except `lm_head` (`llm_int8_skip_modules` default); the CLIP
tower is loaded afterwards in fp16 (`builder.py:138-143`) and the embeddings / norms are not Linear layers.  In state-dict
terms: `model.layers.*.{self_attn.{q,k,v,o}_proj,mlp.{gate,up,down}_proj}.weight` and `model.*mm_projector.{0,2}.weight`.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np
import torch

from . import _lib

_QUANTISED = re.compile(r"^model\.(layers\.\d+\.(self_attn\.[qkvo]_proj|mlp\.(gate|up|down)_proj)|\w*mm\d?_projector\.[02])\.weight$")


def is_bnb_linear(name: str) -> bool:
    return _QUANTISED.match(name) is not None


def _linspace_f32(start, end, n):
    # torch.linspace in float32: start + i*step for the first half, end - (n-1-i)*step for the second
    if n == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32((np.float32(end) - np.float32(start)) / np.float32(n - 1))
    return np.array([np.float32(start) + np.float32(i) * step if i < n // 2 else np.float32(end) - np.float32(n - 1 - i) * step
                     for i in range(n)], dtype=np.float32)


def dynamic_map_8bit() -> np.ndarray:
    """bitsandbytes.functional.create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8): 256 sorted values."""
    data = []
    for i in range(7):
        b = _linspace_f32(0.1, 1.0, 2 ** i + 1)
        means = ((b[:-1] + b[1:]) / np.float32(2.0)).astype(np.float32)
        scale = np.float32(10.0 ** (-6 + i))
        data += (scale * means).astype(np.float32).tolist() + (-(scale * means)).astype(np.float32).tolist()
    data += [0.0, 1.0]
    return np.sort(np.array(data, dtype=np.float32))


_CODE8 = None


def nf4_roundtrip_(w: torch.Tensor, double_quant: bool = True, want_idx: bool = False):
    """In place on a contiguous fp16 CUDA tensor with numel % 64 == 0; returns the level indices when asked."""
    global _CODE8
    if not (w.is_cuda and w.dtype == torch.float16 and w.is_contiguous() and w.numel() % 64 == 0):
        raise ValueError("nf4_roundtrip_: contiguous fp16 CUDA tensor with numel % 64 == 0 expected")
    if _CODE8 is None:
        _CODE8 = np.ascontiguousarray(dynamic_map_8bit())
    idx = torch.empty(w.numel(), dtype=torch.uint8, device=w.device) if want_idx else None
    lib = _lib.load()
    _lib.check(lib.vcb_op_nf4_roundtrip_f16(_lib.ptr(w), w.numel(), int(double_quant), _lib.ptr(_CODE8), _lib.ptr(w), _lib.ptr(idx),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return idx
