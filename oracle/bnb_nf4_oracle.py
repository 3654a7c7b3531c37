"""CPU restatement (numpy) of bitsandbytes' NF4 weight format as the reference requests it — TEST INFRASTRUCTURE ONLY
(imported by tests/ and never by the product path).

Reference call site: vcoder_llava/model/builder.py:28-37 — `load_4bit=True` -> `BitsAndBytesConfig(load_in_4bit=True,
bnb_4bit_compute_dtype=torch.float16, bnb_4bit_use_double_quant=True, bnb_4bit_quant_type='nf4')`, i.e. every nn.Linear of
the language model and of the projectors except `lm_head` becomes `bnb.nn.Linear4bit`; its forward de-quantises the weight
to fp16 and runs an fp16 matmul.

PARITY UNPINNED: `bitsandbytes` (pinned at 0.41.0 in the reference's pyproject.toml:30) is neither under /root/reference nor
installable here, and the reference holds no golden vectors for it.  What is restated is the published algorithm
(Dettmers et al., "QLoRA", 2023, section 3 + appendix E; bitsandbytes/functional.py `quantize_4bit`, `create_dynamic_map`,
`quantize_blockwise`; csrc/kernels.cu `dQuantizeNF4` / `dDequantizeNF4`):
  * blocks of 64 weights; absmax per block (fp32); x / absmax is mapped to the NEAREST of the 16 NF4 levels (the CUDA kernel
    walks a decision tree whose thresholds are the midpoints between adjacent levels, strict `>` on the way up);
  * double quantisation: the fp32 absmax vector minus its mean is quantised in blocks of 256 to the 8-bit "dynamic" code
    (create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8)), nearest code value;
  * de-quantisation: level[idx] * absmax' in fp32, rounded to fp16; absmax' = code8[q] * absmax2 + mean.
"""
import numpy as np

NF4_LEVELS = np.array([
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635, -0.18477343022823334,
    -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
    0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0], dtype=np.float32)


def nf4_thresholds():
    """The 15 decision thresholds of dQuantizeNF4 = midpoints of adjacent levels (as float32)."""
    lv = NF4_LEVELS.astype(np.float64)
    return ((lv[:-1] + lv[1:]) / 2.0).astype(np.float32)


def _linspace_f32(start, end, n):
    """torch.linspace in float32: start + i*step for the first half, end - (n-1-i)*step for the second."""
    if n == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32((np.float32(end) - np.float32(start)) / np.float32(n - 1))
    out = np.empty(n, dtype=np.float32)
    half = n // 2
    for i in range(n):
        out[i] = np.float32(start) + np.float32(i) * step if i < half else np.float32(end) - np.float32(n - 1 - i) * step
    return out


def dynamic_map_8bit():
    """create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8): 256 sorted float32 code values in [-1, 1]."""
    data = []
    max_exponent_bits, non_sign_bits = 7, 7
    for i in range(max_exponent_bits):
        fraction_items = 2 ** (i + non_sign_bits - max_exponent_bits) + 1
        b = _linspace_f32(0.1, 1.0, fraction_items)
        means = ((b[:-1] + b[1:]) / np.float32(2.0)).astype(np.float32)
        scale = np.float32(10.0 ** (-(max_exponent_bits - 1) + i))
        data += (scale * means).astype(np.float32).tolist()
        data += (-(scale * means)).astype(np.float32).tolist()
    data.append(0.0)
    data.append(1.0)
    assert len(data) == 256
    return np.sort(np.array(data, dtype=np.float32))


def _nearest_code(x, code):
    """index of the nearest code value (ties: the lower index), x float32 array, code sorted float32."""
    pos = np.searchsorted(code, x)  # code[pos-1] < x <= code[pos]
    pos = np.clip(pos, 1, len(code) - 1)
    lo, hi = code[pos - 1], code[pos]
    take_hi = (hi.astype(np.float64) - x.astype(np.float64)) < (x.astype(np.float64) - lo.astype(np.float64))
    return np.where(take_hi, pos, pos - 1).astype(np.int32)


def quantize_nf4(w_half, double_quant=True):
    """w_half: fp16 array, size % 64 == 0 -> dict(idx uint8 [n], absmax fp32 [n/64] (what de-quantisation will use),
    and with double_quant the 8-bit codes / second-level scales / mean)."""
    w = np.asarray(w_half, dtype=np.float16).astype(np.float32).reshape(-1, 64)
    absmax = np.abs(w).max(axis=1).astype(np.float32)
    inv = (np.float32(1.0) / np.where(absmax > 0, absmax, np.float32(1.0))).astype(np.float32)  # the kernel multiplies by 1/absmax
    x = (w * inv[:, None]).astype(np.float32)
    thr = nf4_thresholds()
    idx = (x[..., None] > thr[None, None, :]).sum(axis=-1).astype(np.uint8)  # number of thresholds strictly below x
    out = {"idx": idx.reshape(-1), "absmax_raw": absmax}
    if not double_quant:
        out["absmax"] = absmax
        return out
    offset = np.float32(absmax.astype(np.float64).mean())
    a = (absmax - offset).astype(np.float32)
    nb = a.size
    pad = (-nb) % 256
    a2 = np.concatenate([a, np.zeros(pad, np.float32)]).reshape(-1, 256)
    absmax2 = np.abs(a2).max(axis=1).astype(np.float32)
    inv2 = (np.float32(1.0) / np.where(absmax2 > 0, absmax2, np.float32(1.0))).astype(np.float32)
    code = dynamic_map_8bit()
    q = _nearest_code((a2 * inv2[:, None]).astype(np.float32), code)
    deq = (code[q] * absmax2[:, None]).astype(np.float32).reshape(-1)[:nb]
    out.update({"q8": q.reshape(-1)[:nb].astype(np.uint8), "absmax2": absmax2, "offset": offset,
                "absmax": (deq + offset).astype(np.float32)})
    return out


def dequantize_nf4(q):
    """-> fp16 array: level[idx] * absmax in fp32, one rounding to fp16 (dDequantizeNF4 + the cast of the output type)."""
    lv = NF4_LEVELS[q["idx"].reshape(-1, 64)]
    return (lv * q["absmax"][:, None]).astype(np.float32).astype(np.float16).reshape(-1)


def roundtrip_nf4(w_half, double_quant=True):
    """What a Linear4bit layer multiplies by: dequantize(quantize(W)) as fp16, same shape."""
    w = np.asarray(w_half, dtype=np.float16)
    return dequantize_nf4(quantize_nf4(w.reshape(-1), double_quant)).reshape(w.shape)
