"""TMEM read and MUFU ex2 rates on this GPU (vcb_microbench): the two per-SM limits the ViT attention kernel is built
around.  Prints JSON lines."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vcoder_b200 import _lib  # noqa: E402

lib = _lib.load()
r = C.c_float(0)
for warps in (4, 8, 16):
    for ilp in (1, 2, 4):
        _lib.check(lib.vcb_microbench(0, warps, ilp, 2000, C.byref(r)))
        print(json.dumps({"bench": "tmem_read", "warps": warps, "loads_in_flight": ilp, "bytes_per_clk_per_sm": round(r.value, 1)}))
for warps in (4, 8, 16, 32):
    for fma in (0, 1):
        _lib.check(lib.vcb_microbench(1, warps, fma, 4000, C.byref(r)))
        print(json.dumps({"bench": "ex2", "warps": warps, "with_fma": fma, "ex2_per_clk_per_sm": round(r.value, 2)}))
for warps in (4, 8, 16):
    _lib.check(lib.vcb_microbench(2, warps, 0, 4000, C.byref(r)))
    print(json.dumps({"bench": "ex2_f16x2", "warps": warps, "ex2_per_clk_per_sm": round(r.value, 2)}))
VARIANTS = {0: "as in the kernel", 1: "no row sums", 2: "no P stores", 3: "no sums, no stores", 4: "skewed by 16", 5: "skewed, no sums",
            6: "skewed, no stores", 7: "skewed, no sums, no stores"}
for warps in (4, 8):
    for var in range(8):
        _lib.check(lib.vcb_microbench(3, warps, var, 2000, C.byref(r)))
        print(json.dumps({"bench": "softmax_exp_loop", "warps": warps, "variant": VARIANTS[var], "ex2_per_clk_per_sm": round(r.value, 2)}))
