"""Launch the hot kernels at their production shapes a few times each (for `ncu -k regex:...`)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vcoder_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
st = C.c_void_p(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

if which in ("all", "skinny"):
    B = 8
    for (N, K, sw) in [(12288, 4096, 0), (4096, 4096, 0), (22016, 4096, 1), (4096, 11008, 0), (32000, 4096, 0)]:
        X = (torch.randn(B, K, device=dev) * 0.5).half()
        Ws = [(torch.randn(N, K, device=dev) * 0.03).half() for _ in range(3)]
        No = N // 2 if sw else N
        Y = torch.empty(B, No, device=dev, dtype=torch.float16)
        for i in range(3):
            lib.vcb_op_skinny_gemm_f16(_lib.ptr(X), K, _lib.ptr(Ws[i]), B, N, K, None, N, sw, _lib.ptr(Y), No, st)
        torch.cuda.synchronize()
if which in ("all", "attn"):
    Bi, S, heads = 24, 577, 16
    D = heads * 64
    qkv = torch.randn(Bi * S, 3 * D, device=dev).half()
    o = torch.empty(Bi * S, D, device=dev, dtype=torch.float16)
    for _ in range(3):
        flush.zero_()
        lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), Bi, S, heads, 0.125, _lib.ptr(o), st)
    torch.cuda.synchronize()
if which in ("all", "gemm"):
    for (M, N, K, act) in [(13848, 3072, 1024, 0), (13848, 4096, 1024, 1), (13848, 1024, 4096, 0), (9656, 12288, 4096, 0), (9656, 22016, 4096, 3), (9656, 4096, 11008, 0)]:
        A = (torch.randn(M, K, device=dev) * 0.5).half()
        W = (torch.randn(N, K, device=dev) * 0.03).half()
        No = N // 2 if act == 3 else N
        Cc = torch.empty(M, No, device=dev, dtype=torch.float16)
        for _ in range(2):
            lib.vcb_op_gemm_f16(_lib.ptr(A), K, _lib.ptr(W), K, M, N, K, None, None, N, act, _lib.ptr(Cc), No, 0, st)
        torch.cuda.synchronize()
if which in ("all", "attn_llama"):
    # LLaMA attention at the benched shape: prefill B = 8, T' = 1207 (tcgen05 kernel); decode B = 8, ctx = 1271 (streaming kernel)
    import math
    B, heads, T = 8, 32, 1207
    H = heads * 128
    max_pages = (T + 128 + 63) // 64
    n_pages = B * max_pages
    k_pool = torch.randn(n_pages, heads, 64, 128, device=dev).half()
    v_pool = torch.randn(n_pages, heads, 64, 128, device=dev).half()
    bt = torch.arange(n_pages, device=dev, dtype=torch.int32).view(B, max_pages).contiguous()
    q = torch.randn(B * T, 3 * H, device=dev).half()
    out = torch.empty(B * T, H, device=dev, dtype=torch.float16)
    for _ in range(2):
        _lib.check(lib.vcb_op_attention_prefill_f16(_lib.ptr(q), 3 * H, _lib.ptr(k_pool), _lib.ptr(v_pool), _lib.ptr(bt), max_pages, B, heads,
                                                    T, 0, 1.0 / math.sqrt(128), _lib.ptr(out), H, st))
    qkv = torch.randn(B, 3 * H, device=dev).half()
    od = torch.empty(B, H, device=dev, dtype=torch.float16)
    for _ in range(3):
        flush.zero_()
        _lib.check(lib.vcb_op_attention_decode_f16(_lib.ptr(qkv), _lib.ptr(k_pool), _lib.ptr(v_pool), _lib.ptr(bt), max_pages, B, heads, 1271, 1335,
                                                   1, 10000.0, 1.0 / math.sqrt(128), _lib.ptr(od), st))
    torch.cuda.synchronize()
print("done")
