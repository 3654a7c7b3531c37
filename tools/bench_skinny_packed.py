"""A/B of the decode weight stream: row-major W through the 3-D tensor map vs the tile-major packed copy through linear
bulk copies (vcb_op_skinny_gemm_f16 vs vcb_op_skinny_gemm_packed_f16), B = 8, the five decode shapes of LLaMA-7B.
Every launch streams a DIFFERENT weight buffer (>= 600 MB rotating set >> 126 MB L2), launches are back to back (PDL)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(12288, 4096, 0), (4096, 4096, 0), (22016, 4096, 1), (4096, 11008, 0), (32000, 4096, 0)]


def main():
    import torch
    from vcoder_b200 import _lib
    lib = _lib.load()
    dev, st, B = "cuda", C.c_void_p(0), 8
    out = {}
    tot = {"row_major": [0.0, 0.0], "packed": [0.0, 0.0]}
    for (N, K, sw) in SHAPES:
        X = (torch.randn(B, K, device=dev) * 0.5).half()
        n_copies = max(3, int(600e6 // (N * K * 2)) + 1)
        Ws = [(torch.randn(N, K, device=dev) * 0.03).half() for _ in range(n_copies)]
        Wps = [torch.empty(N * K, device=dev, dtype=torch.float16) for _ in range(n_copies)]
        for W, Wp in zip(Ws, Wps):
            _lib.check(lib.vcb_op_pack_skinny_weights(_lib.ptr(W), N, K, _lib.ptr(Wp), st))
        No = N // 2 if sw else N
        Y = torch.empty(B, No, device=dev, dtype=torch.float16)
        res = {}
        for name in ("row_major", "packed", "row_major", "packed"):  # interleaved repeats: the second pair is reported
            it = [0]

            def call():
                i = it[0] % n_copies
                it[0] += 1
                if name == "packed":
                    return lib.vcb_op_skinny_gemm_packed_f16(_lib.ptr(X), K, _lib.ptr(Wps[i]), B, N, K, None, N, sw, _lib.ptr(Y), No, st)
                return lib.vcb_op_skinny_gemm_f16(_lib.ptr(X), K, _lib.ptr(Ws[i]), B, N, K, None, N, sw, _lib.ptr(Y), No, st)
            for _ in range(2 * n_copies):
                call()
            torch.cuda.synchronize()
            reps = 6 * n_copies
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                call()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / reps
        w = 1 if N == 32000 else 32
        for name, ms in res.items():
            tot[name][0] += N * K * 2 * w
            tot[name][1] += ms * w
        out[f"{N}x{K}"] = {k: round(N * K * 2 / v / 1e6) for k, v in res.items()}
        del Ws, Wps
        torch.cuda.empty_cache()
    out["step_GBs"] = {k: round(b / t / 1e6) for k, (b, t) in tot.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
