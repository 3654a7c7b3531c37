"""Micro-benchmark of the skinny GEMM variants at the decode shapes (L2 flushed between launches)."""
import ctypes as C
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(12288, 4096, 0, 0), (4096, 4096, 0, 0), (22016, 4096, 1, 0), (4096, 11008, 0, 0), (32000, 4096, 0, 0)]


def run():
    import torch
    from vcoder_b200 import _lib
    lib = _lib.load()
    dev = "cuda"
    st = C.c_void_p(0)
    B = 8
    # L2 state as in production: every launch streams a DIFFERENT weight buffer (clean lines of other layers in L2,
    # no dirty write-back traffic from a memset-style flush)
    out = {}
    tot_b, tot_t = 0, 0
    for (N, K, sw, norm) in SHAPES:
        X = (torch.randn(B, K, device=dev) * 0.5).half()
        n_copies = max(3, int(600e6 // (N * K * 2)) + 1)
        Ws = [(torch.randn(N, K, device=dev) * 0.03).half() for _ in range(n_copies)]
        wn = torch.ones(K, device=dev).half()
        it = [0]
        No = N // 2 if sw else N
        Y = torch.empty(B, No, device=dev, dtype=torch.float16)

        def call():
            W = Ws[it[0] % n_copies]
            it[0] += 1
            if norm:
                return lib.vcb_op_skinny_rmsnorm_gemm_f16(_lib.ptr(X), K, _lib.ptr(wn), 1e-5, _lib.ptr(W), B, N, K, None, N, sw, _lib.ptr(Y), No, st)
            return lib.vcb_op_skinny_gemm_f16(_lib.ptr(X), K, _lib.ptr(W), B, N, K, None, N, sw, _lib.ptr(Y), No, st)
        for _ in range(2):
            call()
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3 * n_copies):
            call()
        torch.cuda.synchronize()
        reps = 4 * n_copies
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[f"{N}x{K}"] = round(N * K * 2 / ms / 1e6)
        tot_b += N * K * 2 * (1 if N == 32000 else 32)
        tot_t += ms * (1 if N == 32000 else 32)
    out["step_GBs"] = round(tot_b / tot_t / 1e6)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print("RESULT " + json.dumps(run()))
        sys.exit(0)
    for warps, mode, db, pf in itertools.product([0, 4, 8, 16], [0, 1], [0, 1], [0]):
        env = dict(os.environ, VCB_SKINNY_WARPS=str(warps), VCB_SKINNY_MODE=str(mode), VCB_SKINNY_DB=str(db), VCB_SKINNY_PREFETCH=str(pf))
        r = subprocess.run([sys.executable, __file__, "run"], capture_output=True, text=True, timeout=300, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        print(f"warps={warps} mode={mode} db={db} pf={pf}:", lines[-1][7:] if lines else r.stderr[-300:], flush=True)
