"""GPU bring-up for the tcgen05 GEMM: correctness vs an fp32 matmul of the same fp16 inputs, then timing.
Run on the GPU box:  python tools/bringup_gemm.py  (each case in a subprocess so a trap cannot take the rest down).
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_case(M, N, K, act, bias, res, bn, time_it):
    import torch

    lib = ctypes.CDLL(os.path.join(ROOT, "vcoder_b200", "libvcoder_b200.so"))
    lib.vcb_last_error.restype = ctypes.c_char_p
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = (torch.randn(N, device=dev) * 0.1).half() if bias else None
    R = (torch.randn(M, N, device=dev) * 0.5).half() if res else None
    Nout = N // 2 if act == 3 else N
    C = torch.full((M, Nout), float("nan"), device=dev, dtype=torch.float16)
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)

    def call():
        return lib.vcb_op_gemm_f16(vp(A), K, vp(W), K, M, N, K, vp(b), vp(R), N, act, vp(C), Nout, bn, ctypes.c_void_p(0))

    rc = call()
    torch.cuda.synchronize()
    if rc != 0:
        return {"rc": rc, "err": lib.vcb_last_error().decode()}
    ref = A.float() @ W.float().t()
    if b is not None:
        ref = ref + b.float()
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    elif act == 3:
        r = ref.view(M, N // 16, 2, 8)
        ref = (torch.nn.functional.silu(r[:, :, 0]) * r[:, :, 1]).reshape(M, N // 2)
    if R is not None:
        ref = ref + R.float()
    diff = (C.float() - ref).abs()
    out = {
        "rc": 0,
        "max_abs": float(diff.max()),
        "nan": int(torch.isnan(C.float()).sum()),
        "ref_absmax": float(ref.abs().max()),
        "mean_abs": float(diff.mean()),
    }
    if out["max_abs"] > 0.05 or out["nan"]:
        bad = (diff > 0.05) | torch.isnan(C.float())
        idx = bad.nonzero()[:6].tolist()
        out["bad_count"] = int(bad.sum())
        out["bad_first"] = [(i, j, float(C[i, j]), float(ref[i, j])) for i, j in idx]
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        out["bad_rows_minmax"] = [int(rows.min()), int(rows.max()), int(rows.numel())]
        out["bad_cols_minmax"] = [int(cols.min()), int(cols.max()), int(cols.numel())]
    if time_it:
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out["ms"] = ms
        out["tflops"] = 2.0 * M * N * K / ms / 1e9
    return out


CASES = [
    # M, N, K, act, bias, res, bn, time
    (128, 128, 64, 0, 0, 0, 128, 0),
    (128, 128, 256, 0, 0, 0, 128, 0),
    (128, 256, 256, 0, 0, 0, 256, 0),
    (300, 384, 640, 0, 1, 0, 128, 0),
    (300, 512, 640, 1, 1, 1, 256, 0),
    (1000, 1024, 1024, 2, 1, 0, 0, 0),
    (1000, 1024, 1024, 3, 0, 0, 128, 0),
    (4616, 3072, 1024, 0, 1, 0, 128, 1),
    (13848, 3072, 1024, 0, 1, 0, 128, 1),
    (13848, 3072, 1024, 0, 1, 0, 256, 1),
    (13848, 4096, 1024, 1, 1, 0, 256, 1),
    (13848, 1024, 4096, 0, 1, 1, 128, 1),
    (13848, 1024, 4096, 0, 1, 1, 256, 1),
    (9648, 12288, 4096, 0, 0, 0, 256, 1),
    (9648, 22016, 4096, 3, 0, 0, 256, 1),
    (9648, 4096, 11008, 0, 0, 1, 256, 1),
    (9648, 4096, 11008, 0, 0, 1, 128, 1),
]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        args = [int(x) for x in sys.argv[1:]]
        print(json.dumps(run_case(*args)))
        sys.exit(0)
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, *map(str, c)], capture_output=True, text=True, timeout=180)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("NO OUTPUT rc=%d %s" % (r.returncode, r.stderr[-400:]))
        except subprocess.TimeoutExpired:
            line = "TIMEOUT"
        print(c, line, flush=True)
