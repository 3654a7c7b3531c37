"""Bring-up of the tcgen05 ViT attention: compare against fp32 reference for several V-descriptor encodings."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    from vcoder_b200 import _lib
    lib = _lib.load()
    dev = "cuda"
    st = C.c_void_p(0)
    out = {}
    for (B, S, heads, qs) in [(1, 128, 1, 1.0), (1, 64, 1, 1.0), (2, 577, 16, 2.0), (1, 17, 2, 2.0), (3, 130, 4, 3.0), (1, 640, 2, 1.0)]:
        torch.manual_seed(S)
        D = heads * 64
        qkv = torch.randn(B * S, 3 * D, device=dev)
        qkv[:, :D] *= qs
        qkv = qkv.half()
        o = torch.full((B * S, D), float("nan"), device=dev, dtype=torch.float16)
        rc = lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), st)
        torch.cuda.synchronize()
        if rc != 0:
            out[f"{B}x{S}x{heads}"] = {"rc": rc, "err": lib.vcb_last_error().decode()}
            continue
        q, k, v = [t.float().view(B, S, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
        ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B * S, D)
        d = (o.float() - ref).abs()
        out[f"{B}x{S}x{heads}"] = {"max_abs": float(d.max()), "rel_rms": float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
                                  "nan": int(torch.isnan(o.float()).sum())}
    # timing at the production shape
    B, S, heads = 24, 577, 16
    D = heads * 64
    qkv = torch.randn(B * S, 3 * D, device=dev).half()
    o = torch.empty(B * S, D, device=dev, dtype=torch.float16)
    for _ in range(3):
        lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["time_24img"] = {"ms": ms, "us_per_image_layer": ms * 1000 / B, "GBs_alg": B * 4726784 / ms / 1e6, "TFLOPs": B * 1.3637e9 / ms / 1e9}
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print("RESULT " + json.dumps(run()))
        sys.exit(0)
    variants = {"default": None, "lbo1": "40004040,1", "swap": "40004001,40", "lbo128": "40004040,80", "sbo2048": "40004080,0"}
    for name, v in variants.items():
        env = dict(os.environ, VCB_VIT_ATTN="tc5")
        if v:
            env["VCB_VIT_VDESC"] = v
        try:
            r = subprocess.run([sys.executable, __file__, "run"], capture_output=True, text=True, timeout=180, env=env)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            print("==", name, v, lines[-1][7:] if lines else ("FAILED rc=%d %s" % (r.returncode, r.stderr[-600:])), flush=True)
        except subprocess.TimeoutExpired:
            print("==", name, "TIMEOUT", flush=True)
        ok = False
        if lines:
            res = json.loads(lines[-1][7:])
            ok = all(("rel_rms" in x and x["rel_rms"] < 2e-3 and x["nan"] == 0) for k, x in res.items() if k != "time_24img")
        if ok:
            print("VARIANT OK:", name)
            break
