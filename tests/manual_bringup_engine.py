"""GPU bring-up of the op-level kernels and the engine against the CPU oracle / golden fixtures.
Each stage runs in a subprocess with a timeout so a trap or hang cannot take the rest down."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def err(a, b):
    import torch
    a = a.float().cpu(); b = b.float().cpu()
    d = (a - b).abs()
    return {"max_abs": float(d.max()), "ref_absmax": float(b.abs().max()), "rel_rms": float(d.pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-12)),
            "nan": int(torch.isnan(a).sum())}


def stage_ops():
    import ctypes as C
    import torch
    from vcoder_b200 import _lib
    lib = _lib.load()
    out = {}
    dev = "cuda"
    torch.manual_seed(0)
    st = C.c_void_p(0)
    # layernorm
    x = torch.randn(1000, 1024, device=dev).half(); w = (1 + 0.1 * torch.randn(1024, device=dev)).half(); b = (0.1 * torch.randn(1024, device=dev)).half()
    y = torch.empty_like(x)
    _lib.check(lib.vcb_op_layernorm_f16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), 1000, 1024, 1e-5, st))
    out["layernorm"] = err(y, torch.nn.functional.layer_norm(x.float(), (1024,), w.float(), b.float(), 1e-5))
    # rmsnorm
    x = torch.randn(77, 4096, device=dev).half(); w = (1 + 0.1 * torch.randn(4096, device=dev)).half()
    y = torch.empty_like(x)
    _lib.check(lib.vcb_op_rmsnorm_f16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), 77, 4096, 1e-5, st))
    xf = x.float()
    out["rmsnorm"] = err(y, w.float() * xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5))
    # skinny
    for (B, N, K, sw, res) in [(8, 4096, 4096, 0, 1), (8, 22016, 4096, 1, 0), (8, 4096, 11008, 0, 1), (3, 512, 256, 0, 0), (16, 32000, 4096, 0, 0), (32, 1024, 512, 1, 0)]:
        X = (torch.randn(B, K, device=dev) * 0.5).half(); W = (torch.randn(N, K, device=dev) * 0.03).half()
        R = (torch.randn(B, N, device=dev)).half() if res else None
        No = N // 2 if sw else N
        Y = torch.full((B, No), float("nan"), device=dev, dtype=torch.float16)
        _lib.check(lib.vcb_op_skinny_gemm_f16(_lib.ptr(X), K, _lib.ptr(W), B, N, K, _lib.ptr(R), N, sw, _lib.ptr(Y), No, st))
        ref = X.float() @ W.float().t()
        if sw:
            r = ref.view(B, N // 16, 2, 8)
            ref = (torch.nn.functional.silu(r[:, :, 0]) * r[:, :, 1]).reshape(B, N // 2)
        if R is not None:
            ref = ref + R.float()
        e = err(Y, ref)
        # timing
        for _ in range(3):
            lib.vcb_op_skinny_gemm_f16(_lib.ptr(X), K, _lib.ptr(W), B, N, K, _lib.ptr(R), N, sw, _lib.ptr(Y), No, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(5):
            big.zero_()  # flush L2
            e0.record()
            lib.vcb_op_skinny_gemm_f16(_lib.ptr(X), K, _lib.ptr(W), B, N, K, _lib.ptr(R if not sw else None), N, sw, _lib.ptr(Y), No, st)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        e["ms"] = ms; e["GBs"] = N * K * 2 / ms / 1e6
        out[f"skinny_{B}x{N}x{K}_sw{sw}"] = e
        del big
    # vit attention
    for (B, S, heads) in [(2, 577, 16), (1, 17, 2), (3, 130, 4)]:
        D = heads * 64
        qkv = (torch.randn(B * S, 3 * D, device=dev)).half()
        o = torch.full((B * S, D), float("nan"), device=dev, dtype=torch.float16)
        _lib.check(lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), st))
        q, k, v = [t.float().view(B, S, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
        ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B * S, D)
        out[f"attn_vit_{B}x{S}x{heads}"] = err(o, ref)
    B, S, heads = 24, 577, 16
    D = heads * 64
    qkv = (torch.randn(B * S, 3 * D, device=dev)).half(); o = torch.empty(B * S, D, device=dev, dtype=torch.float16)
    for _ in range(3):
        lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["attn_vit_time_24img"] = {"ms": ms, "us_per_image_layer": ms * 1000 / B, "GBs_alg": B * 4726784 / ms / 1e6, "TFLOPs": B * 1.3637e9 / ms / 1e9}
    return out


def stage_engine(preset, variant, n_new):
    import numpy as np
    import torch
    from oracle import vcoder_oracle as orc
    from vcoder_b200 import synth
    from vcoder_b200.engine import Engine
    g = synth.PRESETS[preset].with_variant(variant)
    sd = synth.make_state_dict(g, seed=0)
    B = 2
    ids = synth.make_prompt_ids(g, B, seed=0, n_before=5, n_after=4) if preset == "tiny" else synth.make_prompt_ids(g, B, seed=0)
    images = synth.make_pixels(g, B, 1); segs = synth.make_pixels(g, B, 2)
    depths = synth.make_pixels(g, B, 3) if variant == 1 else None
    eng = Engine(g, max_batch=B, max_seq_len=ids.shape[1] + 3 * g.n_patches + n_new + 8)
    eng.bind(sd)
    out = {}
    tf = eng.tower_forward(images)
    torch.cuda.synchronize()
    out["tower"] = err(tf, orc.clip_tower_forward(images, sd, g))
    out["encode_img"] = err(eng.encode(0, images), orc.encode_images(images, sd, g))
    out["encode_seg"] = err(eng.encode(1, segs), orc.encode_seg_images(segs, sd, g))
    pre = eng.prefill(ids, images, segs, depths, want_logits=True, want_embeds=True)
    torch.cuda.synchronize()
    x_ref = orc.prefill_embeds(ids, sd, g, images, segs, depths)
    out["spliced_len"] = [pre["spliced_len"], x_ref.shape[1]]
    out["embeds"] = err(pre["embeds"], x_ref)
    print("RESULT " + json.dumps(out), flush=True)
    ids_ref, trace_ref = orc.greedy_generate(ids, sd, g, images, segs, depths, n_new, return_logits=True)
    out["prefill_logits"] = err(pre["logits"], trace_ref[0])
    dec_ids, dec_logits = eng.decode(n_new - 1, want_logits=True)
    torch.cuda.synchronize()
    got = torch.cat([pre["next_ids"][:, None].cpu(), dec_ids.cpu()], dim=1).long()
    out["ids_got"] = got.tolist(); out["ids_ref"] = ids_ref.tolist()
    for s in range(1, n_new):
        out[f"decode_logits_{s}"] = err(dec_logits[s - 1], trace_ref[s])
    # graph path + e2e generate
    print("RESULT " + json.dumps(out), flush=True)
    gen = eng.generate(ids, images, segs, depths, max_new_tokens=n_new)
    out["ids_generate"] = gen.tolist()
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        name = sys.argv[1]
        if name == "ops":
            r = stage_ops()
        else:
            r = stage_engine(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        print("RESULT " + json.dumps(r))
        sys.exit(0)
    for args in (["ops"], ["engine", "tiny", "1", "6"], ["engine", "tiny", "0", "6"], ["engine", "vit_l", "1", "4"]):
        try:
            r = subprocess.run([sys.executable, __file__, *args], capture_output=True, text=True, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                res = json.loads(lines[-1][7:])
                print("==", args)
                for k, v in res.items():
                    print("   ", k, v)
            else:
                print("==", args, "FAILED rc=%d" % r.returncode, r.stdout[-1500:], r.stderr[-3000:])
        except subprocess.TimeoutExpired:
            print("==", args, "TIMEOUT")
        sys.stdout.flush()
