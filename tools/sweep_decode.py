"""In-process A/B sweep of the weight-streaming kernel's ring shape per GEMM shape (vcb_tune_skinny) on the config-3 decode
phase (VCoder-DS 7B, B = 8, context 1207 -> 1334): one engine, one line of JSON per configuration.

    python tools/sweep_decode.py [--preset 7b] [--batch 8] [--configs "o=4x3;down=4x3|o=8x2"] [--reps 2]

A configuration is `name=WxD;name=WxD...` over the shapes qkv, o, gu, down, lm (W warps, D stages of 8 KiB per warp).  The
captured decode graph is rebuilt for every configuration (a trace epoch bump invalidates it).  Evidence for DESIGN.md 3.1.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

DEFAULT = ("base|o=4x3|o=4x4|o=8x2|o=2x3|down=4x3|down=4x4|down=8x2|down=8x3|qkv=4x3|qkv=8x2|gu=4x3|gu=8x2|lm=4x3|lm=8x2|"
           "o=4x3;down=4x3|o=8x2;down=8x3")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="7b")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--configs", default=DEFAULT)
    args = ap.parse_args()
    from vcoder_b200 import _lib, synth
    from vcoder_b200.engine import Engine
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = synth.PRESETS[args.preset].with_variant(synth.VARIANT_VCODER_DS)
    B, N = args.batch, args.new_tokens
    H, F, V, L = g.llm_hidden, g.llm_ffn, g.llm_vocab, g.llm_layers
    shapes = {"qkv": (3 * H, H), "o": (H, H), "gu": (2 * F, H), "down": (H, F), "lm": (V, H)}
    ids = synth.make_prompt_ids(g, B, seed=0)
    Tp = ids.shape[1] - 3 + 2 * g.n_patches
    eng = Engine(g, max_batch=B, max_seq_len=((Tp + N + 63) // 64) * 64)
    bench.bind_synthetic(eng, g, seed=0, device=dev)
    px = [synth.make_pixels(g, B, s).to(dev) for s in (1, 2, 3)]
    stream = torch.cuda.current_stream()
    step_bytes = 2 * (L * (4 * H * H + 3 * F * H) + V * H) + B * (Tp + N // 2) * 2 * L * H * 2

    def bump_graph():
        _lib.check(lib.vcb_trace_begin(1))
        la, n, st = np.zeros((1, 4), np.int32), C.c_int32(0), np.zeros(3, np.uint64)
        _lib.check(lib.vcb_trace_read(_lib.ptr(la), 1, C.byref(n), _lib.ptr(st), 3))

    for cfg in args.configs.split("|"):
        for name, (n_, k_) in shapes.items():
            _lib.check(lib.vcb_tune_skinny(n_, k_, 0, 0))
        if cfg != "base":
            for item in cfg.split(";"):
                name, wd = item.split("=")
                w, d = wd.split("x")
                _lib.check(lib.vcb_tune_skinny(shapes[name][0], shapes[name][1], int(w), int(d)))
        bump_graph()
        dec_ms = []
        out = None
        for r in range(args.reps + 1):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            eng.prefill(ids, *px)
            ev[0].record(stream)
            out = eng.decode(N - 1)
            ev[1].record(stream)
            torch.cuda.synchronize()
            if r > 0:
                dec_ms.append(ev[0].elapsed_time(ev[1]) / (N - 1))
        d = min(dec_ms)
        print(json.dumps({"config": cfg, "decode_ms_per_step": round(d, 4), "decode_step_hbm_GBs": round(step_bytes / (d / 1e3) / 1e9, 1),
                          "ids_checksum": int(out.sum().item())}), flush=True)


if __name__ == "__main__":
    main()
