"""Decode-phase timing of the config-3 job (VCoder-DS 7B, B=8, 127 decode steps after a 1207-token prefill) for A/B runs of
environment-selected kernel variants: one process per variant (the switches are read once per process).
    VCB_SKINNY_PF=4 python tools/bench_decode.py [--reps 3]
Prints one JSON line: prefill ms, decode ms/step (median of reps), HBM GB/s of the step (weights + KV)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    from vcoder_b200 import synth
    from vcoder_b200.engine import Engine
    dev = torch.device("cuda", 0)
    g = synth.PRESETS["7b"].with_variant(synth.VARIANT_VCODER_DS)
    B, N = args.batch, args.new_tokens
    ids = synth.make_prompt_ids(g, B, seed=0)
    Tp = ids.shape[1] - 3 + 2 * g.n_patches
    eng = Engine(g, max_batch=B, max_seq_len=((Tp + N + 63) // 64) * 64)
    bench.bind_synthetic(eng, g, seed=0, device=dev)
    px = [synth.make_pixels(g, B, s).to(dev) for s in (1, 2, 3)]
    stream = torch.cuda.current_stream()
    pre_ms, dec_ms = [], []
    for r in range(args.reps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(stream)
        eng.prefill(ids, *px)
        ev[1].record(stream)
        out = eng.decode(N - 1)
        ev[2].record(stream)
        torch.cuda.synchronize()
        if r > 0:
            pre_ms.append(ev[0].elapsed_time(ev[1]))
            dec_ms.append(ev[1].elapsed_time(ev[2]) / (N - 1))
    pre_ms.sort(); dec_ms.sort()
    H, F, V, L = g.llm_hidden, g.llm_ffn, g.llm_vocab, g.llm_layers
    step_bytes = 2 * (L * (4 * H * H + 3 * F * H) + V * H) + B * (Tp + N // 2) * 2 * L * H * 2
    d = dec_ms[len(dec_ms) // 2]
    env = {k: v for k, v in os.environ.items() if k.startswith("VCB_")}
    print(json.dumps({"env": env, "prefill_ms": round(pre_ms[len(pre_ms) // 2], 3), "decode_ms_per_step": round(d, 4),
                      "decode_step_hbm_GBs": round(step_bytes / (d / 1e3) / 1e9, 1), "ids_checksum": int(out.sum().item())}))


if __name__ == "__main__":
    main()
