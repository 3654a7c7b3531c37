"""BASELINE.json configs[4]: ViT-L/14@336 three-tower throughput sweep (RGB + seg + depth through the ONE shared tower,
SURVEY F1) — images/s, tower TFLOP/s and the attention kernel's share / HBM-yardstick fraction per batch size.

    python tools/bench_tower_sweep.py [--batches 1,2,4,8,16,32,64,128,256] [--reps 5]        (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_tower_sweep.py   (N replicas, no collective on the
        data path: the per-rank rates are summed, times are max over ranks)

`batch` = samples per GPU; every sample brings 3 images, so one vcb_tower_forward call sees 3*batch images (the handle's
vision workspace holds 96; larger calls run in chunks of 96 inside the library).
Pixels are resident in HBM; weights (582 MB) + activations of >= 3 images exceed nothing in particular, so the L2 is
flushed between timed calls by writing a 256 MB buffer.  Prints one JSON line per batch size and a final summary line.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from vcoder_b200 import _lib, synth
    from vcoder_b200.engine import Engine
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    g = synth.PRESETS["vit_l"]
    batches = [int(b) for b in args.batches.split(",")]
    eng = Engine(g, max_batch=min(32, max(batches)), max_seq_len=64, device=local)
    eng.bind({k: v.to(dev) for k, v in synth.make_state_dict(g, seed=0, only_prefix="model.vision_tower.").items()})
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    S, D, L = g.n_patches + 1, g.vit_hidden, g.vit_layers - 1
    attn_bytes = 4 * S * D * 2          # Q,K,V in + O out per (image, layer)
    flops_img = 366e9                   # useful FLOP per image through 23 layers
    rows = []
    for B in batches:
        n = 3 * B
        px = torch.cat([synth.make_pixels(g, B, s, device=str(dev)) for s in (1, 2, 3)], dim=0)
        for _ in range(3):
            eng.tower_forward(px)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for e0, e1 in ev:
            flush.zero_()
            e0.record(stream)
            eng.tower_forward(px)
            e1.record(stream)
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[len(ev) // 2]
        # attention share of one more call, per-launch events
        lib.vcb_profile_enable(1)
        eng.tower_forward(px)
        t_ms, t_n = (C.c_float * 8)(), (C.c_uint64 * 8)()
        _lib.check(lib.vcb_profile_read(t_ms, t_n, 8))
        lib.vcb_profile_enable(0)
        attn_ms, all_ms = float(t_ms[2]), sum(float(x) for x in t_ms)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        attn_us = attn_ms * 1e3 / (n * L)
        row = {"batch": B, "images": n * world, "n_gpus": world, "tower_ms": round(ms, 3),
               "images_per_s": round(n * world / (ms / 1e3), 1), "tower_tflops_per_gpu": round(n * flops_img / (ms / 1e3) / 1e12, 1),
               "attn_share": round(attn_ms / all_ms, 4) if all_ms else None, "attn_us_per_image_layer": round(attn_us, 3),
               "attn_hbm_GBs": round(attn_bytes / (attn_us * 1e-6) / 1e9, 1),
               "attn_hbm_frac": round(attn_bytes / (attn_us * 1e-6) / 1e9 / peaks["hbm_gbs"], 4)}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0:
        best = max(rows, key=lambda r: r["images_per_s"])
        print(json.dumps({"summary": "ViT-L/14@336 three-tower sweep", "n_gpus": world, "best": best,
                          "hbm_peak_GBs": peaks["hbm_gbs"], "l2": "256 MB flush write between timed calls"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
