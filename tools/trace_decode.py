"""Device-side timeline of ONE decode step inside the production CUDA graph (vcb_trace_begin / vcb_trace_read).

    python tools/trace_decode.py [--preset 7b] [--batch 8] [--out gpurun_out/trace_decode.csv]

Every CTA of the weight-streaming GEMMs, the decode attention kernels and the RMSNorm launches stamps %globaltimer at
entry, after its grid-dependency wait and at its end.  Per launch the table shows
  gap_us    : first CTA released (min t_wait) minus the previous launch's last CTA end  (the dependency bubble)
  early     : share of this launch's CTAs that were resident BEFORE the previous launch finished (ring prefilled)
  body_us   : last end minus first release
and the step is summed up by kind.  This is evidence for DESIGN.md section 3.1, not a benchmark number.
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

KINDS = {0: "gemm", 1: "attn_split", 2: "attn_combine", 3: "rmsnorm"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="7b")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace_decode.csv"))
    ap.add_argument("--label", default="")
    args = ap.parse_args()
    import bench
    from vcoder_b200 import _lib, synth
    from vcoder_b200.engine import Engine
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = synth.PRESETS[args.preset].with_variant(synth.VARIANT_VCODER_DS)
    B, N = args.batch, args.new_tokens
    ids = synth.make_prompt_ids(g, B, seed=0)
    Tp = ids.shape[1] - 3 + 2 * g.n_patches
    eng = Engine(g, max_batch=B, max_seq_len=((Tp + N + 63) // 64) * 64)
    bench.bind_synthetic(eng, g, seed=0, device=dev)
    px = [synth.make_pixels(g, B, s).to(dev) for s in (1, 2, 3)]
    eng.prefill(ids, *px)
    eng.decode(N // 2)  # warm graph, context in the middle of the job
    torch.cuda.synchronize()
    _lib.check(lib.vcb_trace_begin(400000))
    eng.decode(10)  # new trace epoch: one eager step + capture + replays; the slots hold the LAST replay
    max_l = 4096
    launches = np.zeros((max_l, 4), dtype=np.int32)
    n = C.c_int32(0)
    stamps = np.zeros(400000 * 3, dtype=np.uint64)
    _lib.check(lib.vcb_trace_read(_lib.ptr(launches), max_l, C.byref(n), _lib.ptr(stamps), stamps.size))
    launches = launches[: n.value]
    half = n.value // 2
    cap = launches[half:]  # the captured step (the first half is the eager step that precedes the capture)
    rows = []
    prev_end = None
    t_first = None
    for tag, ncta, base, aux in cap:
        st = stamps[base * 3: (base + ncta) * 3].reshape(ncta, 3).astype(np.int64)
        ok = st[:, 2] > 0
        if not ok.any():
            continue
        st = st[ok]
        enter, wait, end = st[:, 0], st[:, 1], st[:, 2]
        if t_first is None:
            t_first = wait.min()
        row = {"kind": KINDS.get(int(tag), str(tag)), "aux": int(aux), "ctas": int(ncta),
               "t_release_us": (wait.min() - t_first) / 1e3, "body_us": (end.max() - wait.min()) / 1e3,
               "gap_us": None if prev_end is None else (wait.min() - prev_end) / 1e3,
               "early": None if prev_end is None else float((enter < prev_end).mean()),
               "release_spread_us": (np.percentile(wait, 99) - wait.min()) / 1e3,
               "cta_body_med_us": float(np.median(end - wait)) / 1e3}
        prev_end = end.max()
        rows.append(row)
    step_us = (prev_end - t_first) / 1e3
    by = {}
    for r in rows:
        k = r["kind"] + (f"_N{r['aux']}" if r["kind"] == "gemm" else "")
        d = by.setdefault(k, {"n": 0, "body_us": 0.0, "gap_us": 0.0})
        d["n"] += 1
        d["body_us"] += r["body_us"]
        d["gap_us"] += r["gap_us"] or 0.0
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(f"# tools/trace_decode.py {args.label} preset={args.preset} B={B}: one graph-replayed decode step, {len(rows)} traced launches, "
                f"{step_us:.1f} us from first release to last end\n")
        f.write("kind,aux,ctas,t_release_us,gap_us,early,release_spread_us,body_us,cta_body_med_us\n")
        for r in rows[: 9 * 3 + 2]:  # first three layers in full
            f.write(",".join(str(round(r[k], 3)) if isinstance(r[k], float) else str(r[k]) for k in
                             ("kind", "aux", "ctas", "t_release_us", "gap_us", "early", "release_spread_us", "body_us", "cta_body_med_us")) + "\n")
        f.write("# totals per kind over the step: kind,launches,sum_body_us,sum_gap_us,avg_body_us,avg_gap_us\n")
        for k, d in by.items():
            f.write(f"# {k},{d['n']},{d['body_us']:.1f},{d['gap_us']:.1f},{d['body_us'] / d['n']:.2f},{d['gap_us'] / d['n']:.2f}\n")
    print(json.dumps({"label": args.label, "step_us": round(step_us, 1),
                      "by_kind": {k: {"n": d["n"], "avg_body_us": round(d["body_us"] / d["n"], 2), "avg_gap_us": round(d["gap_us"] / d["n"], 2)}
                                  for k, d in by.items()}}))


if __name__ == "__main__":
    main()
