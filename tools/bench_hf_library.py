"""Second baseline: the same job through the LIBRARY path the reference bottoms out in — HF transformers'
`CLIPVisionModel` and `LlamaForCausalLM` (eager attention) in fp16 on the GPU — with the same seeded synthetic weights.

    python tools/bench_hf_library.py [--preset 7b] [--batch 8] [--new-tokens 128] [--reps 3]           (GPU, fp16)
    python tools/bench_hf_library.py --preset tiny --device cpu --dtype fp32 --check-fixture            (CPU dry run)

It restates the reference's glue around those modules (vcoder_ds_llava_arch.py:106-314): three tower calls, tap
`hidden_states[-2][:, 1:]`, `mlp2x_gelu` projectors (depth through `seg_mm_projector`, F3), splice (the C++ planner of
libvcoder_b200 gives the row plan; the gather itself is torch), prefill with `inputs_embeds`, greedy loop with the HF KV
cache.  `vcoder_llava` itself cannot travel to the GPU box, so this is "HF library code + restated glue", not the reference
package; it is a reported baseline and a token-parity cross-check (`--compare-engine`), never part of the product path.
Prints one JSON line.  NOTE: written at the end of round 1 after the GPU budget was spent — the CPU dry run below is
validated against the reference-written fixture, the GPU/fp16/7B path has not been run yet.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_VT = "model.vision_tower.vision_tower."


def build_modules(g, device, dtype, seed=0):
    """HF modules on `device` filled tensor by tensor from the seeded generator (no CPU copy of a 7B model)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel, LlamaConfig, LlamaForCausalLM
    from vcoder_b200 import synth

    vcfg = CLIPVisionConfig(hidden_size=g.vit_hidden, intermediate_size=g.vit_ffn, num_hidden_layers=g.vit_layers,
                            num_attention_heads=g.vit_heads, image_size=g.vit_image_size, patch_size=g.vit_patch,
                            hidden_act="quick_gelu", layer_norm_eps=g.vit_ln_eps, projection_dim=64)
    lcfg = LlamaConfig(hidden_size=g.llm_hidden, intermediate_size=g.llm_ffn, num_hidden_layers=g.llm_layers,
                       num_attention_heads=g.llm_heads, num_key_value_heads=g.llm_heads, vocab_size=g.llm_vocab,
                       rms_norm_eps=g.llm_rms_eps, max_position_embeddings=4096, rope_theta=g.llm_rope_theta,
                       pad_token_id=0, bos_token_id=1, eos_token_id=2, tie_word_embeddings=False)
    vcfg._attn_implementation = "eager"
    lcfg._attn_implementation = "eager"
    with torch.device("meta"):
        tower = CLIPVisionModel(vcfg)
        llm = LlamaForCausalLM(lcfg)
    tower = tower.to_empty(device=device).to(dtype).eval()
    llm = llm.to_empty(device=device).to(dtype).eval()
    # non-persistent buffers are garbage after to_empty(): rebuild them the way the modules' constructors do
    emb = tower.vision_model.embeddings
    emb.position_ids = torch.arange(emb.num_positions, device=device).expand((1, -1))
    rot = llm.model.rotary_emb
    inv_freq, scaling = rot.rope_init_fn(rot.config, device) if hasattr(rot, "rope_init_fn") else rot.compute_default_rope_parameters(rot.config, device)
    rot.inv_freq = inv_freq.to(device)
    rot.original_inv_freq = rot.inv_freq
    rot.attention_scaling = scaling
    proj = {}
    tsd, lsd = dict(tower.named_parameters()), dict(llm.named_parameters())
    gen_dev = device if torch.device(device).type == "cuda" else "cpu"
    for name, shape, kind, std in synth.weight_specs(g):
        gen = torch.Generator(device=gen_dev)
        gen.manual_seed(synth._seed_for(name, seed))
        t = torch.randn(shape, generator=gen, device=gen_dev, dtype=torch.float32)
        t = ((1.0 + std * t) if kind == "ln_weight" else (std * t)).to(torch.float16).to(dtype)  # fp16-valued, as the engine binds
        if name.startswith(_VT):
            key = name[len(_VT):]
            if key in tsd:
                tsd[key].data.copy_(t)
        elif name in lsd:
            lsd[name].data.copy_(t)
        elif name.startswith("model.mm_projector.") or name.startswith("model.seg_mm_projector."):
            proj[name] = t.to(device)
        del t
    return tower, llm, proj


@torch.no_grad()
def encode(tower, proj, which, pixels, select_layer=-2):
    hs = tower(pixels, output_hidden_states=True).hidden_states[select_layer][:, 1:]
    p = "model.mm_projector." if which == "image" else "model.seg_mm_projector."  # depth -> seg projector (F3)
    h = torch.nn.functional.linear(hs, proj[p + "0.weight"], proj[p + "0.bias"])
    h = torch.nn.functional.gelu(h)
    return torch.nn.functional.linear(h, proj[p + "2.weight"], proj[p + "2.bias"])


@torch.no_grad()
def generate(g, tower, llm, proj, ids, images, segs, depths, n_new, keep_trace=True, ds=True):
    """-> (new ids [B, n_new] on the host, per-step logits list or None).  keep_trace=False: no per-step host copy."""
    from vcoder_b200 import synth
    from vcoder_b200.engine import splice_plan
    B = ids.shape[0]
    feats = [encode(tower, proj, "image", images), encode(tower, proj, "seg", segs)]
    if ds:
        feats.append(encode(tower, proj, "depth", depths))
    P = g.n_patches
    variant = synth.VARIANT_VCODER_DS if ds else synth.VARIANT_VCODER
    plan = torch.from_numpy(splice_plan(variant, ids.cpu().numpy(), P, True, [True] * B if ds else None)).to(ids.device).long()
    # feature buffer in the planner's row order: [kind][B][1 + P] (row 0 of every image = CLS, never referenced)
    H = feats[0].shape[-1]
    fbuf = torch.zeros(3, B, P + 1, H, dtype=feats[0].dtype, device=feats[0].device)
    for k in range(len(feats)):
        fbuf[k, :, 1:] = feats[k]
    fbuf = fbuf.view(-1, H)
    emb = llm.model.embed_tokens.weight
    x = torch.where((plan >= 0)[..., None], emb[plan.clamp(min=0)], fbuf[(-plan - 2).clamp(min=0)])
    out = llm.model(inputs_embeds=x, use_cache=True)
    past = out.past_key_values
    logits = llm.lm_head(out.last_hidden_state[:, -1])
    new, trace = [logits.argmax(-1)], ([logits.float().cpu()] if keep_trace else None)
    for _ in range(n_new - 1):
        out = llm.model(input_ids=new[-1][:, None], past_key_values=past, use_cache=True)
        past = out.past_key_values
        logits = llm.lm_head(out.last_hidden_state[:, -1])
        new.append(logits.argmax(-1))
        if keep_trace:
            trace.append(logits.float().cpu())
    return torch.stack(new, dim=1).cpu(), trace


def measure(g, dev, B, N, reps=2, engine_ids=None, seed=0, ds=True):
    """bench.py's `library_baseline`: the job through the HF library modules in fp16 on `dev` (1 warm-up + `reps` timed
    runs, wall clock with synchronize on both sides), and where its greedy ids first differ from `engine_ids` [B, N]."""
    from vcoder_b200 import synth
    tower, llm, proj = build_modules(g, str(dev), torch.float16, seed=seed)
    ids = synth.make_prompt_ids(g, B, seed=0).to(dev)
    px = [synth.make_pixels(g, B, s).to(dev, torch.float16) for s in (1, 2, 3)]
    new, trace = generate(g, tower, llm, proj, ids, *px, N, ds=ds)  # warm-up; its logits serve the token comparison
    times = []
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        new2, _ = generate(g, tower, llm, proj, ids, *px, N, keep_trace=False, ds=ds)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    times.sort()
    t = times[len(times) // 2]
    out = {"value": B * N / t, "unit": "tokens/s", "kind": "hf-fp16-eager", "ms_per_job": 1e3 * t, "reps": reps,
           "transformers": __import__("transformers").__version__, "torch": torch.__version__,
           "what": "transformers CLIPVisionModel + LlamaForCausalLM (eager attention, fp16, HF KV cache) with the reference's glue "
                   "restated in torch (tools/bench_hf_library.py), same seeded weights, same GPU"}
    if engine_ids is not None:
        same = (engine_ids.cpu() == new)
        first = [int((~same[b]).nonzero()[0]) if not bool(same[b].all()) else N for b in range(B)]
        margins = []
        for b in range(B):
            if first[b] < N:
                top2 = trace[first[b]][b].topk(2).values
                margins.append(round(float(top2[0] - top2[1]), 5))
        out["engine_tokens_equal_until"] = first
        out["hf_top2_margin_at_first_divergence"] = margins
    del tower, llm, proj
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="7b")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--check-fixture", action="store_true", help="tiny preset: compare with tests/golden/tiny_ds.npz")
    ap.add_argument("--compare-engine", action="store_true", help="also run libvcoder_b200 and report token agreement")
    args = ap.parse_args()
    from vcoder_b200 import synth
    g = synth.PRESETS[args.preset].with_variant(synth.VARIANT_VCODER_DS)
    dtype = torch.float16 if args.dtype == "fp16" else torch.float32
    dev = torch.device(args.device)
    tower, llm, proj = build_modules(g, args.device, dtype)
    B, N = args.batch, args.new_tokens
    if args.check_fixture:
        import numpy as np
        z = np.load(os.path.join(ROOT, "tests", "golden", "tiny_ds.npz"))
        ids = torch.from_numpy(z["ids"]).to(dev)
        B, N = ids.shape[0], z["greedy_ids"].shape[1]
    else:
        ids = synth.make_prompt_ids(g, B, seed=0).to(dev)
    px = [synth.make_pixels(g, B, s).to(dev, dtype) for s in (1, 2, 3)]
    times = []
    for r in range(args.reps + 1):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        new, trace = generate(g, tower, llm, proj, ids, *px, N)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if r > 0 or args.reps == 0:
            times.append(time.perf_counter() - t0)
    times.sort()
    t = times[len(times) // 2]
    line = {"impl": "hf-library", "metric": f"end-to-end tokens/sec (prefill+decode) VCoder-DS-{args.preset.upper()} bs{B}",
            "value": B * N / t, "unit": "tokens/s", "ms_per_job": 1e3 * t, "dtype": args.dtype, "device": args.device,
            "transformers": __import__("transformers").__version__, "attn_implementation": "eager"}
    if args.check_fixture:
        import numpy as np
        line["fixture_ids_equal"] = bool(np.array_equal(new.numpy(), z["greedy_ids"]))
        line["fixture_max_logit_err"] = float(max(np.abs(trace[s].numpy() - z["step_logits"][s]).max() for s in range(N)))
    if args.compare_engine:
        from vcoder_b200.engine import Engine
        eng = Engine(g, max_batch=B, max_seq_len=((ids.shape[1] + 2 * g.n_patches + N + 63) // 64) * 64)
        sys.path.insert(0, ROOT)
        import bench
        bench.bind_synthetic(eng, g, seed=0, device=dev)
        mine = eng.generate(ids.cpu(), *[p.half() for p in px], max_new_tokens=N)
        same = (mine == new)
        first_div = [int((~same[b]).nonzero()[0]) if not bool(same[b].all()) else N for b in range(B)]
        margins = []
        for b in range(B):
            if first_div[b] < N:
                top2 = trace[first_div[b]][b].topk(2).values
                margins.append(float(top2[0] - top2[1]))
        line["engine_tokens_equal_until"] = first_div
        line["hf_top2_margin_at_first_divergence"] = margins
    print(json.dumps(line))


if __name__ == "__main__":
    main()
