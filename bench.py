#!/usr/bin/env python
"""bench.py — end-to-end tokens/sec (prefill + decode) of VCoder-DS LLaVA-1.5-7B, batch 8 per GPU, 128 new tokens,
greedy, synthetic 336x336 images and seeded random-init weights (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one full generate job over one batch: 3x8 images through the shared CLIP ViT-L/14@336 tower,
projectors, splice, 32-layer prefill over T' = 1208 tokens and 127 greedy decode steps.
  value : tokens/s with pixels already resident in HBM (prefill + decode launched through the C ABI)
  e2e   : tokens/s through the reference-facing call (`Engine.generate` -> vcb_generate) with HOST buffers: pinned
          pixels/ids copied H2D and the generated ids copied D2H inside the timed region
  --impl reference : the oracle port of the reference's CPU path (fp32 torch-on-CPU, all host threads), timed on a
          bounded sample of the same workload and scaled to tokens/s (the full job is ~10 minutes of CPU).
N > 1 (torchrun): every rank runs its own replica on its own batch of 8 (weak scaling); the generated ids are joined
with ONE NCCL all-gather inside the timed region; time = max over ranks (CUDA events).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


# ------------------------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        # median over the upper half of samples = "under load" (idle gaps between steps pull the clock down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, bounded sample scaled to the full job
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample_tokens_per_s(B, n_new, repeats=1):
    """Times, with every host thread, one image through the ViT-L tower + projector, one LLaMA-7B layer of prefill
    (B=1, T'=1208), one layer of a decode step (B=8, ctx ~1270) and the lm_head, then scales by the job's counts:
      total = 3B*t_img + B*32*t_prefill_layer + (n_new-1)*(32*t_decode_layer + t_lm_head) + t_lm_head."""
    from oracle import vcoder_oracle as orc
    from vcoder_b200 import synth

    g = synth.PRESETS["7b"]
    n_threads = pick_threads()
    Tp = 55 + 3 - 3 + 2 * g.n_patches  # 58 ids incl. 3 specials -> 55 text + 1152 visual
    H, dh = g.llm_hidden, 128
    gen = torch.Generator().manual_seed(0)
    # --- tower + projector on one image
    sd_v = synth.make_state_dict(g, seed=0, only_prefix="model.vision_tower.")
    sd_v.update(synth.make_state_dict(g, seed=0, only_prefix="model.mm_projector."))
    sd_v = {k: v.float() for k, v in sd_v.items()}
    px = synth.make_pixels(g, 1, 1).float()
    t_img = []
    for _ in range(1 + repeats):
        t0 = time.perf_counter()
        f = orc.clip_tower_forward(px, sd_v, g)
        orc.projector(f, sd_v, "mm_projector")
        t_img.append(time.perf_counter() - t0)
    del sd_v
    # --- one decoder layer
    sd_l = {k: v.float() for k, v in synth.make_state_dict(g, seed=0, only_prefix="model.layers.0.").items()}
    x = torch.randn(1, Tp, H, generator=gen)
    t_pl = []
    for _ in range(1 + repeats):
        cache = orc.KVCache(1)
        t0 = time.perf_counter()
        orc.llama_layer(x, sd_l, 0, g, cache, 0)
        t_pl.append(time.perf_counter() - t0)
    xd = torch.randn(B, 1, H, generator=gen)
    ctx = Tp + n_new // 2
    t_dl = []
    for _ in range(2 + 2 * repeats):
        cache = orc.KVCache(1)
        cache.k[0] = torch.randn(B, g.llm_heads, ctx, dh, generator=gen)
        cache.v[0] = torch.randn(B, g.llm_heads, ctx, dh, generator=gen)
        t0 = time.perf_counter()
        orc.llama_layer(xd, sd_l, 0, g, cache, ctx)
        t_dl.append(time.perf_counter() - t0)
    del sd_l
    w_lm = torch.randn(g.llm_vocab, H, generator=gen) * 0.02
    xl = torch.randn(B, H, generator=gen)
    t_lm = []
    for _ in range(2 + repeats):
        t0 = time.perf_counter()
        torch.nn.functional.linear(xl, w_lm)
        t_lm.append(time.perf_counter() - t0)
    ti, tp, td, tl = min(t_img[1:]), min(t_pl[1:]), min(t_dl[1:]), min(t_lm[1:])
    total = 3 * B * ti + B * g.llm_layers * tp + (n_new - 1) * (g.llm_layers * td + tl) + tl
    cpu_work = sum(t_img) + sum(t_pl) + sum(t_dl) + sum(t_lm)
    return {
        "value": B * n_new / total, "unit": "tokens/s", "cores": n_threads, "kind": "port", "host_cores": os.cpu_count(),
        "sample": (f"oracle fp32 torch-CPU: 1 image ViT-L tower+projector {ti:.2f}s, 1 LLaMA-7B layer prefill B=1 T'={Tp} "
                   f"{tp:.2f}s, 1 layer decode step B={B} ctx={ctx} {td * 1e3:.0f}ms, lm_head {tl * 1e3:.0f}ms; scaled to the "
                   f"full job ({total:.0f}s); {cpu_work:.0f}s of CPU work sampled"),
        "job_seconds_est": total,
    }


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "vcoder_llava"))


def pick_threads():
    """All host threads — unless oversubscription makes torch slower (seen on the 128-core GPU boxes): probe with one
    LLaMA-sized matmul and keep the fastest of {all cores, 64, 32}."""
    cand = sorted({os.cpu_count(), min(os.cpu_count(), 64), min(os.cpu_count(), 32)}, reverse=True)
    a = torch.randn(1207, 4096)
    w = torch.randn(11008, 4096)
    best_n, best_t = cand[0], None
    for n in cand:
        torch.set_num_threads(n)
        torch.nn.functional.linear(a, w)
        t0 = time.perf_counter()
        torch.nn.functional.linear(a, w)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    return best_n


class ReferenceCpuJob:
    """The UNMODIFIED reference package (pip-installed from /root/reference into baseline/_ref, pure Python) on the host
    cores: its own VCoderDSLlavaLlamaForCausalLM / CLIPVisionTower / prepare_inputs_labels_for_multimodal on top of the
    installed transformers, fp32 (its CPU dtype), random-init at the 7B width.  One step = one bounded sample of the
    headline job: ALL three towers + projectors + splice for the full batch, prefill and `dec_steps` greedy decode steps
    through a decoder truncated to `layers` of the 32 layers.  Decoder layers are identical, so the full job is
    t_towers_splice + (32/layers) * t_prefill_layers + t_lm_head + (N-1) * ((32/layers) * t_dec_layers + t_dec_lm_head);
    the factors go into the JSON line, ms_per_step is the TRUE wall time of the sample.

    The greedy loop is route (ii) of SURVEY 8(c): under transformers >= 4.36 the reference's generate(use_cache=True)
    cannot run (prepare_inputs_for_generation truncates the prompt against a DynamicCache, vcoder_ds_llava_llama.py:123),
    so the loop calls the same modules generate() would: model.model(...) and lm_head."""

    def __init__(self, preset, batch, layers=2, dec_steps=8, n_steps_planned=1, budget_s=None):
        import tempfile
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from transformers import AutoConfig, AutoModelForCausalLM
        # the 6-line compatibility shim of SURVEY Appendix B ('llava' is a registered model_type since transformers 4.36)
        o1, o2 = AutoConfig.register, AutoModelForCausalLM.register
        AutoConfig.register = staticmethod(lambda mt, cfg, exist_ok=False: o1(mt, cfg, exist_ok=True))
        AutoModelForCausalLM.register = classmethod(lambda cls, cc, mc, exist_ok=False: o2(cc, mc, exist_ok=True))
        from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel
        from vcoder_llava.model import VCoderDSLlavaLlamaForCausalLM
        from vcoder_llava.model.language_model.vcoder_ds_llava_llama import VCoderDSLlavaConfig
        from vcoder_b200 import synth

        self.threads = pick_threads()
        g = synth.PRESETS[preset].with_variant(synth.VARIANT_VCODER_DS)
        self.g, self.B, self.layers, self.dec_steps = g, batch, layers, dec_steps
        torch.manual_seed(0)
        self.tmp = tempfile.TemporaryDirectory()
        vcfg = CLIPVisionConfig(hidden_size=g.vit_hidden, intermediate_size=g.vit_ffn, num_hidden_layers=g.vit_layers,
                                num_attention_heads=g.vit_heads, image_size=g.vit_image_size, patch_size=g.vit_patch,
                                hidden_act="quick_gelu", layer_norm_eps=g.vit_ln_eps, projection_dim=64)
        CLIPVisionModel(vcfg).save_pretrained(self.tmp.name)
        CLIPImageProcessor(size={"shortest_edge": g.vit_image_size}, crop_size=g.vit_image_size).save_pretrained(self.tmp.name)
        cfg = VCoderDSLlavaConfig(
            hidden_size=g.llm_hidden, intermediate_size=g.llm_ffn, num_hidden_layers=layers, num_attention_heads=g.llm_heads,
            num_key_value_heads=g.llm_heads, vocab_size=g.llm_vocab, rms_norm_eps=g.llm_rms_eps, max_position_embeddings=4096,
            pad_token_id=0, bos_token_id=1, eos_token_id=2, mm_vision_tower=self.tmp.name, mm_vision_select_layer=g.vit_select_layer,
            mm_vision_select_feature="patch", mm_projector_type="mlp2x_gelu", mm_hidden_size=g.vit_hidden,
            seg_mm_projector_type="mlp2x_gelu", seg_mm_hidden_size=g.vit_hidden, depth_mm_projector_type="mlp2x_gelu",
            depth_mm_hidden_size=g.vit_hidden, mm_vcoder_lm_emb=True, use_mm2_proj=True)
        self.model = VCoderDSLlavaLlamaForCausalLM(cfg).eval().float()
        self.model.get_vision_tower().load_model()
        self.model.float()
        # bounded sample: the run of `n_steps_planned` steps must end within a few minutes on whatever host this is.
        # One image through the tower calibrates the host; if the full-batch sample does not fit the budget, the sample
        # covers the first B_s sequences of the batch (towers / splice / prefill are per-sequence work: scaled by B/B_s;
        # a decode step at B_s is taken AS the B-sequence step — weight-streaming-bound, a bound that favours the reference)
        budget_s = float(os.environ.get("VCB_REF_BUDGET_S", "300")) if budget_s is None else budget_s
        px1 = synth.make_pixels(g, 1, 1).float()
        with torch.no_grad():
            self.model.get_vision_tower()(px1)
            t0 = time.perf_counter()
            self.model.get_vision_tower()(px1)
            t_img = time.perf_counter() - t0
        self.Bs = batch
        while self.Bs > 1 and n_steps_planned * (3 * self.Bs * t_img * 1.7) > budget_s:
            self.Bs //= 2
        self.t_img_probe = t_img
        self.ids = synth.make_prompt_ids(g, batch, seed=0)[: self.Bs]
        self.px = [synth.make_pixels(g, batch, s).float()[: self.Bs] for s in (1, 2, 3)]

    @torch.no_grad()
    def step(self):
        m, t = self.model, time.perf_counter
        t0 = t()
        mask = torch.ones_like(self.ids)
        _, _, _, embeds, _ = m.prepare_inputs_labels_for_multimodal(self.ids, mask, None, None, *self.px)
        t1 = t()
        o = m.model(inputs_embeds=embeds, use_cache=True)
        t2 = t()
        logits = m.lm_head(o.last_hidden_state)  # every position, as the reference's forward does (vcoder_ds_llava_llama.py:92-93)
        nxt = logits[:, -1].argmax(-1)
        t3 = t()
        past = o.past_key_values
        t_layers = t_lm = 0.0
        for _ in range(self.dec_steps):
            a = t()
            o = m.model(input_ids=nxt[:, None], past_key_values=past, use_cache=True)
            b = t()
            nxt = m.lm_head(o.last_hidden_state)[:, -1].argmax(-1)
            c = t()
            past = o.past_key_values
            t_layers += b - a
            t_lm += c - b
        return {"wall": t() - t0, "towers_splice": t1 - t0, "prefill_layers": t2 - t1, "prefill_lm_head": t3 - t2,
                "dec_layers": t_layers / self.dec_steps, "dec_lm_head": t_lm / self.dec_steps, "spliced_len": embeds.shape[1]}

    def job_seconds(self, r, n_new):
        f = self.g.llm_layers / self.layers
        fb = self.B / self.Bs
        return (fb * (r["towers_splice"] + f * r["prefill_layers"] + r["prefill_lm_head"])
                + (n_new - 1) * (f * r["dec_layers"] + r["dec_lm_head"]))

    def describe(self, r, n_new):
        f = self.g.llm_layers / self.layers
        return (f"unmodified reference package (vcoder_llava from baseline/_ref, transformers {__import__('transformers').__version__}, "
                f"fp32, {self.threads} threads): {self.Bs} of the {self.B} sequences: 3x{self.Bs} images through CLIPVisionTower + projectors + splice "
                f"{r['towers_splice']:.2f}s, prefill B={self.Bs} T'={int(r['spliced_len'])} through {self.layers} of {self.g.llm_layers} layers "
                f"{r['prefill_layers']:.2f}s, lm_head over all positions {r['prefill_lm_head']:.2f}s, {self.dec_steps} decode steps "
                f"({r['dec_layers'] * 1e3:.0f} ms layers + {r['dec_lm_head'] * 1e3:.0f} ms lm_head each); layer-linear factor {f:g}, batch factor "
                f"{self.B / self.Bs:g} on towers/prefill (decode step at B={self.Bs} taken as the B={self.B} step), "
                f"{n_new - 1} decode steps -> full job {self.job_seconds(r, n_new):.0f}s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not reference_available():
        # the oracle always exists: fall back to the port (states so in `kind`)
        vals, last = [], None
        t_steps = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            last = cpu_sample_tokens_per_s(args.batch, args.new_tokens, repeats=2)
            if i >= args.warmup:
                vals.append(last["value"])
                t_steps.append(time.perf_counter() - t0)
        v = sum(vals) / len(vals)
        last["value"] = v
        ms = 1e3 * sum(t_steps) / len(t_steps)
    else:
        job = ReferenceCpuJob(args.preset, args.batch, n_steps_planned=args.warmup + args.steps)
        rs = []
        for i in range(args.warmup + args.steps):
            r = job.step()
            if i >= args.warmup:
                rs.append(r)
        mean = {k: sum(x[k] for x in rs) / len(rs) for k in rs[0]}
        est = job.job_seconds(mean, args.new_tokens)
        v = args.batch * args.new_tokens / est
        ms = 1e3 * mean["wall"]
        last = {"value": v, "unit": "tokens/s", "cores": job.threads, "kind": "reference", "host_cores": os.cpu_count(),
                "sample": job.describe(mean, args.new_tokens), "job_seconds_est": est,
                "extrapolation": {"sample_seconds": mean["wall"], "job_seconds_est": est, "factor": est / mean["wall"],
                                  "layers_sampled": job.layers, "layers": job.g.llm_layers, "sequences_sampled": job.Bs,
                                  "sequences": job.B, "decode_steps_sampled": job.dec_steps,
                                  "decode_steps": args.new_tokens - 1}}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": last, "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "ms_per_step = wall time of one bounded sample of the job (see cpu_baseline.sample); value = tokens/s of the "
                "full job extrapolated from it (cpu_baseline.extrapolation)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
METRIC = "end-to-end tokens/sec (prefill+decode) VCoder-DS-7B bs8"


def workload_config(args, world):
    if getattr(args, "variant", "ds") == "vcoder":
        which = ("BASELINE.json configs[1]" if (args.preset, args.batch, args.new_tokens) == ("7b", 1, 32) else "non-headline shape")
        return {"workload": f"VCoder LLaVA-1.5-{args.preset.upper()}, RGB+seg towers (CLIP ViT-L/14@336), batch {args.batch} per GPU, "
                            f"prompt 57 ids -> T'=1207 spliced tokens, {args.new_tokens} new tokens, greedy ({which})",
                "global_batch": args.batch * world, "new_tokens": args.new_tokens, "parallelism": f"dp{world} (replicas, one NCCL all-gather of ids)",
                "l2": "inputs larger than L2: 13.5 GB of weights streamed per decode step vs 126 MB L2; no explicit flush",
                "weights": "seeded random-init fp16 (HF std 0.02)"}
    which = ("BASELINE.json configs[2]" if (args.preset, args.batch, args.new_tokens) == ("7b", 8, 128)
             else "BASELINE.json configs[3], one GPU's slice of the batch-32 job" if (args.preset, args.batch, args.new_tokens) == ("13b", 4, 256)
             else "non-headline shape")
    return {"workload": f"VCoder-DS LLaVA-1.5-{args.preset.upper()}, RGB+seg+depth towers (CLIP ViT-L/14@336), batch {args.batch} "
                        f"per GPU, prompt 58 ids -> T'=1207 spliced tokens, {args.new_tokens} new tokens, greedy ({which})",
            "global_batch": args.batch * world, "new_tokens": args.new_tokens, "parallelism": f"dp{world} (replicas, one NCCL all-gather of ids)",
            "l2": f"inputs larger than L2: {'13.5' if args.preset == '7b' else '26'} GB of weights + KV streamed per decode step vs 126 MB L2; no explicit flush",
            "weights": "seeded random-init fp16 (HF std 0.02)"}


def bind_synthetic(eng, g, seed, device):
    """Generate each tensor on the device and bind it immediately (keeps peak memory at one tensor over the model)."""
    from vcoder_b200 import synth
    specs = synth.weight_specs(g)
    for name, shape, kind, std in specs:
        gen = torch.Generator(device=device)
        gen.manual_seed(synth._seed_for(name, seed))
        t = torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        t = (1.0 + std * t) if kind == "ln_weight" else (std * t)
        eng.bind({name: t.to(torch.float16)})
        del t


def run_b200(args):
    import torch.distributed as dist
    from vcoder_b200 import _lib, synth
    from vcoder_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a GPU: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = None
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner on fd 1 from native code (NCCL_DEBUG_FILE does not
        # catch it), so fd 1 points at stderr for the whole run and the line is written to the saved descriptor at the end
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    ds = args.variant == "ds"
    g = synth.PRESETS[args.preset].with_variant(synth.VARIANT_VCODER_DS if ds else synth.VARIANT_VCODER)
    B, N = args.batch, args.new_tokens
    ids = synth.make_prompt_ids(g, B, seed=rank)  # [B, 58] (DS: IMAGE, DEPTH, SEG) / [B, 57] (VCoder: IMAGE, SEG)
    n_spec = 3 if ds else 2
    Tp = ids.shape[1] - n_spec + 2 * g.n_patches
    eng = Engine(g, max_batch=B, max_seq_len=((Tp + N + 63) // 64) * 64, device=local)
    bind_synthetic(eng, g, seed=0, device=dev)
    torch.cuda.synchronize()
    host_px = [synth.make_pixels(g, B, 10 * rank + s).pin_memory() for s in (1, 2, 3)]
    if not ds:
        host_px[2] = None  # VCoder (BASELINE configs[1]): RGB + seg towers only
    ids_pinned = ids.pin_memory()
    dev_px = [p.to(dev) if p is not None else None for p in host_px]
    gathered = torch.empty(world * B, N, dtype=torch.int32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream()

    def step_resident(collective=True):
        pre = eng.prefill(ids, dev_px[0], dev_px[1], dev_px[2])
        dec = eng.decode(N - 1)
        out = torch.cat([pre["next_ids"][:, None], dec], dim=1)
        if world > 1 and collective:
            dist.all_gather_into_tensor(gathered, out)  # K14: the only collective on the path
        return out

    def step_e2e():
        out = eng.generate(ids_pinned, host_px[0], host_px[1], host_px[2], max_new_tokens=N)  # host in, host out
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.to(dev, torch.int32))
        return out

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = lib.vcb_launch_count()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.vcb_launch_count() - c0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    for _ in range(args.warmup):
        step_resident()
        step_e2e()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, launches = timed(step_resident, args.steps)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    tokens = world * B * N * args.steps
    value = tokens / (ms_res / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)

    # ---- live per-kernel timing of one more step (CUDA events around every launch, graph disabled)
    roofline, breakdown, phases, vit = None, None, None, None
    if rank == 0:
        import ctypes as C
        peaks = load_peaks()
        lib.vcb_profile_enable(1)
        step_resident(collective=False)  # rank 0 only: must not enter a collective
        ms = (C.c_float * 8)()
        nl = (C.c_uint64 * 8)()
        _lib.check(lib.vcb_profile_read(ms, nl, 8))
        lib.vcb_profile_enable(0)
        breakdown = {c: {"ms": round(float(ms[i]), 3), "launches": int(nl[i])} for i, c in enumerate(_lib.PROFILE_CATEGORIES)}
        # phase split in production mode (CUDA graph + PDL): towers+prefill vs decode, device-timed
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(stream)
        eng.prefill(ids, dev_px[0], dev_px[1], dev_px[2])
        ev[1].record(stream)
        eng.decode(N - 1)
        ev[2].record(stream)
        torch.cuda.synchronize()
        phases = {"towers_prefill_ms": round(ev[0].elapsed_time(ev[1]), 3), "decode_ms": round(ev[1].elapsed_time(ev[2]), 3),
                  "decode_ms_per_step": round(ev[1].elapsed_time(ev[2]) / (N - 1), 4)}
        step_bytes_all = 2 * (g.llm_layers * (3 * g.llm_hidden ** 2 + g.llm_hidden ** 2 + 3 * g.llm_ffn * g.llm_hidden) + g.llm_vocab * g.llm_hidden) \
            + B * (Tp + N // 2) * 2 * g.llm_layers * g.llm_hidden * 2
        phases["decode_step_hbm_GBs"] = round(step_bytes_all / (phases["decode_ms_per_step"] / 1e3) / 1e9, 1)
        tot = sum(v["ms"] for v in breakdown.values())
        for v in breakdown.values():
            v["share"] = round(v["ms"] / tot, 4) if tot else None
        # dominant kernel: the decode weight-streaming skinny GEMM (HBM-bound).  Algorithmic bytes per launch =
        # the weight matrix it streams once: per decode step 32 x (qkv 3H*H + o H*H + gate|up 2F*H + down H*F) + lm_head
        H, F, V, L = g.llm_hidden, g.llm_ffn, g.llm_vocab, g.llm_layers
        step_bytes = 2 * (L * (3 * H * H + H * H + 2 * F * H + H * F) + V * H)
        sk = breakdown["skinny_gemm"]
        total_bytes = step_bytes * (N - 1) + 2 * V * H  # + the prefill's last-position lm_head
        ach_events = total_bytes / (sk["ms"] / 1e3) / 1e9 if sk["ms"] else None
        # average launch duration, live: the step's 4L+1 weight-streaming launches on the bound weights, back to back
        # with PDL as in the decode graph, between two CUDA events on the launching stream (13.2 GB per round >> L2)
        rounds = 10
        ms_ws, n_ws = C.c_float(0), C.c_int32(0)
        _lib.check(lib.vcb_bench_weight_stream(eng.h, B, rounds, C.byref(ms_ws), C.byref(n_ws), C.c_void_p(stream.cuda_stream)))
        avg_launch_us = ms_ws.value * 1e3 / max(1, n_ws.value)
        bytes_per_launch = step_bytes * rounds / max(1, n_ws.value)
        ach = bytes_per_launch / (avg_launch_us * 1e-6) / 1e9 if avg_launch_us else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_skinny_traffic.json")
        if not os.path.exists(tp):
            tp = os.path.join(ROOT, "profiles", "skinny_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        roofline = {"kernel": "skinny_stream_kernel (TMA-fed decode weight streaming, balanced multi-tile CTAs: qkv/o/gate-up/down/lm_head)", "bound": "hbm",
                    "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (ach / peaks["hbm_gbs"]) if ach else None,
                    "traffic": traffic, "peak_source": peaks["source"],
                    "bytes_per_launch_avg": bytes_per_launch, "avg_launch_us": avg_launch_us, "launches": n_ws.value,
                    "how": f"{rounds} rounds of the decode step's {n_ws.value // rounds} weight-streaming launches (real weights, "
                           "PDL, one stream) between two CUDA events; algorithmic bytes = weight bytes, read once",
                    "achieved_with_per_launch_events": ach_events,
                    "note": "achieved_with_per_launch_events = same bytes / sum of per-launch event pairs inside one full "
                            "job (graph off, ~2 us event overhead per launch, no PDL overlap): a lower bound"}
        # secondary: the kernel the north star names (CLIP self-attention) and the three-tower throughput
        n_kinds = 3 if ds else 2
        n_img_layers = n_kinds * B * (g.vit_layers - 1)
        av = breakdown["attn_vit"]
        attn_us = av["ms"] * 1e3 / n_img_layers if av["ms"] else None
        attn_bytes = 4 * (g.n_patches + 1) * g.vit_hidden * 2
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        px_all = torch.cat([p for p in dev_px if p is not None], dim=0)
        eng.tower_forward(px_all)
        ev[0].record(stream)
        for _ in range(3):
            eng.tower_forward(px_all)
        ev[1].record(stream)
        torch.cuda.synchronize()
        tower_ms = ev[0].elapsed_time(ev[1]) / 3
        vit = {"images": n_kinds * B, "tower_ms": round(tower_ms, 3), "images_per_s": round(n_kinds * B / (tower_ms / 1e3), 1),
               "tower_tflops": round(n_kinds * B * 366e9 / (tower_ms / 1e3) / 1e12, 1),
               "attn_us_per_image_layer": round(attn_us, 3) if attn_us else None,
               "attn_hbm_frac": round(attn_bytes / (attn_us * 1e-6) / 1e9 / peaks["hbm_gbs"], 4) if attn_us else None,
               "attn_tflops": round(4 * (g.n_patches + 1) ** 2 * g.vit_hidden / (attn_us * 1e-6) / 1e12, 1) if attn_us else None,
               "note": "attention: 4,726,784 algorithmic bytes and 1.364 GFLOP per (image, layer) (SURVEY 8d), per-launch events"}
    # ---- ids of the timed (CUDA-graph) path must equal the eager launches' ids: a wrong decode-attention split, page
    #      table or graph capture would otherwise change neither a test nor the tokens/s
    ids_check = None
    if rank == 0:
        import hashlib
        g_ids = step_resident(collective=False).cpu()
        pre = eng.prefill(ids, dev_px[0], dev_px[1], dev_px[2])
        dec, _lg = eng.decode(N - 1, want_logits=True)  # logits requested => eager launches, no graph
        e_ids = torch.cat([pre["next_ids"][:, None], dec], dim=1).cpu()
        del _lg
        h_ids = eng.generate(ids_pinned, host_px[0], host_px[1], host_px[2], max_new_tokens=N).to(torch.int32)
        ids_check = {"ids_checksum": hashlib.sha256(g_ids.numpy().tobytes()).hexdigest()[:16],
                     "graph_equals_eager": bool(torch.equal(g_ids, e_ids)), "e2e_equals_resident": bool(torch.equal(h_ids, g_ids))}
        if not (ids_check["graph_equals_eager"] and ids_check["e2e_equals_resident"]):
            raise RuntimeError(f"bench: generated ids differ between the graph, eager and host-buffer paths: {ids_check}")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if reference_available():
            job = ReferenceCpuJob(args.preset, B, n_steps_planned=2, budget_s=40.0)  # bounded: ~10-30 s of CPU work
            job.step()
            r = job.step()
            est = job.job_seconds(r, N)
            cpu = {"value": B * N / est, "unit": "tokens/s", "cores": job.threads, "kind": "reference", "host_cores": os.cpu_count(),
                   "sample": job.describe(r, N), "job_seconds_est": est, "sample_seconds": r["wall"]}
            del job
        else:
            cpu = cpu_sample_tokens_per_s(B, N, repeats=20)  # ~15 s of CPU work on the GPU box
    # ---- second baseline: the HF library path the reference bottoms out in on a GPU (CLIPVisionModel +
    #      LlamaForCausalLM, eager attention, fp16, cuBLAS) on THIS GPU, same weights, with a token comparison
    library = None
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_hf_library as hfl
            library = hfl.measure(g, dev, B, N, reps=2, engine_ids=g_ids.long() if ids_check else None, ds=ds)
        except Exception as ex:  # a reported baseline must never take the bench line down
            library = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    if rank == 0:
        h2d = sum(p.numel() * 2 for p in host_px if p is not None) + ids.numel() * 8
        headline = (args.preset, args.batch, args.variant) == ("7b", 8, "ds")
        line = {
            "metric": METRIC if headline else f"end-to-end tokens/sec (prefill+decode) VCoder{'-DS' if ds else ''}-{args.preset.upper()} bs{args.batch}",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f16", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * N * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernel_breakdown": breakdown, "phases": phases, "vit": vit,
            "cpu_baseline": cpu, "library_baseline": library, "impl": "b200",
        }
        if ids_check:
            line.update(ids_check)
        if real_stdout is not None:
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="7b")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--variant", default="ds", choices=["ds", "vcoder"],
                    help="ds: VCoder-DS (RGB+seg+depth, the headline); vcoder: VCoder (RGB+seg; BASELINE configs[1] with --batch 1 --new-tokens 32)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fixed global batch split over the ranks (overrides --batch; reported as scaling=strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.global_batch > 0:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.global_batch % world != 0:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        args.batch = args.global_batch // world
        args.scaling = "strong"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
