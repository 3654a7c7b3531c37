"""`load_pretrained_model` — mirror of vcoder_llava/model/builder.py:25-154 for the accelerated path.

Same signature and the same 6-tuple `(tokenizer, model, image_processor, seg_image_processor, depth_image_processor,
context_len)`; the model object is `vcoder_b200.model.VCoder{,DS}LlavaLlamaForCausalLM` (weights copied into
libvcoder_b200.so) instead of an HF nn.Module.  Class dispatch by substring of `model_name` as in the reference
(`builder.py:93-108`).  `model_base` + `mm_projector.bin` (projector-only checkpoints, builder.py:82-92) loads the base LLM and the projector as
plain LLaVA; `load_4bit` applies bitsandbytes' NF4 + double quantisation to the Linear weights at bind time (numerics of the
reference's `Linear4bit` layers, fp16 storage; parity unpinned — DESIGN.md); LoRA checkpoints (`model_base` + adapter, builder.py:44-81)
are merged here without peft (`merge_lora`); `load_8bit` raises.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict

import torch

from . import synth
from .model import build_model, variant_from_name
from .synth import ModelGeometry

_VT = "model.vision_tower.vision_tower."


def _read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    st = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if st:
        from safetensors.torch import load_file
        for f in st:
            sd.update(load_file(f))
        return sd
    for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))):
        sd.update(torch.load(f, map_location="cpu"))
    if not sd:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {path}")
    return sd


def resolve_tower_path(name: str) -> str:
    """`mm_vision_tower` of the released checkpoints is the hub id `openai/clip-vit-large-patch14-336`, which the
    reference resolves through `from_pretrained` (clip_encoder.py:22-27).  Accepts a local directory, else the local
    HF cache (no network access is attempted)."""
    if os.path.isdir(name):
        return name
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(name, local_files_only=True)
    except Exception as e:
        raise FileNotFoundError(f"vision tower '{name}' is neither a local directory nor in the local HF cache "
                                f"(download it first, e.g. `huggingface-cli download {name}`)") from e


def merge_lora(sd: Dict[str, torch.Tensor], model_path: str) -> Dict[str, torch.Tensor]:
    """builder.py:57-80 without peft: load `non_lora_trainables.bin` over the base weights (same prefix stripping), then
    what `PeftModel.from_pretrained(...).merge_and_unload()` does for LoRA on nn.Linear (peft 0.4.0 `Linear.merge`):
    `W += (lora_B @ lora_A) * (lora_alpha / r)`, evaluated in the weights' own dtype (fp16), adapter keys
    `base_model.model.<module>.lora_{A,B}.weight` (+ `.default` in newer peft files)."""
    nlt_path = os.path.join(model_path, "non_lora_trainables.bin")
    if os.path.exists(nlt_path):
        nlt = torch.load(nlt_path, map_location="cpu")
        nlt = {(k[11:] if k.startswith("base_model.") else k): v for k, v in nlt.items()}
        if any(k.startswith("model.model.") for k in nlt):
            nlt = {(k[6:] if k.startswith("model.") else k): v for k, v in nlt.items()}
        sd.update({k: v.to(torch.float16) for k, v in nlt.items()})
    acfg = json.load(open(os.path.join(model_path, "adapter_config.json")))
    if acfg.get("peft_type", "LORA") != "LORA" or acfg.get("fan_in_fan_out", False):
        raise NotImplementedError("only plain LoRA adapters on nn.Linear are merged")
    scaling = float(acfg["lora_alpha"]) / float(acfg["r"])
    apath = os.path.join(model_path, "adapter_model.safetensors")
    if os.path.exists(apath):
        from safetensors.torch import load_file
        adapter = load_file(apath)
    else:
        adapter = torch.load(os.path.join(model_path, "adapter_model.bin"), map_location="cpu")
    merged = 0
    for k, a in adapter.items():
        if ".lora_A." not in k:
            continue
        kb = k.replace(".lora_A.", ".lora_B.")
        target = k.split(".lora_A.")[0]
        target = target[len("base_model.model."):] if target.startswith("base_model.model.") else target
        name = target + ".weight"
        if name not in sd:
            raise KeyError(f"LoRA adapter targets {name}, which the base checkpoint does not have")
        w = sd[name].to(torch.float16)
        delta = (adapter[kb].to(torch.float16) @ a.to(torch.float16)) * scaling  # peft: in the weight's dtype
        sd[name] = w + delta.to(torch.float16)
        merged += 1
    if merged == 0:
        raise ValueError(f"no lora_A / lora_B pairs found in the adapter under {model_path}")
    return sd


def geometry_from_configs(cfg: dict, vcfg: dict, variant: int) -> ModelGeometry:
    """HF config.json keys the reference reads (vcoder_ds_llava_arch.py:34-49, CLIPVisionConfig, LlamaConfig)."""
    return ModelGeometry(
        name=cfg.get("_name_or_path", "checkpoint"), variant=variant,
        vit_image_size=vcfg.get("image_size", 336), vit_patch=vcfg.get("patch_size", 14),
        vit_hidden=vcfg.get("hidden_size", 1024), vit_heads=vcfg.get("num_attention_heads", 16),
        vit_ffn=vcfg.get("intermediate_size", 4096), vit_layers=vcfg.get("num_hidden_layers", 24),
        vit_select_layer=cfg.get("mm_vision_select_layer", -2), vit_ln_eps=vcfg.get("layer_norm_eps", 1e-5),
        llm_hidden=cfg["hidden_size"], llm_layers=cfg["num_hidden_layers"], llm_heads=cfg["num_attention_heads"],
        llm_ffn=cfg["intermediate_size"], llm_vocab=cfg["vocab_size"], llm_rms_eps=cfg.get("rms_norm_eps", 1e-5),
        llm_rope_theta=cfg.get("rope_theta", 10000.0))


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda", tokenizer=None, max_batch: int = 8):
    if load_8bit:
        raise NotImplementedError("bitsandbytes LLM.int8 loading (builder.py:28-29) is not built: it quantises the activations "
                                  "of every Linear per call (vector-wise int8 + fp16 outlier columns); use load_4bit or fp16")
    if "llava" not in model_name.lower():
        raise ValueError(f"{model_name}: not a LLaVA-family model (plain language models are not on the accelerated path)")
    lora = model_base is not None and "lora" in model_name.lower()  # builder.py:44-81
    projector_only = model_base is not None and not lora  # builder.py:82-92: base LLM + mm_projector.bin, always LlavaLlamaForCausalLM
    if lora:  # builder.py:48-53: VCoderITLlavaLlamaForCausalLM for 'vcoder_it' names, else LlavaLlamaForCausalLM
        variant = synth.VARIANT_VCODER_IT if "vcoder_it" in model_name.lower() else synth.VARIANT_LLAVA
    else:
        variant = synth.VARIANT_LLAVA if projector_only else variant_from_name(model_name)
    cfg = json.load(open(os.path.join(model_path, "config.json")))
    tower_path = resolve_tower_path(cfg["mm_vision_tower"])
    vcfg = json.load(open(os.path.join(tower_path, "config.json")))
    vcfg = vcfg.get("vision_config", vcfg)
    g = geometry_from_configs(cfg, vcfg, variant)

    if tokenizer is None:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(model_base if (projector_only or lora) else model_path, use_fast=False)  # builder.py:47/:85/:96

    if lora:
        sd = merge_lora(_read_state_dict(model_base), model_path)
    elif projector_only:
        sd = _read_state_dict(model_base)
        extra = torch.load(os.path.join(model_path, "mm_projector.bin"), map_location="cpu")  # builder.py:89-91
        sd.update({k: v.to(torch.float16) for k, v in extra.items()})
    else:
        sd = _read_state_dict(model_path)
    # the CLIP tower is delay-loaded from its own checkpoint by the reference (builder.py:138-143)
    if not any(k.startswith(_VT) for k in sd):
        tower_sd = _read_state_dict(tower_path)
        for k, v in tower_sd.items():
            if k.startswith("vision_model."):
                sd[_VT + k] = v
    # clip_encoder.py:24: a failure to load the processor surfaces here, as in the reference (callers dereference it)
    from transformers import CLIPImageProcessor
    image_processor = CLIPImageProcessor.from_pretrained(tower_path)
    # builder.py:138-151: the same processor object, handed out by SUBSTRING of the model name (not by class)
    seg_image_processor = image_processor if "vcoder" in model_name.lower() else None
    depth_image_processor = image_processor if "ds" in model_name.lower() else None

    context_len = cfg.get("max_sequence_length", 2048)  # builder.py:133-136
    # KV-cache capacity: the reference has no hard limit short of the rotary table (max_position_embeddings, 4096 for
    # vicuna-1.5); a VCoder-DS prompt is ~1800 spliced tokens and the callers ask for max_new_tokens=512, so the
    # `context_len` the reference merely reports (2048) must not cap the engine
    max_seq_len = max(int(cfg.get("max_position_embeddings", 2048)), context_len, 2048)
    dev_index = torch.device(device).index or 0
    from .model import LlavaLlamaForCausalLM
    kw = dict(max_batch=max_batch, max_seq_len=max_seq_len, device=dev_index, image_processor=image_processor)
    if load_4bit:
        # builder.py:30-37: NF4 + double quantisation with fp16 compute.  The Linear weights are replaced by their
        # de-quantised values at bind time (same arithmetic as bnb's de-quantise + matmul; not its 4-bit storage)
        kw["load_4bit"] = True
    if lora:
        from .model import VCoderITLlavaLlamaForCausalLM
        model = (VCoderITLlavaLlamaForCausalLM if variant == synth.VARIANT_VCODER_IT else LlavaLlamaForCausalLM)(g, sd, **kw)
    else:
        model = LlavaLlamaForCausalLM(g, sd, **kw) if projector_only else build_model(model_name, g, sd, **kw)
    # the callers hand `model.config` to process_images (eval/model_depth_loader.py:76, serve/chat.py:59): every
    # config.json key (image_aspect_ratio = 'pad', mm_use_im_start_end, ...) must be visible on it
    for k, v in cfg.items():
        setattr(model.config, k, v)
    # generation defaults HF would read: generation_config.json, else the model config's eos/pad ids
    gen = {}
    gpath = os.path.join(model_path, "generation_config.json")
    if os.path.exists(gpath):
        gen = json.load(open(gpath))
    model.generation_config.eos_token_id = gen.get("eos_token_id", cfg.get("eos_token_id"))
    model.generation_config.pad_token_id = gen.get("pad_token_id", cfg.get("pad_token_id"))
    model.generation_config.top_k = gen.get("top_k", 50)
    return tokenizer, model, image_processor, seg_image_processor, depth_image_processor, context_len
