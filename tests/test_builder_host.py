"""CPU tests of `vcoder_b200.builder.load_pretrained_model` (mirror of vcoder_llava/model/builder.py:25-154): which
checkpoint layouts it accepts, the class it dispatches to, the processors it hands out and what it binds — with the C-ABI
engine replaced by a recorder (no GPU here; the GPU version of the full-checkpoint path is tests/test_gpu_shim.py)."""
import json
import os

import pytest
import torch

from vcoder_b200 import synth


class RecordingEngine:
    last = None

    def __init__(self, g, max_batch=8, max_seq_len=1536, device=0):
        self.g, self.bound, self.tdev = g, {}, torch.device("cpu")
        RecordingEngine.last = self

    def bind(self, sd):
        self.bound.update({k: v.dtype for k, v in sd.items()})


def _save(path, sd):
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    save_file({k: v.contiguous() for k, v in sd.items()}, os.path.join(path, "model.safetensors"))


def _configs(g, model_dir, tower):
    from transformers import CLIPImageProcessor
    os.makedirs(model_dir, exist_ok=True)
    json.dump({"hidden_size": g.vit_hidden, "intermediate_size": g.vit_ffn, "num_hidden_layers": g.vit_layers,
               "num_attention_heads": g.vit_heads, "image_size": g.vit_image_size, "patch_size": g.vit_patch},
              open(os.path.join(tower, "config.json"), "w"))
    CLIPImageProcessor(size={"shortest_edge": g.vit_image_size}, crop_size=g.vit_image_size).save_pretrained(tower)
    json.dump({"hidden_size": g.llm_hidden, "intermediate_size": g.llm_ffn, "num_hidden_layers": g.llm_layers,
               "num_attention_heads": g.llm_heads, "vocab_size": g.llm_vocab, "rms_norm_eps": g.llm_rms_eps,
               "mm_vision_tower": tower, "mm_vision_select_layer": -2, "eos_token_id": 2},
              open(os.path.join(model_dir, "config.json"), "w"))


def test_projector_only_checkpoint_with_model_base(tmp_path, monkeypatch):
    """builder.py:82-92: `model_base` (the LLM) + `mm_projector.bin` under model_path -> always LlavaLlamaForCausalLM."""
    from vcoder_b200 import builder
    monkeypatch.setattr("vcoder_b200.model.Engine", RecordingEngine)
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_LLAVA)
    sd = synth.make_state_dict(g, seed=0)
    vt = "model.vision_tower.vision_tower."
    tower, base, model_dir = (str(tmp_path / n) for n in ("tower", "vicuna-tiny", "llava-tiny-pretrain"))
    _save(tower, {k[len(vt):]: v for k, v in sd.items() if k.startswith(vt)})
    _save(base, {k: v for k, v in sd.items() if not k.startswith(vt) and "mm_projector" not in k})
    _configs(g, model_dir, tower)
    torch.save({k: v.float() for k, v in sd.items() if "mm_projector" in k}, os.path.join(model_dir, "mm_projector.bin"))
    tok, model, ip, sip, dip, ctx = builder.load_pretrained_model(model_dir, base, "llava-tiny-pretrain", tokenizer="stub")
    assert type(model).__name__ == "LlavaLlamaForCausalLM" and model.geometry.variant == synth.VARIANT_LLAVA
    assert ip is not None and sip is None and dip is None and ctx == 2048
    eng = RecordingEngine.last
    assert "model.mm_projector.0.weight" in eng.bound and "model.layers.0.self_attn.q_proj.weight" in eng.bound
    assert vt + "vision_model.embeddings.class_embedding" in eng.bound and "lm_head.weight" in eng.bound
    assert model.generation_config.eos_token_id == 2
    # the processors follow SUBSTRINGS of the model name, not the class (builder.py:138-151)
    _, model, ip, sip, dip, _ = builder.load_pretrained_model(model_dir, base, "vcoder_ds_llava-projector", tokenizer="stub")
    assert type(model).__name__ == "LlavaLlamaForCausalLM" and sip is ip and dip is ip


def test_rejected_layouts(tmp_path, monkeypatch):
    from vcoder_b200 import builder
    monkeypatch.setattr("vcoder_b200.model.Engine", RecordingEngine)
    with pytest.raises(NotImplementedError):
        builder.load_pretrained_model(str(tmp_path), None, "llava-tiny", load_4bit=True)
    with pytest.raises(NotImplementedError):
        builder.load_pretrained_model(str(tmp_path), None, "llava-tiny", load_8bit=True)
    with pytest.raises(NotImplementedError):  # LoRA needs peft (builder.py:44-81)
        builder.load_pretrained_model(str(tmp_path), str(tmp_path), "llava-tiny-lora", tokenizer="stub")
    with pytest.raises(ValueError):  # plain language models (builder.py:109-128) are not on the accelerated path
        builder.load_pretrained_model(str(tmp_path), None, "vicuna-7b", tokenizer="stub")
