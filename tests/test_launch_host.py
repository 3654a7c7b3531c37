"""CPU test of `vcoder_b200.launch`: the reference's OWN callers — `vcoder_llava.serve.cli` (serve/cli.py:30-139) and
`vcoder_llava.eval.model_depth_loader` (eval/model_depth_loader.py:116-185) — run unmodified through the launcher,
`vcoder_b200.builder.load_pretrained_model` and the real `vcoder_b200.model` shim.  Only the C-ABI engine underneath
the shim is replaced by a recording fake (there is no GPU here and the product has no CPU fallback), so what is
checked is the host side of the drop-in boundary: the name swap, the 6-tuple, the kwargs `generate` receives, the
streamer / stopping-criteria / prefix-cache flow, and what the callers do with the returned ids.

Needs the reference checkout (this container: /root/reference); skipped where it is absent (the GPU box).
"""
import builtins
import json
import os
import sys
import types

import pytest
import torch

from vcoder_b200 import synth

REFERENCE = os.environ.get("VCODER_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "vcoder_llava")),
                                reason="reference checkout not present")

I, S, D = synth.IMAGE_TOKEN_INDEX, synth.SEG_TOKEN_INDEX, synth.DEPTH_TOKEN_INDEX


class FakeTokenizer:
    """Whitespace tokenizer with a growing vocabulary; '</s>' is id 2, bos is id 1."""
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def __init__(self):
        self.vocab = {"<pad>": 0, "<s>": 1, "</s>": 2}
        self.words = ["<pad>", "<s>", "</s>"]

    def _id(self, w):
        if w not in self.vocab:
            self.vocab[w] = len(self.words)
            self.words.append(w)
        return self.vocab[w]

    def __call__(self, text):
        return types.SimpleNamespace(input_ids=[1] + [self._id(w) for w in text.replace("</s>", " </s> ").split()])

    def decode(self, ids, skip_special_tokens=False, **kw):
        ids = [int(i) for i in (ids.tolist() if hasattr(ids, "tolist") else ids)]
        out = []
        for i in ids:
            w = self.words[i] if 0 <= i < len(self.words) else f"tok{i}"
            if skip_special_tokens and i < 3:
                continue
            out.append(w)
        return " ".join(out)

    def batch_decode(self, ids, skip_special_tokens=False, **kw):
        return [self.decode(r, skip_special_tokens=skip_special_tokens) for r in ids]


class FakeEngine:
    """Stands in for vcoder_b200.engine.Engine: produces the scripted tokens, records every call."""
    instances = []
    script = [11, 12, 13, 2, 14, 15, 16, 17]

    def __init__(self, g, max_batch=8, max_seq_len=1536, device=0):
        self.g, self.max_batch, self.max_seq_len = g, max_batch, max_seq_len
        self.tdev = torch.device("cpu")
        self.calls = []
        self.bound = {}
        self.pos = 0
        FakeEngine.instances.append(self)

    def bind(self, sd):
        self.bound.update({k: tuple(v.shape) for k, v in sd.items()})

    def set_sampling(self, do_sample, temperature=1.0, top_k=50, top_p=1.0, seed=0):
        self.calls.append(("set_sampling", bool(do_sample), temperature, top_k, top_p))

    def _check_px(self, ids, *px):
        for t in px:
            assert t is None or (t.dtype == torch.float16 and tuple(t.shape) == (ids.shape[0], 3, self.g.vit_image_size, self.g.vit_image_size))

    def prefill(self, ids, images=None, segs=None, depths=None, want_logits=False, want_embeds=False, reuse=False):
        self._check_px(ids, images, segs, depths)
        self.calls.append(("prefill", ids.clone(), images is not None, segs is not None, depths is not None, reuse))
        self.pos = 1
        B = ids.shape[0]
        return {"next_ids": torch.full((B,), self.script[0], dtype=torch.int32), "spliced_len": ids.shape[1] + 5,
                "reused_len": 0, "towers_skipped": False}

    def decode_stream(self, max_steps, on_token, lookahead=4):
        n = 0
        for s in range(max_steps):
            tok = self.script[self.pos % len(self.script)]
            self.pos += 1
            n += 1
            if on_token(s, [tok]):
                break
        self.calls.append(("decode_stream", max_steps, n, lookahead))
        return n

    def generate(self, ids, images=None, segs=None, depths=None, max_new_tokens=32):
        self._check_px(ids, images, segs, depths)
        self.calls.append(("generate", ids.clone(), images is not None, segs is not None, depths is not None, max_new_tokens))
        row = [self.script[i % len(self.script)] for i in range(max_new_tokens)]
        return torch.tensor([row] * ids.shape[0], dtype=torch.int64)


def _write_checkpoint(root, g, name="vcoder_ds_llava-tiny"):
    from safetensors.torch import save_file
    from transformers import CLIPImageProcessor
    sd = synth.make_state_dict(g, seed=0)
    model_dir = os.path.join(root, name)
    tower = os.path.join(root, "tower_" + name)
    os.makedirs(model_dir)
    os.makedirs(tower)
    vt = "model.vision_tower.vision_tower."
    save_file({k[len(vt):]: v.contiguous() for k, v in sd.items() if k.startswith(vt)}, os.path.join(tower, "model.safetensors"))
    json.dump({"hidden_size": g.vit_hidden, "intermediate_size": g.vit_ffn, "num_hidden_layers": g.vit_layers,
               "num_attention_heads": g.vit_heads, "image_size": g.vit_image_size, "patch_size": g.vit_patch,
               "layer_norm_eps": g.vit_ln_eps}, open(os.path.join(tower, "config.json"), "w"))
    CLIPImageProcessor(size={"shortest_edge": g.vit_image_size}, crop_size=g.vit_image_size).save_pretrained(tower)
    save_file({k: v.contiguous() for k, v in sd.items() if not k.startswith(vt)}, os.path.join(model_dir, "model.safetensors"))
    json.dump({"hidden_size": g.llm_hidden, "intermediate_size": g.llm_ffn, "num_hidden_layers": g.llm_layers,
               "num_attention_heads": g.llm_heads, "vocab_size": g.llm_vocab, "rms_norm_eps": g.llm_rms_eps,
               "mm_vision_tower": tower, "mm_vision_select_layer": -2, "mm_projector_type": "mlp2x_gelu",
               "max_sequence_length": 256, "eos_token_id": 2, "pad_token_id": 0}, open(os.path.join(model_dir, "config.json"), "w"))
    return model_dir


def _jpg(path, w, h, colour):
    from PIL import Image
    os.makedirs(os.path.dirname(path), exist_ok=True)
    Image.new("RGB", (w, h), colour).save(path)


@pytest.fixture()
def rig(tmp_path, monkeypatch):
    from vcoder_b200 import launch
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    model_dir = _write_checkpoint(str(tmp_path), g)
    tok = FakeTokenizer()
    FakeEngine.instances = []
    monkeypatch.setattr("vcoder_b200.model.Engine", FakeEngine)
    import transformers
    monkeypatch.setattr(transformers.AutoTokenizer, "from_pretrained", classmethod(lambda cls, *a, **k: tok))
    # the callers hard-code .cuda() / .to(device='cuda') (serve/cli.py:112, eval/model_depth_loader.py:144-153)
    real_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return real_to(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "to", to_cpu)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    argv = list(sys.argv)
    yield types.SimpleNamespace(g=g, model_dir=model_dir, tok=tok, tmp=str(tmp_path), launch=launch)
    sys.argv = argv
    launch.uninstall()


def test_install_swaps_the_one_name_the_callers_import(rig):
    ref_builder = rig.launch.install(REFERENCE)
    from vcoder_b200 import builder as b200_builder
    assert ref_builder.load_pretrained_model is b200_builder.load_pretrained_model
    import vcoder_llava.eval.model_seg_loader as seg_loader
    import vcoder_llava.serve.cli as cli
    assert cli.load_pretrained_model is b200_builder.load_pretrained_model
    assert seg_loader.load_pretrained_model is b200_builder.load_pretrained_model
    rig.launch.uninstall()
    assert ref_builder.load_pretrained_model.__module__ == "vcoder_llava.model.builder"
    assert cli.load_pretrained_model.__module__ == "vcoder_llava.model.builder"
    with pytest.raises(FileNotFoundError):
        rig.launch.install(rig.tmp)  # not a VCoder checkout
    # f2 opt-in: mm_utils.process_images -> the GPU version, in the callers that imported it by name as well
    from vcoder_b200 import preprocess as b200_pre
    rig.launch.install(REFERENCE, gpu_preprocess=True)
    import vcoder_llava.mm_utils as mm
    assert mm.process_images is b200_pre.process_images and cli.process_images is b200_pre.process_images
    rig.launch.uninstall()
    assert cli.process_images.__module__ == "vcoder_llava.mm_utils" and mm.process_images.__module__ == "vcoder_llava.mm_utils"


def test_reference_cli_runs_unmodified(rig, monkeypatch, capsys):
    img = os.path.join(rig.tmp, "in", "rgb.jpg")
    seg = os.path.join(rig.tmp, "in", "seg.jpg")
    dep = os.path.join(rig.tmp, "in", "depth.jpg")
    _jpg(img, 40, 30, (200, 10, 10)); _jpg(seg, 40, 30, (10, 200, 10)); _jpg(dep, 40, 30, (90, 90, 90))
    turns = iter(["what is in the picture ?", "and behind it ?", ""])
    monkeypatch.setattr(builtins, "input", lambda prompt="": next(turns))
    rig.launch.main(["--reference", REFERENCE, "vcoder_llava.serve.cli", "--model-path", rig.model_dir,
                     "--image-file", img, "--seg-file", seg, "--depth-file", dep, "--temperature", "0.2",
                     "--max-new-tokens", "16"])
    out = capsys.readouterr().out
    assert "exit..." in out
    eng = FakeEngine.instances[-1]
    assert eng.max_batch == 8 and any(k.startswith("model.layers.0.") for k in eng.bound)
    pre = [c for c in eng.calls if c[0] == "prefill"]
    assert len(pre) == 2  # two turns, both through the interactive (streamer + stopping criteria) path
    for c in pre:
        assert c[2] and c[3] and c[4] and c[5]  # images, segs, depths passed; prefix cache on
    ids1, ids2 = pre[0][1], pre[1][1]
    # _tokenizer_depth_token's special order (mm_utils.py:101-103)
    specials = [int(t) for t in ids1[0] if t < 0]
    assert specials == [I, D, S]
    # turn 2 re-sends the whole conversation: turn 1's prompt is a prefix of it, and the images are re-sent as well
    assert ids2.shape[1] > ids1.shape[1] and torch.equal(ids2[0, : ids1.shape[1] - 1], ids1[0, : ids1.shape[1] - 1])
    # do_sample=True, temperature=0.2, top_k = HF default 50, top_p off
    assert ("set_sampling", True, 0.2, 50, 1.0) in eng.calls
    # KeywordsStoppingCriteria(["</s>"]) stops each turn at the scripted eos: tokens 11 12 13 </s>
    ds = [c for c in eng.calls if c[0] == "decode_stream"]
    assert [c[2] for c in ds] == [3, 3]
    assert rig.tok.decode([11, 12, 13]) in out.replace("\n", " ")  # TextStreamer printed the scripted answer


def test_reference_eval_model_runs_unmodified(rig):
    for i in range(2):
        _jpg(os.path.join(rig.tmp, "rgb", f"{i}.jpg"), 36, 36, (30 * i, 100, 100))
        _jpg(os.path.join(rig.tmp, "seg", "panoptic_inference", f"{i}.jpg"), 36, 36, (0, 30 * i, 100))
        _jpg(os.path.join(rig.tmp, "depth", f"{i}.jpg"), 36, 36, (50, 50, 50 + i))
    out_prefix = os.path.join(rig.tmp, "answers", "depth")
    rig.launch.main(["--reference", REFERENCE, "vcoder_llava.eval.model_depth_loader", "--model-path", rig.model_dir,
                     "--image-folder", os.path.join(rig.tmp, "rgb"), "--use_depth_seg",
                     "--seg-image-folder", os.path.join(rig.tmp, "seg"), "--depth_image-folder", os.path.join(rig.tmp, "depth"),
                     "--output-file", out_prefix, "--temperature", "0"])
    text = open(out_prefix + "_1_0.txt").read()
    assert text.count("<<ANSWER>>") == 2
    # greedy (temperature 0) without streamer / stopping criteria, but the checkpoint's eos id makes the call interactive
    eng = FakeEngine.instances[-1]
    assert ("set_sampling", False, 1.0, 50, 1.0) in eng.calls
    pre = [c for c in eng.calls if c[0] == "prefill"]
    assert len(pre) == 2 and all(c[2] and c[3] and c[4] for c in pre)
    assert [int(t) for t in pre[0][1][0] if t < 0] == [I, D, S]
    # batch_decode(skip_special_tokens=True) of [11, 12, 13, </s>]
    assert f"<<ANSWER>>: {rig.tok.decode([11, 12, 13])}\n" in text


def test_reference_chat_generate_stream_runs_unmodified(rig):
    """serve/chat.py:41-159 (the Gradio worker's `Chat.generate_stream`): base64 images in, `model.generate(inputs=...,
    do_sample, temperature, top_p, streamer=TextIteratorStreamer, stopping_criteria)` — two turns with the same images."""
    import base64
    import io
    import logging
    from PIL import Image
    rig.launch.install(REFERENCE)
    import vcoder_llava.serve.chat as chat
    assert chat.load_pretrained_model.__module__ == "vcoder_b200.builder"

    def b64(colour):
        buf = io.BytesIO()
        Image.new("RGB", (48, 40), colour).save(buf, format="PNG")
        return base64.b64encode(buf.getvalue()).decode()

    c = chat.Chat(rig.model_dir, None, None, False, False, "cuda", logging.getLogger("test"))
    assert c.is_multimodal and c.is_seg and c.is_depth
    imgs, segs, deps = [b64((200, 0, 0))], [b64((0, 200, 0))], [b64((50, 50, 50))]
    p1 = "USER: <depth>\n<seg>\n<image>\nwhat is this ? ASSISTANT:"
    params = {"prompt": p1, "images": imgs, "segs": segs, "depths": deps, "temperature": 0.2, "top_p": 0.7,
              "max_new_tokens": 16, "stop": "</s>"}
    outs = [json.loads(chunk.decode().rstrip("\0")) for chunk in c.generate_stream(params)]
    assert outs and all(o["error_code"] == 0 for o in outs)
    answer = rig.tok.decode([11, 12, 13])
    assert outs[-1]["text"].startswith(p1) and answer in outs[-1]["text"]
    eng = FakeEngine.instances[-1]
    assert ("set_sampling", True, 0.2, 50, 0.7) in eng.calls  # do_sample since temperature > 0.001; top_p forwarded
    pre = [x for x in eng.calls if x[0] == "prefill"]
    assert len(pre) == 1 and pre[0][2] and pre[0][3] and pre[0][4] and pre[0][5]
    assert [int(t) for t in pre[0][1][0] if t < 0] == [I, D, S]
    # second turn: whole conversation + the same base64 images again (what the prefix cache is for)
    params["prompt"] = outs[-1]["text"] + "</s>USER: and now ? ASSISTANT:"
    outs2 = [json.loads(chunk.decode().rstrip("\0")) for chunk in c.generate_stream(params)]
    assert outs2[-1]["error_code"] == 0
    pre = [x for x in eng.calls if x[0] == "prefill"]
    assert len(pre) == 2 and pre[1][1].shape[1] > pre[0][1].shape[1] and pre[1][5]
    # greedy when temperature <= 0.001
    params["temperature"] = 0.0
    list(c.generate_stream(params))
    assert eng.calls[-3][0] == "set_sampling" and eng.calls[-3][1] is False


def test_reference_seg_eval_model_runs_unmodified_on_vcoder(rig):
    """eval/model_seg_loader.py:99-166 with a VCoder (RGB + seg, non-DS) checkpoint: three tasks, each loading the model and
    calling generate(images=, segs=, depths=None, do_sample=False, top_p=None, num_beams=1, max_new_tokens=512)."""
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER)
    model_dir = _write_checkpoint(rig.tmp, g, name="vcoder_llava-tiny")
    _jpg(os.path.join(rig.tmp, "rgb2", "0.jpg"), 40, 36, (10, 100, 200))
    for task in ("semantic", "instance", "panoptic"):
        _jpg(os.path.join(rig.tmp, "seg2", f"{task}_inference", "0.jpg"), 40, 36, (0, 90, 100))
    out_prefix = os.path.join(rig.tmp, "answers_seg", "seg")
    n0 = len(FakeEngine.instances)
    rig.launch.main(["--reference", REFERENCE, "vcoder_llava.eval.model_seg_loader", "--model-path", model_dir,
                     "--image-folder", os.path.join(rig.tmp, "rgb2"), "--use_seg", "--seg-image-folder", os.path.join(rig.tmp, "seg2"),
                     "--output-file", out_prefix, "--temperature", "0"])
    engines = FakeEngine.instances[n0:]
    assert len(engines) == 3  # eval_model builds the model once per task
    for task, eng in zip(("semantic", "instance", "panoptic"), engines):
        assert eng.g.variant == synth.VARIANT_VCODER
        text = open(f"{out_prefix}_{task}_1_0.txt").read()
        assert f"<<ANSWER>>: {rig.tok.decode([11, 12, 13])}\n" in text
        pre = [c for c in eng.calls if c[0] == "prefill"]
        assert len(pre) == 1 and pre[0][2] and pre[0][3] and not pre[0][4]  # images + segs, no depths
        assert [int(t) for t in pre[0][1][0] if t < 0] == [I, S]  # tokenizer_seg_token emits [IMAGE, SEG] (mm_utils.py:65-87)
        assert ("set_sampling", False, 1.0, 50, 1.0) in eng.calls


def test_generate_batches_larger_than_the_handle_run_as_slices(monkeypatch):
    """The reference's generate() takes any batch size; the handle holds `max_batch` sequences, so a larger batch runs
    as consecutive slices (each with its own pixel rows) and the rows come back in order."""
    from vcoder_b200.model import VCoderDSLlavaLlamaForCausalLM
    FakeEngine.instances = []
    monkeypatch.setattr("vcoder_b200.model.Engine", FakeEngine)
    g = synth.PRESETS["tiny"].with_variant(synth.VARIANT_VCODER_DS)
    model = VCoderDSLlavaLlamaForCausalLM(g, None, max_batch=4)
    model.generation_config.eos_token_id = None
    B, px_size = 10, g.vit_image_size
    ids = torch.stack([torch.tensor([1, 5 + b, I, S, D, 9]) for b in range(B)])
    px = [torch.randn(B, 3, px_size, px_size).half() for _ in range(3)]
    out = model.generate(ids, images=px[0], segs=px[1], depths=px[2], do_sample=False, max_new_tokens=3)
    eng = FakeEngine.instances[-1]
    gens = [c for c in eng.calls if c[0] == "generate"]
    assert [c[1].shape[0] for c in gens] == [4, 4, 2]
    assert torch.equal(torch.cat([c[1] for c in gens]), ids)  # every row once, in order
    assert out.shape == (B, ids.shape[1] + 3) and torch.equal(out[:, :ids.shape[1]], ids)
    assert out[:, ids.shape[1]:].tolist() == [[11, 12, 13]] * B
