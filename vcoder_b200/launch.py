"""Run the reference's own entry points, unmodified, on the B200 path.

    python -m vcoder_b200.launch [--reference /path/to/VCoder] vcoder_llava.serve.cli --model-path ... --image-file ...
    python -m vcoder_b200.launch [--reference /path/to/VCoder] vcoder_llava.eval.model_depth_loader --model-path ...
    python -m vcoder_b200.launch --gpu-preprocess vcoder_llava.serve.cli ...     (also swaps mm_utils.process_images)

The reference has no plugin interface: its callers reach the hot path through
`load_pretrained_model(...)` (vcoder_llava/model/builder.py:25-154) and `model.generate(input_ids, images=, segs=,
depths=, ...)` (serve/cli.py:121-132, eval/model_depth_loader.py:146-168, eval/model_seg_loader.py:127-149,
serve/chat.py:141-151).  `install()` swaps the ONE name those callers import — `vcoder_llava.model.builder.
load_pretrained_model` — for `vcoder_b200.builder.load_pretrained_model`, whose model object mirrors
VCoder{,DS,IT}LlavaLlamaForCausalLM / LlavaLlamaForCausalLM on top of libvcoder_b200.so.  Nothing under the reference
tree is edited; the swap happens in this process, before the caller module is imported and run with `runpy`.

Two import-time incompatibilities of the reference with current `transformers` are bridged here as well (SURVEY
Appendix B), because the callers cannot even be imported without them:
  * `AutoConfig.register("llava", ...)` (language_model/llava_llama.py:139) collides with the "llava" model type that
    transformers >= 4.36 ships: registration is retried with `exist_ok=True`;
  * `shortuuid` (imported by eval/model_*_loader.py, never used on the path) is stubbed when absent.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types
from typing import List, Optional

_installed = False


def _allow_reregistration() -> None:
    """Let the reference's Auto* registrations overwrite same-named entries of newer transformers releases."""
    from transformers import AutoConfig, AutoModelForCausalLM

    if getattr(AutoConfig.register, "_vcb_wrapped", False):
        return
    cfg_register = AutoConfig.register
    model_register = AutoModelForCausalLM.register

    def register_config(model_type, config, exist_ok=False):
        return cfg_register(model_type, config, exist_ok=True)

    def register_model(cls, config_class, model_class, exist_ok=False):
        return model_register(config_class, model_class, exist_ok=True)

    register_config._vcb_wrapped = True
    AutoConfig.register = staticmethod(register_config)
    AutoModelForCausalLM.register = classmethod(register_model)


def _stub_missing(name: str) -> None:
    try:
        importlib.import_module(name)
    except ImportError:
        sys.modules[name] = types.ModuleType(name)


def install(reference_path: Optional[str] = None, gpu_preprocess: bool = False):
    """Make `vcoder_llava`'s callers load the B200 model.  Returns the patched `vcoder_llava.model.builder` module.

    reference_path: checkout of SHI-Labs/VCoder to put on sys.path (default: $VCODER_REFERENCE, else `vcoder_llava`
    must already be importable).
    gpu_preprocess: also swap `vcoder_llava.mm_utils.process_images` (mm_utils.py:28-40) for the bit-exact GPU version
    (`vcoder_b200.preprocess.process_images`).  Off by default: the eval loaders call process_images inside
    forked DataLoader workers (eval/model_depth_loader.py:75-84), where CUDA cannot be initialised; the single-process
    callers (serve/cli.py, serve/chat.py) can take it."""
    global _installed
    reference_path = reference_path or os.environ.get("VCODER_REFERENCE")
    if reference_path:
        reference_path = os.path.abspath(reference_path)
        if not os.path.isdir(os.path.join(reference_path, "vcoder_llava")):
            raise FileNotFoundError(f"{reference_path} does not contain the vcoder_llava package")
        if reference_path not in sys.path:
            sys.path.insert(0, reference_path)
    _allow_reregistration()
    _stub_missing("shortuuid")
    try:
        ref_builder = importlib.import_module("vcoder_llava.model.builder")
    except ImportError as e:
        raise ImportError("vcoder_llava is not importable: pass --reference /path/to/VCoder or set VCODER_REFERENCE") from e
    from . import builder as b200_builder

    if not _installed:
        ref_builder._reference_load_pretrained_model = ref_builder.load_pretrained_model
    ref_builder.load_pretrained_model = b200_builder.load_pretrained_model
    # modules that already did `from vcoder_llava.model.builder import load_pretrained_model` hold the old name
    for mod in list(sys.modules.values()):
        if mod is None or mod is ref_builder or not getattr(mod, "__name__", "").startswith("vcoder_llava"):
            continue
        if getattr(mod, "load_pretrained_model", None) is getattr(ref_builder, "_reference_load_pretrained_model", None):
            mod.load_pretrained_model = b200_builder.load_pretrained_model
    _installed = True
    if gpu_preprocess:
        from . import preprocess as b200_pre
        ref_mm = importlib.import_module("vcoder_llava.mm_utils")
        if not hasattr(ref_mm, "_reference_process_images"):
            ref_mm._reference_process_images = ref_mm.process_images
        ref_mm.process_images = b200_pre.process_images
        for mod in list(sys.modules.values()):
            if mod is not None and mod is not ref_mm and getattr(mod, "__name__", "").startswith("vcoder_llava") and \
                    getattr(mod, "process_images", None) is ref_mm._reference_process_images:
                mod.process_images = b200_pre.process_images
    return ref_builder


def uninstall() -> None:
    """Restore the reference's own load_pretrained_model (tests)."""
    global _installed
    ref_builder = sys.modules.get("vcoder_llava.model.builder")
    if _installed and ref_builder is not None and hasattr(ref_builder, "_reference_load_pretrained_model"):
        from . import builder as b200_builder
        orig = ref_builder._reference_load_pretrained_model
        ref_builder.load_pretrained_model = orig
        for mod in list(sys.modules.values()):
            if mod is not None and getattr(mod, "__name__", "").startswith("vcoder_llava") and \
                    getattr(mod, "load_pretrained_model", None) is b200_builder.load_pretrained_model:
                mod.load_pretrained_model = orig
    ref_mm = sys.modules.get("vcoder_llava.mm_utils")
    if ref_mm is not None and hasattr(ref_mm, "_reference_process_images"):
        from . import preprocess as b200_pre
        orig_pi = ref_mm._reference_process_images
        for mod in list(sys.modules.values()):
            if mod is not None and getattr(mod, "__name__", "").startswith("vcoder_llava") and \
                    getattr(mod, "process_images", None) is b200_pre.process_images:
                mod.process_images = orig_pi
    _installed = False


def main(argv: Optional[List[str]] = None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    reference, gpu_preprocess = None, False
    while argv and argv[0].startswith("--"):
        if argv[0] == "--reference":
            if len(argv) < 2:
                raise SystemExit("--reference needs a path")
            reference, argv = argv[1], argv[2:]
        elif argv[0].startswith("--reference="):
            reference, argv = argv[0].split("=", 1)[1], argv[1:]
        elif argv[0] == "--gpu-preprocess":
            gpu_preprocess, argv = True, argv[1:]
        else:
            break
    if not argv or argv[0].startswith("-"):
        raise SystemExit(__doc__)
    target = argv[0]
    install(reference, gpu_preprocess=gpu_preprocess)
    sys.argv = [target] + argv[1:]
    runpy.run_module(target, run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()
