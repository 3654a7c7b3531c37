"""Python handle over the C-ABI engine (vcb_create / vcb_bind_weights / vcb_prefill / vcb_decode / vcb_generate).

torch is used for device memory and streams only; every FLOP on the path runs inside libvcoder_b200.so.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .synth import ModelGeometry


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_half_any(t) -> Optional[torch.Tensor]:
    """fp16, contiguous, on whatever device it already lives (host buffers are copied by the library).  A list of
    [3,S,S] tensors (clip_encoder.py:40-45 accepts one) is stacked."""
    if t is None:
        return None
    if isinstance(t, (list, tuple)):
        if not t or any(not torch.is_tensor(x) or x.dim() != 3 or x.shape != t[0].shape for x in t):
            raise ValueError("a list of images must hold [3,S,S] tensors of one shape (mixed sizes are not supported)")
        t = torch.stack(list(t), dim=0)
    if not torch.is_tensor(t):
        raise ValueError(f"images must be a tensor or a list of tensors, not {type(t).__name__}")
    if t.dtype != torch.float16:
        t = t.to(torch.float16)
    return t.contiguous()


class Engine:
    def __init__(self, g: ModelGeometry, max_batch: int = 8, max_seq_len: int = 1536, device: int = 0):
        if not torch.cuda.is_available():
            # fail loudly: there is no CPU fallback anywhere in this package
            raise RuntimeError("vcoder_b200.Engine needs a CUDA device (no CPU fallback exists)")
        self.lib = _lib.load()
        self.g = g
        self.device = device
        self.max_batch = max_batch
        self.max_seq_len = max_seq_len
        self.lock = threading.Lock()  # a handle is single-stream and not re-entrant
        cfg = _lib.VcbConfig(
            struct_size=C.sizeof(_lib.VcbConfig), variant=g.variant,
            vit_image_size=g.vit_image_size, vit_patch=g.vit_patch, vit_hidden=g.vit_hidden, vit_heads=g.vit_heads,
            vit_ffn=g.vit_ffn, vit_layers=g.vit_layers, vit_select_layer=g.vit_select_layer, vit_ln_eps=g.vit_ln_eps,
            llm_hidden=g.llm_hidden, llm_layers=g.llm_layers, llm_heads=g.llm_heads, llm_ffn=g.llm_ffn,
            llm_vocab=g.llm_vocab, llm_rms_eps=g.llm_rms_eps, llm_rope_theta=g.llm_rope_theta,
            max_batch=max_batch, max_seq_len=max_seq_len, kv_page_tokens=64)
        h = C.c_void_p()
        _lib.check(self.lib.vcb_create(C.byref(cfg), device, C.byref(h)))
        self.h = h
        self.tdev = torch.device("cuda", device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vcb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------
    def bind(self, state_dict: Dict[str, torch.Tensor], chunk: int = 64, nf4: bool = False) -> None:
        """Copies (and packs) the named tensors into the handle; the caller may drop its copies afterwards.
        nf4=True (`load_4bit`): the weights bitsandbytes would hold as NF4 `Linear4bit` layers are replaced by their
        de-quantised fp16 values on the way in (vcoder_b200/nf4.py, vcb_op_nf4_roundtrip_f16)."""
        items = [(k, v) for k, v in state_dict.items() if torch.is_tensor(v) and v.dtype.is_floating_point]
        if nf4:
            from . import nf4 as _nf4
        for i in range(0, len(items), chunk):
            part = items[i:i + chunk]
            keep = []
            arr = (_lib.VcbTensor * len(part))()
            for j, (k, v) in enumerate(part):
                t = v.detach()
                if t.dtype != torch.float16:
                    t = t.to(torch.float16)
                t = t.contiguous()
                if nf4 and _nf4.is_bnb_linear(k):
                    t = t.to(self.tdev, copy=True)
                    _nf4.nf4_roundtrip_(t)
                keep.append(t)
                arr[j].name = k.encode()
                arr[j].data = t.data_ptr()
                arr[j].ndim = t.dim()
                for d in range(t.dim()):
                    arr[j].shape[d] = t.shape[d]
            _lib.check(self.lib.vcb_bind_weights(self.h, arr, len(part)))
            del keep

    # ------------------------------------------------------------------------------------------------
    def _check_pixels(self, t: Optional[torch.Tensor], what: str, batch: Optional[int] = None):
        """The library copies n*3*S*S halves from the pointer it is given: anything else must be refused here."""
        if t is None:
            return
        S = self.g.vit_image_size
        if t.dim() != 4 or tuple(t.shape[1:]) != (3, S, S):
            raise ValueError(f"{what}: expected [n, 3, {S}, {S}] (the tower's image size), got {tuple(t.shape)}")
        if batch is not None and t.shape[0] != batch:
            raise ValueError(f"{what}: {t.shape[0]} images for a batch of {batch} sequences")
        if batch is None and t.shape[0] < 1:
            raise ValueError(f"{what}: empty image batch")

    def tower_forward(self, pixels: torch.Tensor) -> torch.Tensor:
        px = _as_half_any(pixels)
        self._check_pixels(px, "tower_forward")
        n = px.shape[0]
        out = torch.empty(n, self.g.n_patches, self.g.vit_hidden, dtype=torch.float16, device=self.tdev)
        with self.lock:
            _lib.check(self.lib.vcb_tower_forward(self.h, _lib.ptr(px), n, _lib.ptr(out), _stream_ptr()))
        return out

    def encode(self, modality: int, pixels: torch.Tensor) -> torch.Tensor:
        px = _as_half_any(pixels)
        self._check_pixels(px, "encode")
        n = px.shape[0]
        out = torch.empty(n, self.g.n_patches, self.g.llm_hidden, dtype=torch.float16, device=self.tdev)
        with self.lock:
            _lib.check(self.lib.vcb_encode(self.h, modality, _lib.ptr(px), n, _lib.ptr(out), _stream_ptr()))
        return out

    def prefill(self, ids: torch.Tensor, images=None, segs=None, depths=None, want_logits=False, want_embeds=False,
                reuse: bool = False):
        """vcb_prefill, or with `reuse=True` vcb_prefill_cached (f3): towers skipped when the pixels equal the previous
        call's, and only the suffix behind the prefix shared with the cached sequence runs through the decoder; the
        result then also carries `reused_len` and `towers_skipped`."""
        if reuse and want_embeds:
            raise ValueError("want_embeds needs the full prompt: not available with reuse=True")
        ids_h = ids.detach().to("cpu", torch.int64).contiguous()
        B, T = ids_h.shape
        self._last_batch = B
        im, sg, dp = _as_half_any(images), _as_half_any(segs), _as_half_any(depths)
        for t, what in ((im, "images"), (sg, "segs"), (dp, "depths")):
            self._check_pixels(t, what, B)
        nxt = torch.empty(B, dtype=torch.int32, device=self.tdev)
        logits = torch.empty(B, self.g.llm_vocab, dtype=torch.float16, device=self.tdev) if want_logits else None
        n_vis = 3 * self.g.n_patches
        embeds = (torch.empty(B, T + n_vis, self.g.llm_hidden, dtype=torch.float16, device=self.tdev)
                  if want_embeds else None)
        slen, reused, skipped = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        with self.lock:
            if reuse:
                _lib.check(self.lib.vcb_prefill_cached(self.h, _lib.ptr(ids_h), B, T, _lib.ptr(im), _lib.ptr(sg),
                                                       _lib.ptr(dp), _lib.ptr(nxt), _lib.ptr(logits), C.byref(slen),
                                                       C.byref(reused), C.byref(skipped), _stream_ptr()))
            else:
                _lib.check(self.lib.vcb_prefill(self.h, _lib.ptr(ids_h), B, T, _lib.ptr(im), _lib.ptr(sg), _lib.ptr(dp),
                                                _lib.ptr(nxt), _lib.ptr(logits), _lib.ptr(embeds), C.byref(slen),
                                                _stream_ptr()))
        out = {"next_ids": nxt, "spliced_len": slen.value, "reused_len": reused.value,
               "towers_skipped": bool(skipped.value)}
        if want_logits:
            out["logits"] = logits
        if want_embeds:
            out["embeds"] = embeds.view(-1)[: B * slen.value * self.g.llm_hidden].view(B, slen.value, self.g.llm_hidden)
        return out

    def forward_logits(self, ids: torch.Tensor, images=None, segs=None, depths=None) -> torch.Tensor:
        """vcb_forward_logits: lm_head(norm(hidden)) of EVERY spliced position, [B, T', V] fp16 on the device (the
        `logits` of the reference's forward(), vcoder_ds_llava_llama.py:95-118).  Fills the KV cache like prefill."""
        ids_h = ids.detach().to("cpu", torch.int64).contiguous()
        B, T = ids_h.shape
        self._last_batch = B
        im, sg, dp = _as_half_any(images), _as_half_any(segs), _as_half_any(depths)
        for t, what in ((im, "images"), (sg, "segs"), (dp, "depths")):
            self._check_pixels(t, what, B)
        t_max = T + (3 * self.g.n_patches if im is not None else 0)
        logits = torch.empty(B * t_max * self.g.llm_vocab, dtype=torch.float16, device=self.tdev)
        slen = C.c_int32(0)
        with self.lock:
            _lib.check(self.lib.vcb_forward_logits(self.h, _lib.ptr(ids_h), B, T, _lib.ptr(im), _lib.ptr(sg), _lib.ptr(dp),
                                                   _lib.ptr(logits), logits.numel(), C.byref(slen), _stream_ptr()))
        return logits[: B * slen.value * self.g.llm_vocab].view(B, slen.value, self.g.llm_vocab)

    def splice_embeds(self, ids: torch.Tensor, images=None, segs=None, depths=None) -> torch.Tensor:
        """vcb_splice_embeds: towers + projectors + splice only -> `new_input_embeds` [B, T', H] fp16 of
        prepare_inputs_labels_for_multimodal (vcoder_ds_llava_arch.py:126-314).  Leaves the handle un-prefilled."""
        ids_h = ids.detach().to("cpu", torch.int64).contiguous()
        B, T = ids_h.shape
        im, sg, dp = _as_half_any(images), _as_half_any(segs), _as_half_any(depths)
        for t, what in ((im, "images"), (sg, "segs"), (dp, "depths")):
            self._check_pixels(t, what, B)
        t_max = T + (3 * self.g.n_patches if im is not None else 0)
        embeds = torch.empty(B * t_max * self.g.llm_hidden, dtype=torch.float16, device=self.tdev)
        slen = C.c_int32(0)
        with self.lock:
            _lib.check(self.lib.vcb_splice_embeds(self.h, _lib.ptr(ids_h), B, T, _lib.ptr(im), _lib.ptr(sg), _lib.ptr(dp),
                                                  _lib.ptr(embeds), C.byref(slen), _stream_ptr()))
        return embeds[: B * slen.value * self.g.llm_hidden].view(B, slen.value, self.g.llm_hidden)

    def decode(self, n_steps: int, want_logits=False):
        B = self._batch()
        ids = torch.empty(B, n_steps, dtype=torch.int32, device=self.tdev)
        trace = (torch.empty(n_steps, B, self.g.llm_vocab, dtype=torch.float16, device=self.tdev)
                 if want_logits else None)
        with self.lock:
            _lib.check(self.lib.vcb_decode(self.h, n_steps, _lib.ptr(ids), _lib.ptr(trace), _stream_ptr()))
        return (ids, trace) if want_logits else ids

    def generate(self, ids: torch.Tensor, images=None, segs=None, depths=None, max_new_tokens: int = 32) -> torch.Tensor:
        """HOST in, HOST out: the e2e entry (H2D of pixels/ids and D2H of the ids happen inside the call)."""
        ids_h = ids.detach().to("cpu", torch.int64).contiguous()
        B, T = ids_h.shape
        self._last_batch = B
        im, sg, dp = _as_half_any(images), _as_half_any(segs), _as_half_any(depths)
        for t, what in ((im, "images"), (sg, "segs"), (dp, "depths")):
            self._check_pixels(t, what, B)
        out = torch.empty(B, max_new_tokens, dtype=torch.int64)
        with self.lock:
            _lib.check(self.lib.vcb_generate(self.h, _lib.ptr(ids_h), B, T, _lib.ptr(im), _lib.ptr(sg), _lib.ptr(dp),
                                             max_new_tokens, _lib.ptr(out), _stream_ptr()))
        return out

    # ---- f1: sampling + per-token streaming --------------------------------------------------------------
    def set_sampling(self, do_sample: bool, temperature: float = 1.0, top_k: int = 50, top_p: float = 1.0,
                     seed: int = 0) -> None:
        """do_sample / temperature / top_k / top_p of the following prefill / decode / generate calls (HF sample())."""
        with self.lock:
            _lib.check(self.lib.vcb_set_sampling(self.h, int(bool(do_sample)), float(temperature), int(top_k),
                                                 float(top_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream_ptr()))

    def decode_stream(self, max_steps: int, on_token, lookahead: int = 4) -> int:
        """Up to `max_steps` decode steps; `on_token(step, ids: list[int]) -> bool` runs on the host for every token
        while the GPU already works on the next `lookahead` steps; returning True stops after that token (the KV
        length and RNG step are rolled back to it).  Returns the number of tokens delivered."""
        err = []

        def _cb(_user, step, ids_p, B):
            try:
                return 1 if on_token(int(step), [int(ids_p[i]) for i in range(B)]) else 0
            except BaseException as ex:  # never unwind through the C frames
                err.append(ex)
                return 1

        cb = _lib.TOKEN_CALLBACK(_cb)
        n = C.c_int32(0)
        with self.lock:
            rc = self.lib.vcb_decode_stream(self.h, int(max_steps), int(lookahead), cb, None, C.byref(n), _stream_ptr())
        if err:
            raise err[0]
        _lib.check(rc)
        return n.value

    def sample_op(self, logits: torch.Tensor, temperature: float, top_k: int, top_p: float, seed: int, step: int = 0):
        """Operator-level entry (parity tests): logits [B,V] fp16 on the device -> (next_ids [B], kept [B]) int32."""
        lg = logits.to(self.tdev, torch.float16).contiguous()
        B, V = lg.shape
        nxt = torch.empty(B, dtype=torch.int32, device=self.tdev)
        kept = torch.empty(B, dtype=torch.int32, device=self.tdev)
        _lib.check(self.lib.vcb_op_sample_f16(_lib.ptr(lg), B, V, float(temperature), int(top_k), float(top_p),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, int(step), _lib.ptr(nxt), _lib.ptr(kept),
                                              _stream_ptr()))
        return nxt, kept

    def reset(self):
        _lib.check(self.lib.vcb_reset(self.h))

    def _batch(self) -> int:
        return getattr(self, "_last_batch", None) or 0


def prefix_match(cached, plan, features_same: bool = True, limit: Optional[int] = None) -> int:
    """Host-only: vcb_prefix_match on two [B, len] int32 arrays of vcb_splice_plan entries (no GPU needed)."""
    lib = _lib.load()
    c = np.ascontiguousarray(np.asarray(cached, dtype=np.int32))
    p = np.ascontiguousarray(np.asarray(plan, dtype=np.int32))
    if c.ndim != 2 or p.ndim != 2 or c.shape[0] != p.shape[0]:
        raise ValueError("prefix_match: cached and plan must be [B, len] with the same B")
    out = C.c_int32(0)
    lim = p.shape[1] if limit is None else int(limit)
    _lib.check(lib.vcb_prefix_match(_lib.ptr(c), c.shape[1], _lib.ptr(p), p.shape[1], c.shape[0], int(features_same),
                                    lim, C.byref(out)))
    return out.value


def splice_plan(variant: int, ids, n_patches: int, has_seg: bool = True, depth_nonzero=None):
    """Host-only: the C++ splice planner (no GPU needed). Returns (plan [B, T'] int32 ndarray)."""
    lib = _lib.load()
    ids_np = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
    B, T = ids_np.shape
    cap = B * (T + 3 * n_patches)
    plan = np.empty(cap, dtype=np.int32)
    out_len = C.c_int32(0)
    nz = None
    if depth_nonzero is not None:
        nz = np.ascontiguousarray(np.asarray(depth_nonzero, dtype=np.uint8))
    _lib.check(lib.vcb_splice_plan(variant, _lib.ptr(ids_np), B, T, n_patches, int(has_seg), _lib.ptr(nz),
                                   _lib.ptr(plan), cap, C.byref(out_len)))
    return plan[: B * out_len.value].reshape(B, out_len.value)
