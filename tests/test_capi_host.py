"""CPU-side checks of the C ABI: the library loads, exports every symbol include/vcoder_b200.h declares, refuses to
compute without a GPU (no CPU fallback), and its host-only splice planner agrees with the oracle's restatement of
prepare_inputs_labels_for_multimodal (which tests/test_oracle_golden.py pins to the reference)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import vcoder_oracle as orc
from vcoder_b200 import _lib, synth
from vcoder_b200.engine import splice_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I, S, D = synth.IMAGE_TOKEN_INDEX, synth.SEG_TOKEN_INDEX, synth.DEPTH_TOKEN_INDEX


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vcoder_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vcoder_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "python binding and header disagree"
    assert b"sm_100a" in lib.vcb_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.vcb_device_count() == 0
    cfg = _lib.VcbConfig(struct_size=C.sizeof(_lib.VcbConfig))
    h = C.c_void_p()
    assert lib.vcb_create(C.byref(cfg), 0, C.byref(h)) == -7  # VCB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.vcb_last_error()
    x = np.zeros((8, 64), dtype=np.float16)
    rc = lib.vcb_op_gemm_f16(_lib.ptr(x), 64, _lib.ptr(x), 64, 8, 8, 64, None, None, 0, 0, _lib.ptr(x), 64, 0, None)
    assert rc == -7
    from vcoder_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(synth.PRESETS["tiny"])


def _oracle_plan_rows(ids, ds, has_seg, nz, P, B):
    """Expand the oracle's segment plan to the flat row encoding of vcb_splice_plan."""
    plans = orc.splice_plan(ids, ds, has_seg, nz)
    kind = {"image": 0, "seg": 1, "depth": 2}
    rows = []
    for segs in plans:
        r = []
        for k, payload in segs:
            if k == "text":
                r += list(payload)
            else:
                base = (kind[k] * B + payload) * (P + 1) + 1
                r += [-(base + p) - 2 for p in range(P)]
        rows.append(r)
    return rows


CASES = [
    # (variant_ds, ids, has_seg, depth_nonzero)
    (True, [[1, 5, I, D, S, 9], [1, 6, I, D, S, 7]], True, [True, True]),   # reference tokenizer order (F5)
    (True, [[1, I, 7, S, 8, D, 9]], True, [True]),                          # depth after seg: spliced
    (True, [[1, I, 9, 4]], True, [True]),                                   # DS tolerates image-only
    (True, [[1, 2, 3, 4, 5]], True, None),                                  # text only
    (False, [[1, I, 7, S, 9], [2, I, 8, S, 3]], True, None),                # VCoder keeps text between
    (True, [[1, I, S, 4, I, 5]], True, [True]),                             # second image after seg is never consumed? (error)
    (True, [[I, S, D]], True, [True]),
]


@pytest.mark.parametrize("ds,ids,has_seg,nz", CASES)
def test_splice_plan_matches_oracle(ds, ids, has_seg, nz):
    P = 4
    B = len(ids)
    variant = synth.VARIANT_VCODER_DS if ds else synth.VARIANT_VCODER
    try:
        want = _oracle_plan_rows(ids, ds, has_seg, nz, P, B)
        want_err = None
    except orc.SpliceError as e:
        want, want_err = None, e
    if want is not None and len({len(r) for r in want}) > 1:
        want, want_err = None, "ragged"
    try:
        got = splice_plan(variant, ids, P, has_seg, nz)
    except _lib.VcbError as e:
        assert want is None, f"C planner failed ({e}) where the oracle succeeded"
        assert e.code in (-5, -6, -1)
        return
    assert want is not None, f"oracle raised {want_err} but the C planner returned a plan"
    assert got.tolist() == want


def test_splice_plan_error_codes():
    # VCoder with <image> but no <seg>: dummy branch feeds -200 to nn.Embedding -> IndexError (F11)
    with pytest.raises(_lib.VcbError) as ei:
        splice_plan(synth.VARIANT_VCODER, [[1, I, 9]], 4, True, None)
    assert ei.value.code == -6
    # is_depth_zero with a reachable <depth> token: the id reaches the embedding
    with pytest.raises(_lib.VcbError) as ei:
        splice_plan(synth.VARIANT_VCODER_DS, [[1, I, 7, S, 8, D, 9]], 4, True, [False])
    assert ei.value.code == -6
    # ragged batch: the reference dies with NameError (_new_labels)
    with pytest.raises(_lib.VcbError) as ei:
        splice_plan(synth.VARIANT_VCODER_DS, [[1, I, S, 9], [1, 5, 5, 9]], 4, True, [True, True])
    assert ei.value.code == -5


def test_splice_plan_random_fuzz():
    rng = np.random.RandomState(0)
    P = 3
    n_ok = 0
    for trial in range(300):
        ds = bool(rng.randint(2))
        B = rng.randint(1, 4)
        T = rng.randint(3, 9)
        ids = rng.randint(3, 50, size=(B, T)).astype(np.int64)
        # same special layout in every row so lengths agree
        specials = [I, S] if not ds else [[I, D, S], [I, S, D], [I, S], [I]][rng.randint(4)]
        pos = sorted(rng.choice(np.arange(1, T), size=min(len(specials), T - 1), replace=False))
        for b in range(B):
            for p, s in zip(pos, specials):
                ids[b, p] = s
        nz = [bool(rng.randint(2)) for _ in range(B)] if ds else None
        variant = synth.VARIANT_VCODER_DS if ds else synth.VARIANT_VCODER
        try:
            want = _oracle_plan_rows(ids.tolist(), ds, True, nz, P, B)
        except orc.SpliceError:
            want = None
        if want is not None and len({len(r) for r in want}) > 1:
            want = None
        try:
            got = splice_plan(variant, ids, P, True, nz).tolist()
        except _lib.VcbError:
            got = None
        assert got == want, (ds, ids.tolist(), nz)
        n_ok += want is not None
    assert n_ok > 50


# ---- plain LLaVA (llava_arch.py:117-199) and VCoder-IT (vcoder_it_llava_arch.py:165-277) splice rules ----
def _oracle_rows_variant(ids, variant, has_seg, nz, P, B):
    plans = orc.splice_plan(ids, False, has_seg, nz if variant == synth.VARIANT_VCODER_DS else None, variant=variant,
                            seg_nonzero=nz if variant == synth.VARIANT_VCODER_IT else None)
    kind = {"image": 0, "seg": 1, "depth": 2}
    rows = []
    for segs in plans:
        r = []
        for k, payload in segs:
            if k == "text":
                r += list(payload)
            else:
                if payload >= B:  # image_features[cur_image_idx] past the batch: the reference's IndexError
                    raise IndexError(payload)
                base = (kind[k] * B + payload) * (P + 1) + 1
                r += [-(base + p) - 2 for p in range(P)]
        rows.append(r)
    return rows


def test_splice_plan_llava_and_it_known_cases():
    P = 4
    L, IT = synth.VARIANT_LLAVA, synth.VARIANT_VCODER_IT
    img = [-(1 + p) - 2 for p in range(P)]
    seg = [-((1 * 1 + 0) * (P + 1) + 1 + p) - 2 for p in range(P)]
    assert splice_plan(L, [[1, I, 7, 9]], P, False, None).tolist() == [[1] + img + [7, 9]]
    assert splice_plan(L, [[1, 2, 3]], P, False, None).tolist() == [[1, 2, 3]]  # no <image>: plain text
    with pytest.raises(_lib.VcbError) as ei:  # LLaVA never consumes <seg>: it reaches embed_tokens
        splice_plan(L, [[1, I, S, 9]], P, False, None)
    assert ei.value.code == -6
    # IT: ids between <image> and <seg> are dropped; zero seg map skips the seg splice
    assert splice_plan(IT, [[1, I, 7, S, 9]], P, True, [True]).tolist() == [[1] + img + seg + [9]]
    assert splice_plan(IT, [[1, I, 7, 9]], P, True, [True]).tolist() == [[1] + img + [7, 9]]
    assert splice_plan(IT, [[1, I, 7, 9]], P, True, [False]).tolist() == [[1] + img + [7, 9]]
    assert splice_plan(IT, [[1, I, 7, 9]], P, False, None).tolist() == [[1] + img + [7, 9]]
    with pytest.raises(_lib.VcbError) as ei:
        splice_plan(IT, [[1, I, S, 9]], P, True, [False])
    assert ei.value.code == -6
    with pytest.raises(_lib.VcbError) as ei:  # segs=None is treated as all-zero (:158)
        splice_plan(IT, [[1, I, S, 9]], P, False, None)
    assert ei.value.code == -6


def test_splice_plan_fuzz_llava_and_it():
    rng = np.random.RandomState(1)
    P = 3
    n_ok = 0
    for trial in range(300):
        variant = [synth.VARIANT_LLAVA, synth.VARIANT_VCODER_IT][rng.randint(2)]
        B = rng.randint(1, 4)
        T = rng.randint(3, 9)
        ids = rng.randint(3, 50, size=(B, T)).astype(np.int64)
        specials = [[I], [I, S], [I, I], [S, I], [I, S, S], []][rng.randint(6)]
        pos = sorted(rng.choice(np.arange(1, T), size=min(len(specials), T - 1), replace=False))
        for b in range(B):
            for p, s in zip(pos, specials):
                ids[b, p] = s
        has_seg = variant == synth.VARIANT_VCODER_IT and bool(rng.randint(4))
        nz = [bool(rng.randint(2)) for _ in range(B)] if has_seg else None
        try:
            want = _oracle_rows_variant(ids.tolist(), variant, has_seg, nz, P, B)
        except (orc.SpliceError, IndexError):
            want = None
        if want is not None and len({len(r) for r in want}) > 1:
            want = None
        try:
            got = splice_plan(variant, ids, P, has_seg, nz).tolist()
        except _lib.VcbError:
            got = None
        assert got == want, (variant, ids.tolist(), has_seg, nz)
        n_ok += want is not None
    assert n_ok > 50


# ---------------------------------------------------------------------------------------------------------------
# f3: the host half of the multi-turn prefix cache (vcb_prefix_match) against a plain-Python restatement
# ---------------------------------------------------------------------------------------------------------------
def _prefix_match_py(cached, plan, features_same, limit):
    best = min(len(cached[0]), len(plan[0]), limit)
    for c, p in zip(cached, plan):
        i = 0
        while i < best and c[i] == p[i] and (features_same or p[i] > -2):
            i += 1
        best = i
    return max(best, 0)


def test_prefix_match_host_cases():
    from vcoder_b200.engine import prefix_match
    P = 4
    plan1 = splice_plan(synth.VARIANT_VCODER_DS, [[1, 5, I, D, S, 9]], P)           # turn 1
    plan2 = splice_plan(synth.VARIANT_VCODER_DS, [[1, 5, I, D, S, 9, 40, 41, 7]], P)  # turn 2: + answer + question
    cached = np.concatenate([plan1, np.array([[40]], dtype=np.int32)], axis=1)      # one generated token fed back
    T1 = plan1.shape[1]
    assert T1 == 2 + 2 * P + 1
    assert prefix_match(cached, plan2) == T1 + 1                       # prompt + the fed token
    assert prefix_match(cached, plan2, features_same=False) == 2      # stops at the first visual row
    assert prefix_match(cached, plan2, limit=5) == 5
    assert prefix_match(plan1, plan1, limit=T1 - 1) == T1 - 1         # regenerate: the last position always runs
    assert prefix_match(plan1[:, :0], plan2) == 0                      # nothing cached
    # batch: the shortest common prefix wins
    a = np.array([[1, 2, 3, 4, 5], [1, 2, 3, 4, 5]], dtype=np.int32)
    b = np.array([[1, 2, 3, 4, 5, 6], [1, 2, 9, 4, 5, 6]], dtype=np.int32)
    assert prefix_match(a, b) == 2
    with pytest.raises(ValueError):
        prefix_match(a, b[:1])


def test_prefix_match_host_random():
    from vcoder_b200.engine import prefix_match
    rng = np.random.default_rng(0)
    for _ in range(200):
        B = int(rng.integers(1, 4))
        lc, lp = int(rng.integers(0, 12)), int(rng.integers(1, 12))
        base = rng.integers(-6, 5, size=(B, max(lc, lp))).astype(np.int32)
        cached, plan = base[:, :lc].copy(), base[:, :lp].copy()
        for b in range(B):  # inject a divergence in some rows
            if rng.random() < 0.5 and lp > 0:
                plan[b, int(rng.integers(0, lp))] += 1
        fs = bool(rng.integers(0, 2))
        limit = int(rng.integers(0, lp + 2))
        assert prefix_match(cached, plan, fs, limit) == _prefix_match_py(cached.tolist(), plan.tolist(), fs, limit)
