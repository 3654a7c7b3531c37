"""Build libvcoder_b200.so (sm_100a only) in-tree with nvcc.

`python -m vcoder_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU, so this is also
the "does it build" check.  The library links cudart statically and resolves the one driver symbol it needs
(cuTensorMapEncodeTiled) at run time through cudaGetDriverEntryPoint, so it has no link-time libcuda dependency
and can be dlopen'ed (symbols inspected) on a machine without a driver.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "_obj"
LIB = PKG / "libvcoder_b200.so"
INCLUDE = PKG.parent / "include"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", str(INCLUDE), "-I", str(CSRC)]
# VCB_EXPERIMENTAL=1 at build time also compiles the A/B variants that were measured and rejected (packed weight stream,
# ring-shape sweep, tcgen05 skinny GEMM, one-tile ViT attention, bulk-copy / cluster decode attention): DESIGN.md section 3.3
if os.environ.get("VCB_EXPERIMENTAL", "0") not in ("", "0"):
    CFLAGS.append("-DVCB_EXPERIMENTAL=1")
# VCB_EXTRA_CFLAGS: extra nvcc flags for diagnostic builds (e.g. -DVCB_VIT_DIAG=1: timing switches inside the ViT attention)
CFLAGS += os.environ.get("VCB_EXTRA_CFLAGS", "").split()


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sources()
    hdrs = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    OBJ.mkdir(exist_ok=True)
    stamp = OBJ / "stamp.txt"
    hdr_digest = _digest(hdrs)

    def needs(src: Path) -> bool:
        obj = OBJ / (src.stem + ".o")
        tag = OBJ / (src.stem + ".tag")
        want = hdr_digest + _digest([src])
        return force or not obj.exists() or not tag.exists() or tag.read_text() != want

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        (OBJ / (src.stem + ".tag")).write_text(hdr_digest + _digest([src]))
        return obj

    todo = [s for s in srcs if needs(s)]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(compile_one, todo))
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    link_tag = _digest(objs)
    if todo or not LIB.exists() or not stamp.exists() or stamp.read_text() != link_tag:
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(link_tag)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
