"""Per-kernel SASS evidence table from the built library: which kernels carry tcgen05 (UTC*MMA), TMEM loads/stores
(LDTM / STTM), TMA (UTMALDG / UBLKCP), legacy tensor-core (HMMA) and cp.async (LDGSTS) instructions.

    python tools/sass_ops.py [--out profiles/r02_sass_ops.csv]

Runs `cuobjdump -sass` on vcoder_b200/libvcoder_b200.so (no GPU needed).  Template instantiations of one kernel are summed
into one row (with the instantiation count)."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAPF", "HMMA", "LDGSTS", "MUFU.EX2", "ATOMG", "BAR.SYNC"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "vcoder_b200", "libvcoder_b200.so"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_ops.csv"))
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    inst = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        d = counts.setdefault(cur, collections.Counter())
        d["_total"] += 1
        for o in OPS:
            if o == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    d[o] += 1
            elif op == o or op.startswith(o + "."):
                d[o] += 1
    names = list(counts)
    if names:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    rows = collections.OrderedDict()
    for n, d in counts.items():
        full = demangle.get(n, n)
        base = re.sub(r"^void ", "", full).replace("(anonymous namespace)::", "").replace("vcb::", "")
        base = re.sub(r"<.*", "", base).split("(")[0]
        r = rows.setdefault(base, collections.Counter())
        r.update(d)
        inst[base] += 1
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(f"# tools/sass_ops.py: cuobjdump -sass {os.path.relpath(args.lib, ROOT)} ({os.path.getsize(args.lib)} bytes); counts summed over template instantiations\n")
        f.write("kernel,instantiations,sass_instructions," + ",".join(OPS) + "\n")
        for k, d in sorted(rows.items()):
            f.write(f"{k},{inst[k]},{d['_total']}," + ",".join(str(d[o]) for o in OPS) + "\n")
    print(open(args.out).read())


if __name__ == "__main__":
    sys.exit(main())
