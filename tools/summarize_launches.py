"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total us, share, average us."""
import collections
import csv
import re
import sys

rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("vcb::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    v = float(r[iv].replace(",", ""))
    if r[iu] == "ns":
        v /= 1e3
    elif r[iu] == "ms":
        v *= 1e3
    d = rows.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += v
tot = sum(d[1] for d in rows.values())
print("kernel,launches,total_us,share,avg_us")
for k, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{n},{us:.1f},{us / tot:.4f},{us / n:.2f}")
