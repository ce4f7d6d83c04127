"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rNN_launches.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    k = r[kn].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")
print(f"\ntotal {tot / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches")
