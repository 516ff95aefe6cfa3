#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [skip_first_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1 + skip:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = r[ik].split("(")[0].replace("void ", "").replace("vqcpc::", "")
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print(f"{len(rows) - 1 - skip} launches, {s / 1e3:.3f} ms total")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in tot.most_common(25):
    print(f"| {k[:90]} | {cnt[k]} | {v:.1f} | {100 * v / s:.1f}% |")
