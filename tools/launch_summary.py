"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel mean time and share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "")[:58]
    agg.setdefault(name, []).append(float(r["Metric Value"]))
skip = ("distribution_elementwise", "vectorized_elementwise", "unrolled_elementwise")
steps = len(max(agg.values(), key=sum))   # launches of the kernel with the largest total time: one per step (one-off kernels of bench.py's parity checks aside)
tot = sum(sum(v) / steps for k, v in agg.items() if not any(s in k for s in skip))
print(f"{'kernel':60s} {'n':>4s} {'mean us':>9s} {'us/step':>9s} {'share':>7s}")
for k, v in agg.items():
    if any(s in k for s in skip):
        continue
    per_step = sum(v) / steps
    print(f"{k:60s} {len(v):4d} {sum(v)/len(v)/1e3:9.2f} {per_step/1e3:9.2f} {per_step/tot:7.1%}")
print(f"{'sum per step':60s} {'':4s} {'':9s} {tot/1e3:9.2f}")
