"""Mean duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_summary.py <launches.csv> [steps]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
tot = 0.0
for k, v in agg.items():
    print(f"{k[:72]:72s} n={len(v):3d} mean={sum(v)/len(v)/1000:9.1f} us  per-step={sum(v)/steps/1000:9.1f} us")
    tot += sum(v) / steps / 1000
print(f"sum per step {tot:.1f} us")
