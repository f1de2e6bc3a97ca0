"""Per-source-line instruction and stall-sample totals from an ncu capture (needs -lineinfo and --import-source on).
Usage: python tools/ncu_lines.py <file.ncu-rep> [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.defaultdict(lambda: [0, 0, ""])
cur_file = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); ism = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= iex: continue
    try: line = int(r[0])
    except ValueError: continue
    try: ex = int(r[iex] or 0); sm = int(r[ism] or 0)
    except ValueError: continue
    a = agg[(cur_file, line)]; a[0] += ex; a[1] += sm; a[2] = r[1].strip()[:90]
tot_ex = sum(a[0] for a in agg.values()); tot_sm = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot_ex}, samples {tot_sm}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:20s}:{l:5d} inst {100*a[0]/max(tot_ex,1):5.1f}%  samples {100*a[1]/max(tot_sm,1):5.1f}%  {a[2]}")
