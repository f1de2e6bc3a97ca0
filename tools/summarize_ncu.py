"""Summarise an ncu --set full capture (.ncu-rep) + launch list csv into profiles/<tag>_*.{json,md,csv}.
Usage: python tools/summarize_ncu.py <tag> [prof.ncu-rep] [launches.csv] [bench.json]"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "prof_embed.ncu-rep")
launches = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "launches.csv")
bench = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", "bench.json")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__icc_request_hit_rate.pct", "lts__t_bytes.sum"]
summary = {"tag": tag, "source": os.path.basename(rep), "kernels": []}
for r in rows[2:]:
    k = {}
    for key in KEYS:
        for i, h in enumerate(H):
            if h == key or h.endswith("." + key):
                k[key] = r[i] + (" " + U[i] if U[i] else "")
                break
    stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i]) for i, h in enumerate(H)
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    k["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    summary["kernels"].append(k)

def to_bytes(s):
    v, unit = s.split()[:2]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
k0 = summary["kernels"][0]
summary["dram_bytes_per_launch"] = to_bytes(k0["dram__bytes_read.sum"]) + to_bytes(k0["dram__bytes_write.sum"])

# opcode mix from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
hdr = srows[1]; isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
byop, tot = collections.Counter(), 0
for r in srows[2:]:
    if len(r) < 10 or r[0] in ("Address",):
        continue
    if r[0] == "Kernel Name":
        break
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_]+)", r[isrc])
    byop[m.group(2) if m else "?"] += int(r[iex]); tot += int(r[iex])
summary["opcode_mix_pct"] = {op: round(100 * c / tot, 1) for op, c in byop.most_common(14)}
summary["sass_instructions_static"] = sum(1 for r in srows[2:] if len(r) > 10 and r[0] != "Address")

# launch list
lrows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(lrows) if r and r[0] == "ID"][0]
LH = lrows[hi]; ki, vi = LH.index("Kernel Name"), LH.index("Metric Value")
per = collections.defaultdict(list)
for r in lrows[hi + 1:]:
    if len(r) > vi:
        per[r[ki]].append(float(r[vi].replace(",", "")))
total = sum(sum(v) for v in per.values())
summary["launch_list"] = {k[:70]: {"launches": len(v), "avg_us": sum(v) / len(v) / 1e3, "share_of_gpu_time": sum(v) / total} for k, v in per.items()}
if os.path.exists(bench):
    try:
        summary["bench_line_same_build"] = json.loads(open(bench).read().strip().splitlines()[-1])
    except Exception:
        pass
json.dump(summary, open(os.path.join(out_dir, f"{tag}_ncu_summary.json"), "w"), indent=1)
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("\n".join(",".join(r) for r in lrows[hi:hi + 41]) + "\n")
print(json.dumps({k: summary[k] for k in ("dram_bytes_per_launch", "opcode_mix_pct", "launch_list")}, indent=1)[:1500])
