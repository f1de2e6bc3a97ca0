"""Collect one GPU trip of tools/gpu_train_round.sh into profiles/<tag>_train_*: bench lines, per-kernel launch times,
and the key ncu metrics + hottest source lines of the EdgeConv forward / backward captures.
Usage: python tools/summarize_train.py <tag>"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1]
prof = os.path.join(ROOT, "profiles")

shutil.copy(os.path.join(OUT, "train_bench.jsonl"), os.path.join(prof, f"{tag}_train_bench.jsonl"))
shutil.copy(os.path.join(OUT, "train_launches.csv"), os.path.join(prof, f"{tag}_train_launches.csv"))

rows = list(csv.reader(open(os.path.join(OUT, "train_launches.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        per.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) / 1000)
steps = len(per.get("sgpr_train_adam_kernel(TrainWs, AdamArgs)", [1]))
summary = {"tag": tag, "steps_in_launch_list": steps,
           "kernel_us_per_step": {k: round(sum(v) / steps, 1) for k, v in per.items()},
           "note": "ncu launch list: cold-cache, serialised launches; shares of the step, not absolute times"}
summary["sum_us_per_step"] = round(sum(summary["kernel_us_per_step"].values()), 1)

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sass__inst_executed_local_loads", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for which in ("fwd", "bwd"):
    rep = os.path.join(OUT, f"prof_train_{which}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    H, U, R = rr[0], rr[1], rr[2]
    k = {"kernel": R[H.index("Kernel Name")]}
    for i, name in enumerate(H):
        if name in WANT:
            k[name] = (R[i] + " " + U[i]).strip()
    st = {n.split("issue_stalled_")[1].split("_per_issue")[0]: float(R[i]) for i, n in enumerate(H)
          if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")}
    k["warp_stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "12"], capture_output=True, text=True).stdout
    k["hot_lines"] = lines.strip().splitlines()
    summary[f"edge_{which}_layer2"] = k
json.dump(summary, open(os.path.join(prof, f"{tag}_train_summary.json"), "w"), indent=1)
print(json.dumps(summary["kernel_us_per_step"], indent=1), summary["sum_us_per_step"])
