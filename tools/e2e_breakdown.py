#!/usr/bin/env python
"""Where the end-to-end step (SG.forward on pinned host tensors + .cpu()) spends its time: host time until the launch call
returns, device time of the launch with the inputs read over PCIe, the .cpu() read-back.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from sg_pr_b200 import synth
from sg_pr_b200.parser_sg import sgpr_args
from sg_pr_b200.sg_net import SG

state = B.load_state()
a = sgpr_args(); a.K, a.node_num, a.gpu, a.cuda = 20, 64, 0, "0"
model = SG(a, 12); model.load_state_dict(state); model.cuda(0).eval()
f1, f2 = B.build_batches(170, seed=3)
f1p, f2p = f1.pin_memory(), f2.pin_memory()
f1d, f2d = f1.cuda(), f2.cuda()
eng = model.engine()
n = 300
def loop(fn):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(n): fn(20 + i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6
out = {}
def full(i):
    j = i % 170
    with torch.no_grad():
        p, _, _ = model({"features_1": f1p[j], "features_2": f2p[j]})
    return p.cpu()
out["e2e_us"] = loop(full)
def nosync(i):
    j = i % 170
    with torch.no_grad():
        p, _, _ = model({"features_1": f1p[j], "features_2": f2p[j]})
    return p
out["module_call_async_us (launch-bound rate, no read-back)"] = loop(nosync)
host = []
def host_only(i):
    j = i % 170
    t = time.perf_counter()
    with torch.no_grad():
        p, _, _ = model({"features_1": f1p[j], "features_2": f2p[j]})
    host.append(time.perf_counter() - t)
    torch.cuda.synchronize()
    return p
loop(host_only)
out["host_time_inside_module_call_us"] = sum(host[20:]) / len(host[20:]) * 1e6
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def kern(src1, src2):
    tot = 0.0
    for i in range(20, 20 + n):
        j = i % 170
        ev[0].record(); eng.forward_pairs(src1[j], src2[j], 20); ev[1].record(); torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[1])
    return tot / n * 1e3
out["kernel_us_inputs_in_hbm (isolated launches)"] = kern(f1d, f2d)
out["kernel_us_inputs_pinned_host (zero-copy over PCIe)"] = kern(f1p, f2p)
p, _, _ = model({"features_1": f1p[0], "features_2": f2p[0]}); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(n): p.cpu()
out["cpu_readback_of_a_ready_score_us"] = (time.perf_counter() - t) / n * 1e6
t = time.perf_counter()
for i in range(n): x = f1p[i % 170]; y = f2p[i % 170]
out["batch_indexing_us (bench harness)"] = (time.perf_counter() - t) / n * 1e6
print(json.dumps(out))
