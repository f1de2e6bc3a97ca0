#!/usr/bin/env python
"""Host-side cost of each piece of SG.forward's eval fast path (pinned inputs), each timed alone in a tight loop."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from sg_pr_b200.parser_sg import sgpr_args
from sg_pr_b200.sg_net import SG

state = B.load_state()
a = sgpr_args(); a.K, a.node_num, a.gpu, a.cuda = 20, 64, 0, "0"
model = SG(a, 12); model.load_state_dict(state); model.cuda(0).eval()
f1, f2 = B.build_batches(4, seed=3)
f1p, f2p = f1.pin_memory()[0], f2.pin_memory()[0]
eng = model.engine()
dev = eng.device
def t(fn, n=3000, sync_every=50):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn()
        if sync_every and i % sync_every == 0: torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / n * 1e6, 2)
out = {}
data = {"features_1": f1p, "features_2": f2p}
out["checks"] = t(lambda: (f1p.device.type == "cpu" and f2p.device.type == "cpu" and f1p.dtype == torch.float32 and f1p.dim() == 3 and f1p.shape[1] == 15 and f2p.shape == f1p.shape and f1p.is_contiguous() and f2p.is_contiguous() and f1p.is_pinned() and f2p.is_pinned()), sync_every=0)
out["weights_version"] = t(model._weights_version, sync_every=0)
out["3x torch.empty cuda"] = t(lambda: (torch.empty(128, dtype=torch.float32, device=dev), torch.empty((128, 64, 1), dtype=torch.float32, device=dev), torch.empty((128, 64, 1), dtype=torch.float32, device=dev)), sync_every=0)
out["current_stream + c_void_p"] = t(lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), sync_every=0)
s = torch.empty(128, device=dev); a1 = torch.empty(128, 64, 1, device=dev); a2 = torch.empty(128, 64, 1, device=dev)
st = eng._stream()
lib = eng._lib
out["ctypes sgpr_forward_pairs (launch, GPU kept shallow)"] = t(lambda: lib.sgpr_forward_pairs(eng._ctx, f1p.data_ptr(), f2p.data_ptr(), 128, 64, 20, s.data_ptr(), a1.data_ptr(), a2.data_ptr(), st), n=600, sync_every=1)
out["engine.forward_pairs"] = t(lambda: eng.forward_pairs(f1p, f2p, 20, True, True), n=600, sync_every=1)
ev = torch.cuda.Event()
out["event.record"] = t(lambda: ev.record(), sync_every=0)
out["event.query"] = t(lambda: ev.query(), sync_every=0)
with torch.no_grad():
    out["model.forward(data) (no nn.Module.__call__)"] = t(lambda: model.forward(data), n=600, sync_every=1)
    out["model(data)"] = t(lambda: model(data), n=600, sync_every=1)
out["torch.no_grad enter/exit"] = t(lambda: torch.no_grad().__enter__(), sync_every=0)
print(json.dumps(out))
