"""CTA -> SM placement, per-CTA wall time and graph weight of the fused kernel (needs a -DSGPR_TIMELINE build:
SGPR_EXTRA_NVCC_FLAGS=-DSGPR_TIMELINE SGPR_BUILD_TAG=tl python -m sg_pr_b200.build, then
SGPR_B200_LIB=tools/variants/libsgpr_b200_tl.so python tools/cta_map.py 128)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import defaultdict
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth, _lib
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz("tests/golden/model_kitti.npz")
eng = Engine(0); eng.set_weights(sd)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
f1, f2 = synth.make_pair_batch(B, 64, 20, seed=1)
f1, f2 = f1.cuda(), f2.cuda()
for _ in range(3): eng.forward_pairs(f1, f2, 20)
torch.cuda.synchronize()
lib = _lib.load()
smid = np.zeros(1024, dtype=np.int32); t = np.zeros(2048, dtype=np.int64); gg = np.zeros(1024, dtype=np.int32)
lib.sgpr_debug_ctas.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.sgpr_debug_cta_graphs.argtypes = [ctypes.c_void_p]
assert lib.sgpr_debug_ctas(smid.ctypes.data, t.ctypes.data) == 0
assert lib.sgpr_debug_cta_graphs(gg.ctypes.data) == 0
G = 2 * B
t = t.reshape(1024, 2)[:G]; smid = smid[:G]; rows = (gg[:G] >> 16); graph = gg[:G] & 0xffff
assert sorted(graph.tolist()) == list(range(G)), "every graph exactly once"
t0 = t[:, 0].min()
dur = (t[:, 1] - t[:, 0]) / 1e3
end = (t[:, 1] - t0) / 1e3
print("smid range", int(smid.min()), int(smid.max()), "distinct", len(set(smid.tolist())))
print("kernel span us: %.1f   CTA start spread us: %.1f" % ((t[:, 1].max() - t0) / 1e3, (t[:, 0].max() - t0) / 1e3))
per = defaultdict(list)
for b in range(G): per[int(smid[b])].append(b)
sizes = sorted(len(v) for v in per.values())
print("SMs used", len(per), "CTAs/SM histogram", {k: sizes.count(k) for k in set(sizes)})
solo = [v[0] for v in per.values() if len(v) == 1]
both = [v for v in per.values() if len(v) == 2]
print("per-CTA duration us: min %.1f mean %.1f max %.1f" % (dur.min(), dur.mean(), dur.max()))
if solo:
    print("lone CTAs  : rows mean %.1f [%d..%d]  end us mean %.1f max %.1f" % (rows[solo].mean(), rows[solo].min(), rows[solo].max(), end[solo].mean(), end[solo].max()))
if both:
    e2 = np.array([max(end[v[0]], end[v[1]]) for v in both]); r2 = np.array([rows[v[0]] + rows[v[1]] for v in both])
    print("shared SMs : rows-sum mean %.1f [%d..%d]  end us mean %.1f max %.1f min %.1f" % (r2.mean(), r2.min(), r2.max(), e2.mean(), e2.max(), e2.min()))
    print("corr(rows-sum, end) on shared SMs: %.2f" % np.corrcoef(r2, e2)[0, 1])
# duration of a lone CTA / shared CTA by rows-per-warp class
for name, idx in (("lone", solo), ("shared", [b for v in both for b in v])):
    if not idx: continue
    idx = np.array(idx)
    for lo, hi in ((0, 32), (33, 40), (41, 64)):
        m = (rows[idx] >= lo) & (rows[idx] <= hi)
        if m.any(): print("  %-6s rows %2d..%2d: n=%3d dur us mean %.1f" % (name, lo, hi, m.sum(), dur[idx][m].mean()))
