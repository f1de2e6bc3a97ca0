"""CTA -> SM placement and per-CTA wall time of the fused kernel (needs the -DSGPR_TIMELINE build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth, _lib
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz("tests/golden/model_kitti.npz")
eng = Engine(0); eng.set_weights(sd)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
f1, f2 = synth.make_pair_batch(B, 64, 20, seed=1)
nreal = torch.stack([(f1[:, 3:].sum(1) > 0).sum(1), (f2[:, 3:].sum(1) > 0).sum(1)], 1).reshape(-1)   # graph g = 2b+side
f1, f2 = f1.cuda(), f2.cuda()
for _ in range(3): eng.forward_pairs(f1, f2, 20)
torch.cuda.synchronize()
lib = _lib.load()
smid = np.zeros(1024, dtype=np.int32); t = np.zeros(2048, dtype=np.int64)
lib.sgpr_debug_ctas.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.sgpr_debug_ctas(smid.ctypes.data, t.ctypes.data) == 0
G = 2 * B
t = t.reshape(1024, 2)[:G]; smid = smid[:G]
t0 = t[:, 0].min()
dur = (t[:, 1] - t[:, 0]) / 1e3
print("kernel span us:", (t[:, 1].max() - t0) / 1e3, " CTA start spread us:", (t[:, 0].max() - t0) / 1e3)
from collections import defaultdict
per = defaultdict(list)
for b in range(G): per[int(smid[b])].append(b)
sizes = sorted(len(v) for v in per.values())
print("SMs used", len(per), "CTAs/SM histogram", {k: sizes.count(k) for k in set(sizes)})
print("first 12 CTA->SM:", smid[:12].tolist(), " CTA 148..159 ->", smid[148:160].tolist() if G > 160 else "")
both = [v for v in per.values() if len(v) == 2]
print("pairs sharing an SM are (b, b+148)?", sum(1 for v in both if abs(v[0] - v[1]) == 148), "of", len(both))
print("per-CTA duration us: min %.1f mean %.1f max %.1f" % (dur.min(), dur.mean(), dur.max()))
solo = [dur[v[0]] for v in per.values() if len(v) == 1]
duo = [max(dur[v[0]], dur[v[1]]) for v in both]
if solo: print("solo CTAs: mean %.1f max %.1f" % (np.mean(solo), np.max(solo)))
if duo: print("shared SMs (max of the two): mean %.1f max %.1f min %.1f" % (np.mean(duo), np.max(duo), np.min(duo)))
