"""Per-phase clock stamps of CTA 0 from a -DSGPR_TIMELINE build (tools/variants/lib_timeline.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth, _lib
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz("tests/golden/model_kitti.npz")
eng = Engine(0); eng.set_weights(sd)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
f1, f2 = synth.make_pair_batch(B, 64, 20, seed=1)
f1, f2 = f1.cuda(), f2.cuda()
for _ in range(3): eng.forward_pairs(f1, f2, 20)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros(8 * 128, dtype=np.int64)
lib.sgpr_debug_timeline.argtypes = [ctypes.c_void_p]
assert lib.sgpr_debug_timeline(buf.ctypes.data) == 0
t = buf.reshape(8, 128)
t0 = t[:, 0].min()
names = {0: "start", 1: "input landed", 58: "layers done", 60: "attention done", 62: "end"}
for l in range(6):
    for j, nm in enumerate(["front start", "gram done", "select done", "front done", "barrier B", "back done", "barrier A"]):
        names[8 + l * 8 + j] = f"L{l} {nm}"
print("n_real side0 graph0:", int((f1[0, 3:].sum(0) > 0).sum()))
print(f"{'slot':18s}" + "".join(f"   w{w}" .rjust(9) for w in range(8)))
prev = None
for slot in sorted(names):
    row = t[:, slot] - t0
    if (t[:, slot] == 0).all(): continue
    print(f"{names[slot]:18s}" + "".join(f"{int(v):9d}" for v in row))

# aggregated phase durations (cycles), mean over warps
import numpy as _np
def d(a, b):
    return float(_np.mean(t[:, b] - t[:, a]))
agg = {"gram": 0.0, "select": 0.0, "gemm/xyz": 0.0, "barrierB": 0.0, "back": 0.0, "barrierA+next": 0.0}
for l in range(6):
    s0 = 8 + l * 8
    agg["gram"] += d(s0, s0 + 1); agg["select"] += d(s0 + 1, s0 + 2); agg["gemm/xyz"] += d(s0 + 2, s0 + 3)
    if l: agg["barrierB"] += d(s0 + 3, s0 + 4); agg["back"] += d(s0 + 4, s0 + 5)
    else: agg["back"] += d(s0 + 3, s0 + 5)
    nxt = 8 + (l + 1) * 8 if l < 5 else 58
    agg["barrierA+next"] += d(s0 + 5, nxt)
agg["attention"] = d(58, 60); agg["head+end"] = d(60, 62); agg["total"] = d(0, 62)
print("PHASES " + " ".join(f"{k}={v:.0f}" for k, v in agg.items()))
