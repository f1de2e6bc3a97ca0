import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz("tests/golden/model_kitti.npz")
eng = Engine(0); eng.set_weights(sd)
f1, f2 = synth.make_pair_batch(128, 64, 20, seed=3)
ref, _, _ = eng.forward_pairs(f1.cuda(), f2.cuda(), 20)
p1, p2 = f1.pin_memory(), f2.pin_memory()
z, _, _ = eng.forward_pairs(p1, p2, 20)
torch.cuda.synchronize()
print("zero-copy == device:", torch.equal(z, ref))
out = (torch.empty(128).pin_memory(), None, None)
for name, fn in (("host zero-copy", lambda: eng.forward_pairs_host(p1, p2, 20, want_att=False, out=out)),
                 ("host staged (pageable)", lambda: eng.forward_pairs_host(f1, f2, 20, want_att=False))):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(300): fn()
    print(name, "us/call:", (time.perf_counter() - t0) / 300 * 1e6)
os.environ["X"] = "1"
