import ctypes as C, os, numpy as np, torch, time
os.environ["SGPR_EMBED_TC"] = "1"
from oracle import sgpr_oracle as orc
from sg_pr_b200 import _lib, synth
from sg_pr_b200.engine import Engine
from tests.emu import build_emu
lib = _lib.bind(C.CDLL(build_emu.build()), _lib.SYMBOLS)
sd = orc.load_state_npz('tests/golden/model_kitti.npz')
eng = Engine(lib=lib); eng.set_weights(sd)
del os.environ["SGPR_EMBED_TC"]
ref = Engine(lib=lib); ref.set_weights(sd)
for n,k,b in ((64,20,3),(40,10,2),(33,8,2)):
    f1, f2 = synth.make_pair_batch(b, n, k, seed=5)
    t=time.time()
    a = eng.forward_pairs(f1,f2,k)
    dt=time.time()-t
    r = ref.forward_pairs(f1,f2,k)
    want = orc.forward_pairs(f1,f2,k,sd)
    print(n,k,"tc vs ffma", float((a[0]-r[0]).abs().max()), "tc vs oracle", float((a[0]-want["score"]).abs().max()), "att", float((a[1]-want["att_1"]).abs().max()), f"{dt:.1f}s")
    e = eng.embed(f1, k, want_emb=True); eo = orc.embed_graphs(f1,k,sd)
    print("   emb", float((e["emb"]-eo["emb"]).abs().max()), "pooled", float((e["pooled"]-eo["pooled"].squeeze(-1)).abs().max()))
