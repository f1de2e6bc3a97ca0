import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz("tests/golden/model_kitti.npz")
eng = Engine(0); eng.set_weights(sd)
for n, k in ((64, 20), (64, 10), (60, 20), (64, 19)):
    g = synth.make_graphs(3, n, k, seed=11)
    want = orc.embed_graphs(g, k, sd, want_trace=True)
    got = eng.embed(g.cuda(), k, want_att=True, want_emb=True, trace=True)
    layers = got["layers"].cpu()
    print(f"== N={n} k={k}")
    for layer in range(6):
        ref = want["layer_out"][layer].permute(0, 2, 1)
        d = (layers[:, layer, :, :ref.shape[2]] - ref).abs()
        bad_nodes = (d.amax(dim=(0, 2)) > 1e-4).nonzero().flatten().tolist()
        bad_ch = (d.amax(dim=(0, 1)) > 1e-4).nonzero().flatten().tolist()
        print(f" layer {layer}: max diff {float(d.max()):.3g}; bad nodes {bad_nodes[:70]}; bad ch {bad_ch[:70]}")
    print(" emb diff", float((got["emb"].cpu() - want["emb"]).abs().max()))
