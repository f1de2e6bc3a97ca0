"""Count k-NN set mismatches (beyond exact ties) between the CUDA kernel and the oracle, per layer, and show the
reference-distance gap at the k-th boundary for each mismatching row.  Usage: python tools/flip_stats.py [graphs] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n, k = 64, 20
sd = orc.load_state_npz("tests/golden/model_kitti.npz")
eng = Engine(0); eng.set_weights(sd)
tot_rows = 0; flips = [0] * 6; bad_graphs = set()
for start in range(0, M, 256):
    g = synth.make_graphs(min(256, M - start), n, k, seed=seed * 1000 + start)
    want = orc.embed_graphs(g, k, sd, want_trace=True)
    got = eng.embed(g.cuda(), k, want_att=True, want_emb=True, trace=True)
    knn = got["knn"].cpu().long()
    for layer in range(6):
        ok = orc.knn_sets_equivalent(want["knn_pd"][layer], want["knn_idx"][layer], knn[:, layer], want["layer_in"][layer])
        bad = (~ok).nonzero()
        flips[layer] += len(bad)
        for b, i in bad.tolist()[:6]:
            pd = want["knn_pd"][layer][b, i]
            srt = pd.sort(descending=True)[0]
            ref_set = set(want["knn_idx"][layer][b, i].tolist()); my_set = set(knn[b, layer, i].tolist())
            print(f"graph {start + b} layer {layer} row {i}: kth={srt[k-1].item():.9g} next={srt[k].item():.9g} "
                  f"gap={(srt[k-1]-srt[k]).item():.3g} rel={(srt[k-1]-srt[k]).item()/abs(srt[k-1].item()+1e-30):.3g} "
                  f"ref-only={sorted(ref_set - my_set)} mine-only={sorted(my_set - ref_set)}")
            bad_graphs.add(start + b)
    tot_rows += g.shape[0] * n * 6
    demb = (got["pooled"].cpu() - want["pooled"].squeeze(-1)).abs().amax(dim=1)
    print(f"[{start}] max pooled diff {float(demb.max()):.3g}; graphs with pooled diff > 1e-4: {(demb > 1e-4).nonzero().flatten().tolist()}")
print("rows", tot_rows, "flips per layer", flips, "graphs affected", sorted(bad_graphs))
