"""Test helpers: materialise the committed golden fixtures as the files the reference's scripts expect."""
import json
import os

import numpy as np
import torch


def write_fixture_tree(golden_dir, root):
    """Creates <root>/data/{0,3,250}.json, <root>/model/model.pth (DataParallel-style keys) and
    <root>/config/config.yml equivalent to the reference's shipped tree.  Returns the config path."""
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    os.makedirs(os.path.join(root, "model"), exist_ok=True)
    os.makedirs(os.path.join(root, "config"), exist_ok=True)
    with np.load(os.path.join(golden_dir, "fixtures.npz")) as z:
        for name in ("0", "3", "250"):
            with open(os.path.join(root, "data", f"{name}.json"), "w") as f:
                json.dump({"centers": z[f"centers_{name}"].tolist(), "nodes": z[f"nodes_{name}"].tolist(),
                           "pose": z[f"pose_{name}"].tolist()}, f)
    with np.load(os.path.join(golden_dir, "model_kitti.npz")) as z:
        state = {"module." + k: torch.from_numpy(z[k].copy()) for k in z.files}
    torch.save(state, os.path.join(root, "model", "model.pth"))
    cfg = os.path.join(root, "config", "config.yml")
    with open(cfg, "w") as f:
        f.write(f"""common:
  model: "{root}/model/model.pth"
  cuda: "0"
  batch_size: 128
  p_thresh: 3
  graph_pairs_dir: "{root}/data"
  pair_list_dir: "{root}/lists"
arch:
  keep_node: 1
  filters_1: 64
  filters_2: 64
  filters_3: 32
  tensor_neurons: 16
  bottle_neck_neurons: 16
  K: 10
train:
  epochs: 2
  train_sequences: ['00']
  eval_sequences: ['08']
  dropout: 0
  learning_rate: 0.001
  weight_decay: 0.0005
  gpu: 0
  logdir: "{root}/logs"
  node_num: 100
eva_batch:
  sequences: ["00"]
  output_path: "{root}/eva"
  show: False
eva_pair:
  pair_file: ["{root}/data/0.json", "{root}/data/250.json"]
""")
    return cfg


NEAR_TIE_REL = 2e-6     # a k-th / (k+1)-th reference distance closer than this (relative) is decided by sgemm rounding


def near_tie_flips(eng, state, graphs, k):
    """k-NN rows where the kernel and the oracle pick different (non-equivalent) sets, with the relative gap of the
    reference distances at the k-th boundary: [(graph, layer, row, rel_gap)].  SURVEY §7 hard part 2: such a row
    flips with the last bit of the Gram matrix — in the reference too (MKL vs cuBLAS vs fp64 disagree on them) —
    and moves the score by up to ~1e-2; it is classified, not hidden."""
    from oracle import sgpr_oracle as orc
    want = orc.embed_graphs(graphs, k, state, want_trace=True)
    got = eng.embed(graphs.cuda(), k, trace=True)
    knn = got["knn"].cpu().long()
    out = []
    for layer in range(6):
        ok = orc.knn_sets_equivalent(want["knn_pd"][layer], want["knn_idx"][layer], knn[:, layer], want["layer_in"][layer])
        for b, i in (~ok).nonzero().tolist():
            srt = want["knn_pd"][layer][b, i].sort(descending=True)[0]
            out.append((b, layer, i, float((srt[k - 1] - srt[k]).abs() / srt[k - 1].abs().clamp_min(1e-30))))
    return out


def assert_scores_match_or_near_tie(eng, state, f1, f2, k, got, want, tol=1e-5, max_bad=2):
    """Every score within tol of the oracle, except (at most max_bad) pairs whose deviation is explained row by row by
    k-NN near-ties (reference gap < NEAR_TIE_REL).  Returns the indices of the explained pairs."""
    import torch
    err = (got.detach().cpu() - want).abs()
    bad = (err > tol).nonzero().flatten().tolist()
    assert len(bad) <= max_bad, f"{len(bad)} of {len(err)} pairs off by more than {tol}: {bad}"
    for p in bad:
        flips = near_tie_flips(eng, state, torch.stack([f1[p], f2[p]]).cpu(), k)
        # only the FIRST divergence of each graph/branch has to be a near-tie: once one neighbour set differs, the
        # features of the later layers of that branch (xyz = layers 0-2, sem = 3-5) legitimately differ too
        first = {}
        for b, layer, row, gap in flips:
            key = (b, layer // 3)
            if key not in first or layer < first[key][0]:
                first[key] = (layer, [])
            if layer == first[key][0]:
                first[key][1].append(gap)
        assert first and all(g < NEAR_TIE_REL for _, gaps in first.values() for g in gaps), \
            f"pair {p}: |dscore| {float(err[p]):.3g} not explained by a k-NN near-tie: {flips}"
    return bad
