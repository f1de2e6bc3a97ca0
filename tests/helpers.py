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
