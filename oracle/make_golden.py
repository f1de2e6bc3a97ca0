"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python oracle/make_golden.py
TEST INFRASTRUCTURE — never imported by the product.  The reference has no tests and records no expected
outputs, so these files ARE the pin: inputs + what the reference's own `SG.forward`
(sg_net.py:112-138) returned for them, with the k-NN index tensors its `dgcnn.knn` (dgcnn.py:14-20) produced
and each EdgeConv layer's output, captured by wrapping — not editing — the reference functions.

Files written:
  model_kitti.npz            the shipped checkpoint model/model.pth (keys without the 'module.' prefix)
  model_<tag>.npz            two more checkpoints out of model/release_model.zip
  fixtures.npz               data/{0,3,250}.json as arrays (centers / nodes / pose)
  ref_fixture_pairs.npz      6 ordered fixture pairs x {(K10,N100),(K20,N64)}: features + score/att (+trace)
  ref_synth_<cfg>.npz        seeded KITTI-shape batches at several (N,k): features + score/att + trace
  ref_ckpt_scores.npz        scores of one synthetic batch under the extra checkpoints
"""
from __future__ import annotations

import io
import json
import os
import sys
import zipfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from sg_pr_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = ref_shim.REFERENCE_ROOT


def state_to_npz(state_dict, path):
    arrays = {}
    for name, value in state_dict.items():
        arrays[name[7:] if name.startswith("module.") else name] = value.detach().cpu().numpy()
    np.savez_compressed(path, **arrays)


class Tracer:
    """Wraps dgcnn.knn and hooks the six EdgeConv blocks of the reference model to capture intermediates."""

    def __init__(self, ref, module):
        self.ref, self.module = ref, module
        self.knn, self.layers = [], []
        self._orig_knn = ref.dgcnn.knn
        self._hooks = []

    def __enter__(self):
        def knn(x, k):
            idx = self._orig_knn(x, k)
            self.knn.append(idx.clone())
            return idx

        self.ref.dgcnn.knn = knn
        for name in ("dgcnn_s_conv1", "dgcnn_s_conv2", "dgcnn_s_conv3",
                     "dgcnn_f_conv1", "dgcnn_f_conv2", "dgcnn_f_conv3"):
            block = getattr(self.module, name)
            self._hooks.append(block.register_forward_hook(
                lambda m, i, o: self.layers.append(o.max(dim=-1)[0].clone())))
        return self

    def __exit__(self, *exc):
        self.ref.dgcnn.knn = self._orig_knn
        for h in self._hooks:
            h.remove()


@torch.no_grad()
def run_reference(trainer, f1, f2, trace=True):
    ref = ref_shim.load_reference()
    module = ref_shim.reference_module(trainer)
    data = {"features_1": f1, "features_2": f2, "target": torch.zeros(f1.shape[0])}
    emb = []
    orig_pass = module.dgcnn_conv_pass

    def conv_pass(x):
        out = orig_pass(x)
        emb.append(out.clone())
        return out

    module.dgcnn_conv_pass = conv_pass
    try:
        with Tracer(ref, module) as t:
            score, a1, a2 = module(data)
    finally:
        module.dgcnn_conv_pass = orig_pass
    out = {"features_1": f1.numpy(), "features_2": f2.numpy(), "score": score.numpy(),
           "att_1": a1.numpy(), "att_2": a2.numpy(), "emb_1": emb[0].numpy(), "emb_2": emb[1].numpy()}
    if trace:
        # call order inside dgcnn_conv_pass: xyz layers 1-3 then sem layers 1-3 (sg_net.py:84-102), side 1 then 2
        for side in (0, 1):
            for layer in range(6):
                out[f"knn_idx_{side + 1}_{layer}"] = t.knn[side * 6 + layer].numpy().astype(np.uint8)
                out[f"layer_out_{side + 1}_{layer}"] = t.layers[side * 6 + layer].numpy()
    return out


def fixture_features(trainer, ref, a, b):
    """The reference's own host prep: utils.process_pair (utils.py:21-38) + transfer_to_torch (sg_net.py:241-310)."""
    pair = [os.path.join(REF, "data", f"{a}.json"), os.path.join(REF, "data", f"{b}.json")]
    data = trainer.transfer_to_torch(ref.utils.process_pair(pair), False)
    f1 = torch.FloatTensor(np.array([data["features_1"]]))
    f2 = torch.FloatTensor(np.array([data["features_2"]]))
    return f1, f2, data["target"]


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    ref = ref_shim.load_reference()

    # --- checkpoints -------------------------------------------------------------------------------
    shipped = torch.load(os.path.join(REF, "model", "model.pth"))
    state_to_npz(shipped, os.path.join(OUT, "model_kitti.npz"))
    extra = {}
    with zipfile.ZipFile(os.path.join(REF, "model", "release_model.zip")) as z:
        for tag, member in (("3_20_08", "release_model/3_20/08/model.pth"),
                            ("10_20_05", "release_model/10_20/05/model.pth")):
            sd = torch.load(io.BytesIO(z.read(member)))
            state_to_npz(sd, os.path.join(OUT, f"model_{tag}.npz"))
            tmp = f"/tmp/sgpr_ckpt_{tag}.pth"
            torch.save(sd, tmp)
            extra[tag] = tmp

    # --- raw fixtures ------------------------------------------------------------------------------
    fx = {}
    for name in ("0", "3", "250"):
        with open(os.path.join(REF, "data", f"{name}.json")) as f:
            js = json.load(f)
        fx[f"centers_{name}"] = np.asarray(js["centers"], dtype=np.float64)
        fx[f"nodes_{name}"] = np.asarray(js["nodes"], dtype=np.int64)
        fx[f"pose_{name}"] = np.asarray(js["pose"], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "fixtures.npz"), **fx)

    # --- fixture pairs under both configurations ---------------------------------------------------
    pairs = [("0", "250"), ("0", "3"), ("3", "0"), ("0", "0"), ("250", "0"), ("3", "250")]
    fp = {}
    for K, N in ((10, 100), (20, 64)):
        trainer = ref_shim.reference_trainer(K=K, node_num=N)
        for a, b in pairs:
            # the reference exits on 3 m < d < 20 m (sg_net.py:302-309); none of these pairs is in that band
            f1, f2, target = fixture_features(trainer, ref, a, b)
            r = run_reference(trainer, f1, f2, trace=True)
            for key, value in r.items():
                fp[f"K{K}_N{N}_{a}_{b}_{key}"] = value
            fp[f"K{K}_N{N}_{a}_{b}_target"] = np.float32(target)
            print(f"fixture K={K} N={N} ({a},{b}) score={r['score'][0]:.9g}")
    np.savez_compressed(os.path.join(OUT, "ref_fixture_pairs.npz"), **fp)

    # --- synthetic KITTI-shape batches -------------------------------------------------------------
    cfgs = [("n64_k20", 64, 20, 8, False), ("n100_k10", 100, 10, 4, False), ("n32_k10", 32, 10, 4, False),
            ("n128_k20", 128, 20, 4, False), ("n16_k10", 16, 10, 4, False), ("n64_k20_dense", 64, 20, 4, True)]
    for tag, N, K, B, dense in cfgs:
        trainer = ref_shim.reference_trainer(K=K, node_num=N)
        f1, f2 = synth.make_pair_batch(B, N, K, seed=1234, dense=dense)
        r = run_reference(trainer, f1, f2, trace=True)
        r["K"], r["N"], r["seed"] = np.int64(K), np.int64(N), np.int64(1234)
        np.savez_compressed(os.path.join(OUT, f"ref_synth_{tag}.npz"), **r)
        print(f"synth {tag}: scores {r['score'][:4]}")

    # --- the other checkpoints on one batch ----------------------------------------------------------
    f1, f2 = synth.make_pair_batch(8, 64, 20, seed=1234)
    ck = {"features_1": f1.numpy(), "features_2": f2.numpy()}
    for tag, path in extra.items():
        trainer = ref_shim.reference_trainer(K=20, node_num=64, model_path=path)
        r = run_reference(trainer, f1, f2, trace=False)
        ck[f"score_{tag}"], ck[f"att_1_{tag}"] = r["score"], r["att_1"]
        print(f"ckpt {tag}: {r['score'][:4]}")
    np.savez_compressed(os.path.join(OUT, "ref_ckpt_scores.npz"), **ck)


if __name__ == "__main__":
    main()
