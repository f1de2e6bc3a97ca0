"""CPU checks of the host logic: the C library's weight packing and the kernel's arithmetic form (numpy model in
tests/kernel_model.py) against the oracle.  No GPU, no compute calls into the library."""
import os

import numpy as np
import pytest
import torch

from oracle import sgpr_oracle as orc
from sg_pr_b200 import engine, synth
from tests import kernel_model as km


@pytest.fixture(scope="module")
def packed(kitti_state):
    return engine.pack_weights_host(kitti_state)


def test_pack_layout_and_sign_fold(kitti_state, packed):
    blob, head, offs = packed
    for off in offs.values():
        assert off % 4 == 0          # 16-byte aligned sections (cp.async.bulk source alignment)
    for name, cout in (("ab_s2", 64), ("ab_s3", 32), ("ab_f1", 64), ("ab_f2", 64), ("ab_f3", 32)):
        assert (blob[offs[name]:offs[name] + cout] >= 0).all(), name     # alpha made non-negative
    assert (blob[offs["s1"]:offs["s1"] + 512].reshape(64, 8)[:, 6] >= 0).all()
    # unfolding the sign must give back the reference conv matrix and BN terms
    sd = kitti_state
    inv = 1.0 / torch.sqrt(sd["dgcnn_s_conv2.1.running_var"] + 1e-5)
    alpha = (inv * sd["dgcnn_s_conv2.1.weight"]).numpy()
    sign = np.where(alpha < 0, -1.0, 1.0).astype(np.float32)
    w = km.unpack_pairs(blob[offs["w_s2"]:offs["w_s2"] + 64 * 128], 64, 128)
    ref = sd["dgcnn_s_conv2.0.weight"].numpy().reshape(64, 128)
    np.testing.assert_array_equal(w[:, :64], (ref[:, :64] * sign[:, None]).T)
    np.testing.assert_array_equal(w[:, 64:], (ref[:, 64:] * sign[:, None]).T)
    np.testing.assert_allclose(blob[offs["ab_s2"]:offs["ab_s2"] + 64], np.abs(alpha), rtol=1e-6)
    assert int((alpha < 0).sum()) > 0   # the shipped checkpoint does have negative BN scales
    np.testing.assert_array_equal(head[:256], sd["fully_connected_first.weight"].numpy().reshape(-1))


def test_unsupported_architecture_is_refused(kitti_state):
    bad = dict(kitti_state)
    bad["tensor_network.bias"] = torch.zeros(8, 1)
    with pytest.raises(Exception) as ei:
        engine.pack_weights_host(bad)
    assert "unsupported" in str(ei.value) or "code -4" in str(ei.value)


@pytest.mark.parametrize("n,k,dense", [(64, 20, False), (100, 10, False), (32, 10, False), (64, 20, True)])
def test_kernel_form_matches_oracle(kitti_state, packed, n, k, dense):
    blob, head, offs = packed
    f1, f2 = synth.make_pair_batch(3, n, k, seed=77, dense=dense)
    want = orc.forward_pairs(f1, f2, k, kitti_state, want_trace=True)
    for b in range(f1.shape[0]):
        g1 = km.embed_graph(f1[b].numpy(), k, blob, offs)
        g2 = km.embed_graph(f2[b].numpy(), k, blob, offs)
        if not dense:
            for layer in range(6):
                ok = orc.knn_sets_equivalent(want["knn_pd_1"][layer][b:b + 1], want["knn_idx_1"][layer][b:b + 1],
                                             torch.from_numpy(g1["knn"][layer])[None],
                                             want["layer_in_1"][layer][b:b + 1])
                assert bool(ok.all()), (b, layer)
                np.testing.assert_allclose(g1["layers"][layer], want["layer_out_1"][layer][b].numpy().T, atol=2e-5)
            np.testing.assert_allclose(g1["emb"], want["emb_1"][b].numpy(), atol=2e-5)
            np.testing.assert_allclose(g1["att"], want["att_1"][b, :, 0].numpy(), atol=1e-5)
            s = km.pair_score(g1["pooled"], g2["pooled"], blob, head, offs)
            assert abs(float(s) - float(want["score"][b])) <= 1e-5
        else:
            # dense one-hot graphs are tie-dominated: only the head arithmetic is comparable
            s = km.pair_score(want["pooled_1"][b, :, 0].numpy(), want["pooled_2"][b, :, 0].numpy(), blob, head, offs)
            assert abs(float(s) - float(want["score"][b])) <= 1e-5
