"""CPU tests of the host side: config parsing, graph-pair loading, transfer_to_torch against what the REFERENCE's
own host prep produced for the fixtures (tests/golden/ref_fixture_pairs.npz), checkpoint round trip, C-ABI symbols."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import write_fixture_tree


@pytest.fixture(scope="module")
def tree(golden_dir, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("sgpr_tree"))
    return root, write_fixture_tree(golden_dir, root)


def test_abi_exports_every_declared_symbol():
    """The library loads without a GPU and exports every function include/*.h declares."""
    from sg_pr_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "sgpr_b200.h")).read() + \
        open(os.path.join(root, "include", "sgpr_b200_train.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # prose in comments is not a declaration
    declared = set(re.findall(r"\b(sgpr_[a-z0-9_]+)\s*\(", header))
    declared -= {"sgpr_ctx", "sgpr_train"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sgpr_abi_version() == 1
    assert lib.sgpr_packed_size() > 47000


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back to anything."""
    from sg_pr_b200 import _lib
    lib = _lib.load()
    handle = ctypes.c_void_p()
    assert lib.sgpr_create(ctypes.byref(handle), 0) != 0
    assert b"no CUDA device" in lib.sgpr_last_error() or b"CUDA" in lib.sgpr_last_error()
    from sg_pr_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)
    assert lib.sgpr_train_create(ctypes.byref(handle), 0) != 0           # the training step has no CPU path either
    from sg_pr_b200.train_engine import TrainEngine
    with pytest.raises(RuntimeError):
        TrainEngine(0)


def test_parser_reads_reference_layout(tree):
    from sg_pr_b200.parser_sg import sgpr_args
    root, cfg = tree
    args = sgpr_args()
    assert args.K == 10 and args.node_num == 100 and args.batch_size == 128     # parser_sg.py:4-34 defaults
    args.load(cfg)
    assert args.cuda == "0" and args.gpu == 0 and args.K == 10 and args.filters_3 == 32
    assert args.pair_file == [f"{root}/data/0.json", f"{root}/data/250.json"]
    assert args.train_sequences == ["00"] and args.show is False


def test_process_pair_and_transfer_to_torch_match_reference(tree, golden_dir):
    """Features built by our host prep == features the reference's process_pair + transfer_to_torch built."""
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SGTrainer
    from sg_pr_b200.utils import process_pair
    root, cfg = tree
    z = np.load(os.path.join(golden_dir, "ref_fixture_pairs.npz"))
    for K, N in ((10, 100), (20, 64)):
        args = sgpr_args().load(cfg)
        args.K, args.node_num = K, N
        trainer = SGTrainer(args, False)
        for a, b in (("0", "250"), ("0", "3"), ("3", "0"), ("0", "0")):
            pair = process_pair([f"{root}/data/{a}.json", f"{root}/data/{b}.json"])
            data = trainer.transfer_to_torch(pair, False)
            p = f"K{K}_N{N}_{a}_{b}_"
            assert data["features_1"].shape == (15, N)
            np.testing.assert_array_equal(torch.FloatTensor(np.array([data["features_1"]])).numpy(), z[p + "features_1"])
            np.testing.assert_array_equal(torch.FloatTensor(np.array([data["features_2"]])).numpy(), z[p + "features_2"])
            assert data["target"] == float(z[p + "target"])
    pair = process_pair([f"{root}/data/0.json", f"{root}/data/250.json"])
    assert abs(pair["distance"] - 133.13) < 0.01


def test_graph_store_matches_reference_host_prep(tree, golden_dir):
    """GraphStore blocks / targets == what the reference's process_pair + transfer_to_torch produced (golden)."""
    from sg_pr_b200.graph_store import GraphStore
    root, _ = tree
    z = np.load(os.path.join(golden_dir, "ref_fixture_pairs.npz"))
    for K, N in ((10, 100), (20, 64)):
        store = GraphStore(N, 12)
        for a, b in (("0", "250"), ("0", "3"), ("3", "0"), ("0", "0"), ("250", "0"), ("3", "250")):
            pa, pb = f"{root}/data/{a}.json", f"{root}/data/{b}.json"
            f1, f2, gt = store.pair_batch([[pa, pb]], 3)
            p = f"K{K}_N{N}_{a}_{b}_"
            np.testing.assert_array_equal(f1.numpy(), z[p + "features_1"])
            np.testing.assert_array_equal(f2.numpy(), z[p + "features_2"])
            assert gt[0] == float(z[p + "target"]) and store.is_static(pa)
            from sg_pr_b200.engine import compact_graphs
            rec = store.compact_record(pa)              # the 13-byte-per-node record == compact form of the reference's block
            assert torch.equal(rec, compact_graphs(f1)[0]) and rec.numel() == ((13 * N + 15) // 16) * 16
        assert len(store) == 3                          # three files parsed once each, however many pairs
    small = GraphStore(16, 12)
    np.random.seed(0)
    b1 = small.block(f"{root}/data/0.json")
    assert b1.shape == (15, 16) and not small.is_static(f"{root}/data/0.json") and float(b1[3:].sum()) == 16.0


def test_subsampling_when_graph_exceeds_node_num(tree):
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SGTrainer
    from sg_pr_b200.utils import process_pair
    root, cfg = tree
    args = sgpr_args().load(cfg)
    args.node_num = 16
    trainer = SGTrainer(args, False)
    np.random.seed(0)
    data = trainer.transfer_to_torch(process_pair([f"{root}/data/0.json", f"{root}/data/3.json"]), False)
    f = data["features_1"]
    assert f.shape == (15, 16) and (f[3:].sum(axis=0) == 1).all()      # 16 real nodes, one label each


def test_checkpoint_contract(tree, kitti_state):
    """strict load of the reference checkpoint; state_dict round trip keeps the DataParallel key format."""
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SGTrainer
    root, cfg = tree
    trainer = SGTrainer(sgpr_args().load(cfg), False)
    sd = trainer.model.state_dict()
    assert all(k.startswith("module.") for k in sd) and len(sd) == 50
    for name, value in kitti_state.items():
        assert torch.equal(sd["module." + name].cpu(), value), name
    assert hasattr(trainer.model, "module") and not trainer.model.training or True
    trainer.model.eval()
    assert not trainer.model.module.training


def test_training_path_matches_oracle_in_eval_math(kitti_state):
    """The stock-PyTorch baseline form (benchmarks/tests only) computes the same function as the oracle (CPU, eval BN)."""
    from sg_pr_b200.torch_baseline import forward_torch
    from oracle import sgpr_oracle as orc
    from sg_pr_b200 import synth
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SG
    args = sgpr_args()
    args.K, args.node_num = 20, 64
    model = SG(args, 12)
    model.load_state_dict(kitti_state)
    model.eval()
    f1, f2 = synth.make_pair_batch(4, 64, 20, seed=8)
    with torch.no_grad():
        score, a1, a2 = forward_torch(model, f1, f2)
    want = orc.forward_pairs(f1, f2, 20, kitti_state)
    assert float((score - want["score"]).abs().max()) <= 1e-5
    assert float((a1 - want["att_1"]).abs().max()) <= 1e-5


def test_augmentations_shapes_and_ranges():
    from sg_pr_b200 import utils
    np.random.seed(1)
    pts = np.random.rand(1, 50, 3) * 10
    rot = utils.rotate_point_cloud(pts.copy())
    np.testing.assert_allclose(np.linalg.norm(rot, axis=2), np.linalg.norm(pts, axis=2), rtol=1e-5)
    np.testing.assert_allclose(rot[..., 2], pts[..., 2], rtol=1e-6)          # rotation about z
    jit = utils.jitter_point_cloud(pts.copy())
    assert np.abs(jit - pts).max() <= 0.05 + 1e-12
    sc = utils.random_scale_point_cloud(pts.copy())
    ratio = sc / pts
    assert 0.8 <= ratio.min() and ratio.max() <= 1.25 and np.ptp(ratio) < 1e-9
    sh = utils.shift_point_cloud(pts.copy())
    assert np.abs(sh - pts).max() <= 0.3 + 1e-12
    per = utils.rotate_perturbation_point_cloud(pts.copy())
    np.testing.assert_allclose(np.linalg.norm(per, axis=2), np.linalg.norm(pts, axis=2), rtol=1e-5)


def test_load_paires_and_listdir(tmp_path):
    from sg_pr_b200.utils import listDir, load_paires
    lst = tmp_path / "00.txt"
    lst.write_text("1.json 2.json\n3.json 4.json\n")
    pairs = load_paires(str(lst), "/graphs")
    assert pairs == [["/graphs/1.json", "/graphs/2.json"], ["/graphs/3.json", "/graphs/4.json"]]
    (tmp_path / "sub").mkdir()
    (tmp_path / "sub" / "a.json").write_text("{}")
    found = []
    listDir(str(tmp_path), found)
    assert sorted(os.path.basename(p) for p in found) == ["00.txt", "a.json"]


def test_fast_training_features_are_bit_identical_to_the_reference_shaped_path(tree):
    """process_batch's training host path (_training_features: parsed-once arrays + one fused augmentation function) builds
    the same float64 blocks, with the same consumption of numpy's and random's global streams, as
    transfer_to_torch(process_pair(pair), training=True) — padding, subsampling (node_num below the graph size) and flip."""
    import random
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SGTrainer
    from sg_pr_b200.utils import process_pair
    root, cfg = tree
    pairs = [[f"{root}/data/{a}.json", f"{root}/data/{b}.json"] for a, b in (("0", "3"), ("0", "250"), ("3", "250"), ("0", "0"))]
    for node_num in (64, 100, 16):
        args = sgpr_args().load(cfg)
        args.node_num = node_num
        trainer = SGTrainer(args, False)
        for seed in range(6):
            np.random.seed(seed); random.seed(seed)
            want = [trainer.transfer_to_torch(process_pair(p), True) for p in pairs]
            tail_w = (np.random.rand(), random.random())
            np.random.seed(seed); random.seed(seed)
            got = [trainer._training_features(p) for p in pairs]
            tail_g = (np.random.rand(), random.random())
            assert tail_w == tail_g                                  # both RNG streams advanced identically
            for w, (a, b, t) in zip(want, got):
                assert a.dtype == w["features_1"].dtype == np.float64 and a.shape == (15, node_num)
                assert np.array_equal(a, w["features_1"]) and np.array_equal(b, w["features_2"]) and t == w["target"]
            # the batched form process_batch uses: draws in the same order, arithmetic over the whole batch at once
            np.random.seed(seed); random.seed(seed)
            feats, targets = trainer._training_batch(pairs)
            assert (np.random.rand(), random.random()) == tail_w
            assert feats.dtype == np.float32 and feats.shape == (2 * len(pairs), 15, node_num) and targets.dtype == np.float32
            for i, w in enumerate(want):
                for side, key in enumerate(("features_1", "features_2")):
                    ref = w[key].astype(np.float32)
                    assert np.array_equal(feats[2 * i + side], ref)
                    assert np.array_equal(np.signbit(feats[2 * i + side]), np.signbit(ref))     # flipped pads are -0.0
                assert targets[2 * i] == targets[2 * i + 1] == w["target"]


def test_compact_from_blocks_matches_the_python_packer():
    """sgpr_compact_from_blocks (host-only C entry point: one-hot [15][N] blocks -> 13-byte-per-node records) against
    engine.compact_graphs, and its refusal of blocks that have no compact form.  No device needed."""
    import ctypes as C
    import torch
    from sg_pr_b200 import _lib, synth
    from sg_pr_b200.engine import compact_graphs
    lib = _lib.bind(C.CDLL(_lib.LIB_PATH), _lib.SYMBOLS)
    for n, k in ((64, 20), (33, 8), (128, 20), (5, 2)):
        g = synth.make_graphs(40, n, k, seed=n)
        g[3, 0, :] = -g[3, 0, :]                                   # flipped cloud: -0.0 coordinates on the pads stay as they are
        out = torch.full((40, int(lib.sgpr_compact_stride(n))), 7, dtype=torch.uint8)
        assert lib.sgpr_compact_from_blocks(g.data_ptr(), 40, n, out.data_ptr()) == 0
        assert torch.equal(out, compact_graphs(g))
    out = torch.empty(4, 832, dtype=torch.uint8)
    for edit in (lambda t: t[2, 5].__setitem__(3, 0.5), lambda t: t[1, 4].__setitem__(60, -0.0),
                 lambda t: (t[1, 4].__setitem__(0, 1.0), t[1, 5].__setitem__(0, 1.0))):
        g = synth.make_graphs(4, 64, 20, seed=1)
        edit(g)
        assert lib.sgpr_compact_from_blocks(g.data_ptr(), 4, 64, out.data_ptr()) == 1
    assert lib.sgpr_compact_from_blocks(None, 4, 64, out.data_ptr()) < 0 and b"NULL" in lib.sgpr_last_error()
