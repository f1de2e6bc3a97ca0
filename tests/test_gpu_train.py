"""Training step on the B200 through the C-ABI (include/sgpr_b200_train.h) against the reference's golden training
vectors: the same assertions as the emulated CPU twin, plus shapes the goldens do not cover checked against the
oracle's own train step run live."""
import numpy as np
import pytest
import torch

from tests import train_checks as tc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from sg_pr_b200.train_engine import TrainEngine
    e = TrainEngine(0)
    yield e
    e.close()


@pytest.mark.parametrize("mirrored", [False, True])
@pytest.mark.parametrize("tag,tol", [("n32_k10", 1e-5), ("n64_k20", 5e-5)])
def test_gradients_match_reference(eng, tag, tol, mirrored):
    g, sd = tc.load_case(tag)
    tc.check_gradients(eng, g, sd, "cuda", pred_tol=tol, mirrored=mirrored)


@pytest.mark.parametrize("mirrored", [False, True])
@pytest.mark.parametrize("tag,tol", [("n32_k10", 1e-5), ("n64_k20", 5e-5)])
def test_two_optimiser_steps_match_reference(eng, tag, tol, mirrored):
    g, sd = tc.load_case(tag)
    tc.check_two_steps(eng, g, sd, "cuda", pred_tol=tol, mirrored=mirrored)


def test_mirrored_equals_two_sided(eng, kitti_state):
    """Embedding every graph once (mirrored) gives the two-sided step's predictions and gradients up to rounding."""
    from oracle.make_golden_train import train_batch
    f1, f2, target = train_batch(24, 64, 20, seed=11)
    out = []
    for mirrored in (False, True):
        eng.set_state(kitti_state, reset_optimizer=True)
        _, pred = eng.step(f1.cuda(), None if mirrored else f2.cuda(), target.cuda(), 20, apply=False, mirrored=mirrored)
        out.append((pred.cpu().clone(), eng.grads()))
    assert float((out[0][0] - out[1][0]).abs().max()) <= 2e-6
    for n, a in out[0][1].items():
        assert float((a - out[1][1][n]).abs().max()) <= 2e-5 * max(float(a.abs().max()), 1e-8), n
    with pytest.raises(RuntimeError, match="odd"):
        eng.step(f1[:3].cuda(), None, target[:3].cuda(), 20, mirrored=True)


@pytest.mark.parametrize("N,k,listed", [(100, 10, 6), (30, 7, 5), (128, 20, 3), (64, 40, 3)])
def test_step_matches_live_oracle(eng, kitti_state, N, k, listed):
    """Shapes without golden vectors (the shipped config N=100/k=10, ragged N, the largest N): one full step vs the
    oracle's autograd + Adam on the CPU."""
    from oracle import sgpr_oracle_train as ort
    from oracle.make_golden_train import train_batch
    f1, f2, target = train_batch(listed, N, k, seed=5)
    sd = {n: v.clone() for n, v in kitti_state.items()}
    adam = ort.new_adam_state(sd)
    want = ort.train_step(sd, f1, f2, target, k, adam, 1e-3, 5e-4)
    eng.set_state(kitti_state, reset_optimizer=True)
    eng.set_optimizer(1e-3, 5e-4)
    loss, pred = eng.step(f1.cuda(), f2.cuda(), target.cuda(), k, apply=True)
    np.testing.assert_allclose(pred.cpu().numpy(), want["pred"].numpy(), rtol=0, atol=5e-5)
    assert abs(float(loss) - want["loss"]) < 5e-5
    grads = eng.grads()
    for name, ref in want["grads"].items():
        scale = max(float(ref.abs().max()), 1e-8)
        err = float((grads[name].reshape(ref.shape) - ref).abs().max())
        assert err <= 2e-3 * scale, (name, err, scale)
    state = eng.get_state()
    for name, value in state.items():
        if "running_" in name:
            ref = sd[name]
            assert float((value.reshape(ref.shape) - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), name


def test_step_is_deterministic_in_everything_but_fp64_atomics(eng, kitti_state):
    """Gradients come from fixed-order partial sums; only the fp64 statistics use atomics (order noise ~1e-16)."""
    from oracle.make_golden_train import train_batch
    f1, f2, target = train_batch(16, 64, 20, seed=9)
    outs = []
    for _ in range(2):
        eng.set_state(kitti_state, reset_optimizer=True)
        loss, pred = eng.step(f1.cuda(), f2.cuda(), target.cuda(), 20, apply=False)
        outs.append((pred.cpu().clone(), {n: v.clone() for n, v in eng.grads().items()}))
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 1e-6
    for n in outs[0][1]:
        a, b = outs[0][1][n], outs[1][1][n]
        assert float((a - b).abs().max()) <= 1e-5 * max(float(a.abs().max()), 1e-8), n


def test_loud_errors(eng, kitti_state):
    from sg_pr_b200 import _lib
    from sg_pr_b200.train_engine import TrainEngine
    fresh = TrainEngine(0)
    f = torch.zeros(2, 15, 32, device="cuda")
    t = torch.zeros(2, device="cuda")
    with pytest.raises(_lib.SgprError, match="set_state"):
        fresh.step(f, f, t, 10)
    fresh.set_state(kitti_state)
    with pytest.raises(_lib.SgprError, match="topk"):
        fresh.step(f, f, t, 40)
    with pytest.raises(RuntimeError):
        fresh.step(f.cpu(), f.cpu(), t.cpu(), 10)
    fresh.close()


def test_batch_assembly_matches_host_augmentation(eng, monkeypatch):
    """sgpr_train_assemble vs the reference-shaped host path fed with the kernel's own random draws."""
    from tests import assemble_checks as ac
    ac.check_assemble(eng, "cuda", monkeypatch, M=12, N=64, P=9)
    ac.check_assemble(eng, "cuda", monkeypatch, M=6, N=100, P=4, seed=99, step=2 ** 33 + 5)


def test_batch_assembly_draw_statistics(eng):
    from sg_pr_b200 import synth
    from tests import assemble_checks as ac
    a, b = synth.make_pair_batch(8, 64, 20, seed=2)
    graphs = torch.cat([a, b]).cuda()
    pair_idx = torch.randint(0, 16, (4096, 2), dtype=torch.int32).cuda()
    _, draws, jitter = eng.assemble(graphs, pair_idx, seed=5, step=0, want_draws=True)
    ac.check_draw_statistics(draws.cpu(), jitter.cpu())


def test_split_forward_backward_matches_reference(eng):
    g, sd = tc.load_case("n64_k20")
    tc.check_split_forward_backward(eng, g, sd, "cuda", pred_tol=5e-5)
    g, sd = tc.load_case("n32_k10")
    tc.check_split_forward_backward(eng, g, sd, "cuda", pred_tol=1e-5, mirrored=True)


def test_module_forward_in_train_mode_is_an_autograd_node(kitti_state):
    """model.train(); model(data); loss.backward(); torch.optim.Adam.step() — the reference's own training idiom
    (sg_net.py:332-338) on the drop-in module, against the golden gradients and the state after one step."""
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SG
    g, sd = tc.load_case("n32_k10")
    args = sgpr_args()
    args.K, args.node_num, args.gpu = int(g["K"]), int(g["N"]), 0
    model = SG(args, 12)
    model.load_state_dict(sd)
    model.cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    data = {"features_1": torch.from_numpy(g["features_1"]), "features_2": torch.from_numpy(g["features_2"]),
            "target": torch.from_numpy(g["target"])}
    opt.zero_grad()
    pred, a1, a2 = model(data)
    assert pred.requires_grad and not a1.requires_grad and a1.shape == (8, 32, 1)
    loss = torch.mean(torch.nn.functional.binary_cross_entropy(pred, data["target"].cuda()))
    loss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["pred1"], rtol=0, atol=1e-5)
    assert abs(float(loss) - float(g["loss1"])) < 1e-5
    for name, p in model.named_parameters():
        ref = g["grad1." + name]
        scale = max(float(np.abs(ref).max()), 1e-8)
        assert float(np.abs(p.grad.cpu().numpy() - ref).max()) <= 1e-3 * scale, name
    opt.step()
    state = model.state_dict()
    assert int(state["dgcnn_s_conv1.1.num_batches_tracked"]) == int(g["state1.dgcnn_s_conv1.1.num_batches_tracked"])
    for name in ("dgcnn_s_conv2.1.running_mean", "dgcnn_conv_end.1.running_var"):
        ref = g["state1." + name]
        assert float(np.abs(state[name].cpu().numpy() - ref).max()) <= 1e-4 * max(1.0, float(np.abs(ref).max())), name
    # eval mode afterwards: the fused eval kernel sees the stepped weights (re-packed on version change)
    model.eval()
    with torch.no_grad():
        s, _, _ = model(data)
    assert s.shape == (8,) and bool(torch.isfinite(s).all())


@pytest.mark.parametrize("B,N,k", [(5, 32, 10), (7, 64, 20)])
def test_general_two_sided_batch(eng, kitti_state, B, N, k):
    tc.check_general_two_sided_batch(eng, kitti_state, "cuda", B=B, N=N, k=k)
