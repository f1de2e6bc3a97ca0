"""Checks shared by the CPU (emulated kernels) and GPU tests of the training step: the same assertions the oracle
itself has to pass against the reference's golden training vectors (tests/test_oracle_golden.py)."""
import os

import numpy as np
import torch

from oracle import sgpr_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_case(tag):
    g = np.load(os.path.join(GOLDEN, f"ref_train_{tag}.npz"))
    sd = orc.load_state_npz(os.path.join(GOLDEN, "model_kitti.npz"))
    return g, sd


def check_gradients(eng, g, sd, device, pred_tol, mirrored=False):
    """forward + backward without an update: loss, predictions and every parameter gradient vs the reference."""
    f1 = torch.from_numpy(g["features_1"]).to(device)
    f2 = torch.from_numpy(g["features_2"]).to(device)
    target = torch.from_numpy(g["target"]).to(device)
    eng.set_state(sd)
    before = {k: v.clone() for k, v in eng.get_state().items()}
    # the golden batches have process_batch's structure (features_2[p] == features_1[p ^ 1]): valid input for both modes
    loss, pred = eng.step(f1, None if mirrored else f2, target, int(g["K"]), apply=False, mirrored=mirrored)
    # train mode has no 1e-5 bar (tests/test_train_form.py): a near-tie k-NN flip moves every prediction through the
    # batch statistics; n64_k20 holds one such flip
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred1"], rtol=0, atol=pred_tol)
    assert abs(float(loss) - float(g["loss1"])) < pred_tol
    grads = eng.grads()
    for name, got in grads.items():
        ref = g["grad1." + name]
        scale = max(float(np.abs(ref).max()), 1e-8)
        err = float(np.abs(got.numpy().reshape(ref.shape) - ref).max())
        assert err <= 1e-3 * scale, (name, err, scale)
    after = eng.get_state()
    for name in before:                                   # apply=0 must leave parameters and statistics alone
        assert torch.equal(before[name], after[name]), name
    assert eng.step_count() == 0


def check_two_steps(eng, g, sd, device, pred_tol, mirrored=False):
    """two optimiser steps: predictions, losses, parameters and running statistics vs the reference's own run."""
    f1 = torch.from_numpy(g["features_1"]).to(device)
    f2 = torch.from_numpy(g["features_2"]).to(device)
    target = torch.from_numpy(g["target"]).to(device)
    lr = float(g["lr"])
    eng.set_state(sd, reset_optimizer=True)
    eng.set_optimizer(lr, float(g["weight_decay"]))
    for step in (1, 2):
        loss, pred = eng.step(f1, None if mirrored else f2, target, int(g["K"]), apply=True, mirrored=mirrored)
        np.testing.assert_allclose(pred.cpu().numpy(), g[f"pred{step}"], rtol=0, atol=pred_tol * step)
        assert abs(float(loss) - float(g[f"loss{step}"])) < pred_tol * step
        state = eng.get_state()
        for name, value in state.items():
            ref = g[f"state{step}." + name]
            diff = np.abs(value.numpy().reshape(ref.shape) - ref)
            if "running_" in name:
                assert diff.max() <= 1e-4 * max(1.0, float(np.abs(ref).max())), (name, step, float(diff.max()))
                continue
            # Adam's first steps move every weight by ~lr whatever the gradient's size (update = lr * g / (|g| + eps)), so
            # where the gradient is below this implementation's rounding noise the DIRECTION is noise: at step 1 (the
            # golden file holds that gradient) every element whose decayed gradient clears the noise floor must agree
            # tightly; everything else, and step 2, is bounded by the largest moves Adam can make (+lr vs -lr per step).
            tight = diff <= 2e-5 + 1e-5 * np.abs(ref)
            assert diff.max() <= 2 * lr * step * 1.01, (name, step, float(diff.max()))    # opposite signs: 2 lr apart
            if step == 1:
                gref = g["grad1." + name] + float(g["weight_decay"]) * sd[name].numpy().reshape(ref.shape)
                clear = np.abs(gref) > 2e-3 * max(float(np.abs(g["grad1." + name]).max()), 1e-8)
                assert tight[clear].all(), (name, float(diff[clear].max()))
            else:
                assert tight.mean() >= 0.85, (name, step, float(tight.mean()))
    assert eng.step_count() == 2


def check_split_forward_backward(eng, g, sd, device, pred_tol, mirrored=False):
    """sgpr_train_forward + sgpr_train_backward (the step split at the loss, autograd driven by the caller): feeding
    d mean-BCE / d prediction computed by torch must reproduce the reference's gradients, and the forward alone must
    update the running statistics the way the reference's first forward did."""
    f1 = torch.from_numpy(g["features_1"]).to(device)
    f2 = torch.from_numpy(g["features_2"]).to(device)
    target = torch.from_numpy(g["target"]).to(device)
    eng.set_state(sd)
    pred, att1, att2 = eng.forward(f1, None if mirrored else f2, int(g["K"]), update_running=True, mirrored=mirrored)
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred1"], rtol=0, atol=pred_tol)
    assert att1.shape == (f1.shape[0], f1.shape[2], 1) and att2.shape == att1.shape
    leaf = pred.detach().clone().requires_grad_(True)
    loss = torch.mean(torch.nn.functional.binary_cross_entropy(leaf, target))
    loss.backward()
    flat = eng.backward(leaf.grad)
    assert abs(loss.item() - float(g["loss1"])) < pred_tol
    grads = eng.grads()
    off = 0
    for name, got in grads.items():
        ref = g["grad1." + name]
        scale = max(float(np.abs(ref).max()), 1e-8)
        err = float(np.abs(got.numpy().reshape(ref.shape) - ref).max())
        assert err <= 1e-3 * scale, (name, err, scale)
        assert torch.equal(flat[off:off + got.numel()].cpu(), got.reshape(-1))           # the device copy is the same vector
        off += got.numel()
    state = eng.get_state()
    for name, value in state.items():
        ref = g["state1." + name]
        if "running_" in name:                       # updated by the forward, both sides
            assert float(np.abs(value.numpy().reshape(ref.shape) - ref).max()) <= 1e-4 * max(1.0, float(np.abs(ref).max())), name
        else:                                        # parameters untouched
            assert torch.equal(value, sd[name].reshape(value.shape)), name
    flat_state = eng.get_state_flat()
    eng.set_state_flat(flat_state)
    assert torch.equal(eng.get_state_flat(), flat_state)


def check_general_two_sided_batch(eng, sd, device, B=5, N=32, k=10, seed=31):
    """features_2 unrelated to features_1 (odd batch, different graphs per side): each side is its own BatchNorm batch
    with its own statistics.  One applied step vs the oracle's autograd + Adam."""
    from oracle import sgpr_oracle_train as ort
    from sg_pr_b200 import synth
    f1, _ = synth.make_pair_batch(B, N, k, seed=seed)
    f2, _ = synth.make_pair_batch(B, N, k, seed=seed + 1)
    target = (torch.arange(B) % 2).float()
    work = {n: v.clone() for n, v in sd.items()}
    want = ort.train_step(work, f1, f2, target, k, ort.new_adam_state(work), 1e-3, 5e-4)
    eng.set_state(sd, reset_optimizer=True)
    eng.set_optimizer(1e-3, 5e-4)
    loss, pred = eng.step(f1.to(device), f2.to(device), target.to(device), k, apply=True)
    np.testing.assert_allclose(pred.cpu().numpy(), want["pred"].numpy(), rtol=0, atol=5e-5)
    assert abs(float(loss) - want["loss"]) < 5e-5
    grads = eng.grads()
    for name, ref in want["grads"].items():
        scale = max(float(ref.abs().max()), 1e-8)
        assert float((grads[name].reshape(ref.shape) - ref).abs().max()) <= 2e-3 * scale, name
    state = eng.get_state()
    for name, value in state.items():
        if "running_" in name:            # side 1 then side 2, each with its own batch statistics
            ref = work[name]
            assert float((value.reshape(ref.shape) - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), name
