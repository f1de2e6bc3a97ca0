"""The reference's own entry points, UNMODIFIED, on top of the drop-in modules (VERDICT r01 task 8).

/root/reference/eval_pair.py and eval_batch.py are executed as they are (runpy) with sg_pr_b200/dropin first on the module
path, a stub matplotlib (not in the image) and the kernels' CPU-emulator build standing in for the GPU
(tests/stubs/emu_hook.py — a test-only hook; on the B200 box the same flow is covered through the CUDA library by
tests/test_gpu_dropin.py, where /root/reference does not exist).  Expected numbers are the reference's own outputs
(tests/golden/ref_fixture_pairs.npz).  Also: sg_pr_b200.eval_batch (device metrics) writes
byte-identical .npy / F1_max files.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import write_fixture_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "eval_batch.py")), reason="reference tree not on this box")

PAIR_LINES = "0.json 3.json\n0.json 250.json\n3.json 250.json\n0.json 0.json\n"


@pytest.fixture(scope="module")
def tree(tmp_path_factory, golden_dir):
    root = str(tmp_path_factory.mktemp("ref_tree"))
    cfg = write_fixture_tree(golden_dir, root)
    os.makedirs(os.path.join(root, "lists"), exist_ok=True)
    with open(os.path.join(root, "lists", "00.txt"), "w") as f:
        f.write(PAIR_LINES)
    return root, cfg


def _run(code, cwd):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "sg_pr_b200", "dropin"),
                                                       os.path.join(ROOT, "tests", "stubs"), ROOT]),
               PYTHONWARNINGS="ignore")
    return subprocess.run([sys.executable, "-c", code], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)


def test_unmodified_eval_pair_prints_the_reference_score(tree):
    root, _ = tree
    out = _run(f"import tests.stubs.emu_hook as h; h.run('{REF}/eval_pair.py')", root)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("Score:")]
    assert line, out.stdout[-2000:]
    score = float(line[0].split()[1])
    assert abs(score - 1.34899221e-06) <= 1e-9          # K=10, N=100, (0.json, 250.json); 1e-5 bar, met to 1e-9
    assert "1.34" in line[0] and "e-06" in line[0]


def test_unmodified_eval_batch_and_device_metrics_driver_write_identical_files(tree, golden_dir):
    root, cfg = tree
    ref_out = os.path.join(root, "eva")
    out = _run(f"import tests.stubs.emu_hook as h; h.run('{REF}/eval_batch.py')", root)
    assert out.returncode == 0, out.stderr[-3000:]
    names = ("00_gt_db.npy", "00_DL_db.npy", "00_DL_F1_max.txt")
    for name in names + ("00_DL_roc_curve.png", "00_DL_pr_curve.png"):
        assert os.path.isfile(os.path.join(ref_out, name)), name
    pred = np.load(os.path.join(ref_out, "00_DL_db.npy"))
    gt = np.load(os.path.join(ref_out, "00_gt_db.npy"))
    assert pred.dtype == np.float32 and gt.dtype == np.float64
    assert gt.tolist() == [1.0, 0.0, 0.0, 1.0]
    with np.load(os.path.join(golden_dir, "ref_fixture_pairs.npz")) as z:
        want = [float(z[f"K10_N100_{a}_{b}_score"][0]) for a, b in (("0", "3"), ("0", "250"), ("3", "250"), ("0", "0"))]
    assert np.abs(pred - np.array(want)).max() <= 1e-5
    # ---- our driver (device metrics) into a second directory: byte-identical files ----
    mine = os.path.join(root, "eva_b200")
    code = (f"import tests.stubs.emu_hook as h; h.install(); import yaml\n"
            f"from sg_pr_b200 import eval_batch as eb\n"
            f"cfg = yaml.safe_load(open('{cfg}')); cfg['eva_batch']['output_path'] = '{mine}'\n"
            f"open('{root}/cfg_b200.yml', 'w').write(yaml.safe_dump(cfg))\n"
            f"print(eb.main(['{root}/cfg_b200.yml']))")
    out2 = _run(code, root)
    assert out2.returncode == 0, out2.stderr[-3000:]
    for name in names:
        with open(os.path.join(ref_out, name), "rb") as a, open(os.path.join(mine, name), "rb") as b:
            assert a.read() == b.read(), f"{name} differs from what the reference's eval_batch.py wrote"
