"""`python -m sg_pr_b200.eval_batch [config.yml]` — the reference's eval_batch.py (/root/reference/eval_batch.py:14-91) with
the evaluation metrics computed on the device.

Same inputs (config sections `common`, `eva_batch`; `<pair_list_dir>/<seq>.txt`), same outputs in `output_path`:
    <seq>_gt_db.npy   float64 ground truth        <seq>_DL_db.npy   float32 predictions       (eval_batch.py:43-48)
    <seq>_DL_F1_max.txt   str(max F1 over the precision-recall curve)                           (eval_batch.py:84-91)
byte-identical to what the reference script writes for the same predictions (tests/test_reference_scripts.py runs the
unmodified script and this module side by side).  The ROC / PR curves are computed with sg_pr_b200.metrics (one sort +
cumsum on the GPU instead of sklearn on Python lists: after an all-pairs scan that is 10^7 scores); the PNG plots
(eval_batch.py:57-82) are drawn only when matplotlib is installed.
"""
import os
import sys

import numpy as np
import torch

from . import metrics
from .parser_sg import sgpr_args
from .sg_net import SGTrainer
from .utils import load_paires, tab_printer

try:
    from tqdm import tqdm
except ImportError:  # pragma: no cover
    def tqdm(it, **kw):
        return it


def evaluate_sequence(trainer, args, sequence: str) -> dict:
    """One sequence of eval_batch.py's loop (:26-91): predictions for every listed pair, files, metrics."""
    graph_pairs = load_paires(os.path.join(args.pair_list_dir, sequence + ".txt"), args.graph_pairs_dir)
    batches = [graph_pairs[i:i + args.batch_size] for i in range(0, len(graph_pairs), args.batch_size)]
    pred_db, gt_db = [], []
    for batch in tqdm(batches):
        pred, gt = trainer.eval_batch_pair(batch)
        pred_db.extend(pred)
        gt_db.extend(gt)
    assert len(pred_db) == len(gt_db)
    assert np.sum(gt_db) > 0  # gt_db should have positive samples
    pred_db, gt_db = np.array(pred_db), np.array(gt_db)
    np.save(os.path.join(args.output_path, sequence + "_gt_db.npy"), gt_db)
    np.save(os.path.join(args.output_path, sequence + "_DL_db.npy"), pred_db)
    dev = torch.device("cuda", int(args.gpu)) if torch.cuda.is_available() else torch.device("cpu")
    y = torch.as_tensor(gt_db.astype(np.float64), device=dev)
    s = torch.as_tensor(pred_db.astype(np.float64), device=dev)
    fpr, tpr, _ = metrics.roc_curve(y, s)
    roc_auc = metrics.auc(fpr, tpr)
    precision, recall, _ = metrics.pr_curve(y, s)
    f1 = torch.nan_to_num(2 * precision * recall / (precision + recall), nan=0.0)
    f1_max = np.float64(f1.max().item())
    print("roc_auc: ", roc_auc)
    print("F1 max score", f1_max)
    with open(os.path.join(args.output_path, sequence + "_DL_F1_max.txt"), "w") as out:
        out.write(str(f1_max))
    try:
        from matplotlib import pyplot as plt
    except ImportError:
        plt = None
    if plt is not None:  # pragma: no cover - matplotlib is not in the build image
        for fig, (xs, ys, xl, yl, title, name) in enumerate((
                (fpr, tpr, "False Positive Rate", "True Positive Rate", "DL ROC Curve", "_DL_roc_curve.png"),
                (recall, precision, "Recall", "Precision", "DL Precision-Recall Curve", "_DL_pr_curve.png"))):
            plt.figure(fig)
            plt.plot(xs.cpu().numpy(), ys.cpu().numpy(), color="darkorange", lw=2)
            plt.xlabel(xl)
            plt.ylabel(yl)
            plt.title(title)
            plt.savefig(os.path.join(args.output_path, sequence + name))
        if args.show:
            plt.show()
    return {"pairs": len(pred_db), "roc_auc": roc_auc, "f1_max": float(f1_max)}


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    args = sgpr_args()
    args.load(argv[0] if argv else "./config/config.yml")
    tab_printer(args)
    trainer = SGTrainer(args, False)
    trainer.model.eval()
    os.makedirs(args.output_path, exist_ok=True)
    results = {}
    for sequence in tqdm(args.sequences):
        print("sequence: ", sequence)
        results[sequence] = evaluate_sequence(trainer, args, sequence)
    return results


if __name__ == "__main__":
    main()
