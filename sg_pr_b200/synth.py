"""Synthetic semantic-graph generators (KITTI-shape), shared by tests, bench and the golden-vector script.

The distribution follows SURVEY.md §8(d): what `SGTrainer.transfer_to_torch`
(/root/reference/sg_net.py:241-310) emits for SemanticKITTI graphs — a `[15, N]` channel-major fp32
block per graph (rows 0-2 = node centre xyz in metres, rows 3-14 = one-hot of 12 semantic labels),
with the trailing `N - n_real` nodes all-zero exactly like the reference's padding
(sg_net.py:258-262, 276-278).

Everything is generated on the CPU from a seeded `torch.Generator`, so the same seed gives the same
bytes here, on the GPU box and inside the golden-vector script.
"""
from __future__ import annotations

import torch

NUM_LABELS = 12
NUM_CHANNELS = 3 + NUM_LABELS
_EXTENT = torch.tensor([100.0, 100.0, 4.0])


def real_node_range(node_num: int, k: int) -> tuple[int, int]:
    """Range of real-node counts that keeps `#pads >= k` (tie-rule independent k-NN)."""
    hi = min(node_num - k, max(1, int(round(node_num * 44 / 64))))
    lo = min(hi, max(1, int(round(node_num * 25 / 64))))
    if hi < 1:
        raise ValueError(f"node_num={node_num} cannot hold k={k} pads plus a real node")
    return lo, hi


def make_graphs(num_graphs: int, node_num: int = 64, k: int = 20, seed: int = 0,
                dense: bool = False) -> torch.Tensor:
    """`[num_graphs, 15, node_num]` fp32 graphs.

    dense=False: KITTI shape, `n_real ~ U{lo..hi}` with at least k zero pads.
    dense=True : every node real (tie-dominated k-NN on the one-hot branch; reported separately).
    """
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(num_graphs, NUM_CHANNELS, node_num, dtype=torch.float32)
    if dense:
        n_real = torch.full((num_graphs,), node_num, dtype=torch.int64)
    else:
        lo, hi = real_node_range(node_num, k)
        n_real = torch.randint(lo, hi + 1, (num_graphs,), generator=g)
    xyz = (torch.rand(num_graphs, node_num, 3, generator=g) - 0.5) * _EXTENT
    lab = torch.randint(0, NUM_LABELS, (num_graphs, node_num), generator=g)
    live = (torch.arange(node_num)[None, :] < n_real[:, None])
    out[:, :3, :] = (xyz * live[:, :, None] + 0.0).permute(0, 2, 1)      # + 0.0: pads are +0.0 like np.zeros, never -0.0
    onehot = torch.nn.functional.one_hot(lab, NUM_LABELS).to(torch.float32) * live[:, :, None]
    out[:, 3:, :] = onehot.permute(0, 2, 1)
    return out.contiguous()


def make_pair_batch(batch: int, node_num: int = 64, k: int = 20, seed: int = 0,
                    dense: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """Two independent `[batch, 15, node_num]` sides (features_1, features_2)."""
    gs = make_graphs(2 * batch, node_num, k, seed, dense)
    return gs[:batch].contiguous(), gs[batch:].contiguous()


def make_sequence_pairs(num_graphs: int, num_pairs: int, seed: int = 0) -> torch.Tensor:
    """`[num_pairs, 2]` int64 ordered index pairs drawn from a sequence of `num_graphs` graphs."""
    g = torch.Generator().manual_seed(seed + 7919)
    return torch.randint(0, num_graphs, (num_pairs, 2), generator=g)


def make_train_batch(listed: int, node_num: int = 64, k: int = 20, seed: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """What `SGTrainer.process_batch` feeds the model for `listed` pairs (sg_net.py:324-331): every pair in both orders,
    so features_1 is `[2*listed, 15, N]` with rows (a0, b0, a1, b1, ...) and features_2[p] == features_1[p ^ 1].
    Returns (features_1, target [2*listed]) — half the listed pairs positive (BASELINE config 3: p_thresh 3 m)."""
    a, b = make_pair_batch(listed, node_num, k, seed=seed)
    f1 = torch.stack([a, b], dim=1).reshape(2 * listed, NUM_CHANNELS, node_num).contiguous()
    target = torch.tensor([float(i % 2 == 0) for i in range(listed)]).repeat_interleave(2)
    return f1, target
