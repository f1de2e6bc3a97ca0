"""Host input pipeline for evaluation: parse each semantic-graph JSON ONCE and keep its padded
`[15, node_num]` fp32 block, instead of re-reading and re-padding both files for every listed pair as the reference
does (utils.process_pair, /root/reference/utils.py:21-38 + SGTrainer.transfer_to_torch, sg_net.py:241-310 — measured
0.5 ms per pair, i.e. ~2 k pairs/s per thread).

The block a graph gets here is bit-identical to what `transfer_to_torch(..., training=False)` builds for it
(tests/test_host_logic.py checks it against the reference's own output).  Graphs with MORE than node_num nodes are
randomly subsampled by the reference on every use (np.random.choice, sg_net.py:252-256); those are not cached as
blocks — `block()` re-samples them per call exactly like the reference.
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch


class GraphStore:
    def __init__(self, node_num: int, number_of_labels: int = 12):
        self.node_num = int(node_num)
        self.labels = int(number_of_labels)
        self._raw: Dict[str, tuple] = {}            # path -> (nodes int64[n], centers float64[n,3], (x, z))
        self._blocks: Dict[str, torch.Tensor] = {}  # path -> [15, node_num] float32 (only graphs with n <= node_num)
        self._compact: Dict[str, torch.Tensor] = {}  # path -> uint8 [ceil16(13 node_num)] compact record (same graphs)

    def __len__(self):
        return len(self._raw)

    def _load(self, path: str):
        hit = self._raw.get(path)
        if hit is None:
            with open(path) as handle:
                g = json.load(handle)
            nodes = np.asarray(g["nodes"], dtype=np.int64)
            centers = np.asarray(g["centers"], dtype=np.float64).reshape(len(nodes), 3)
            hit = (nodes, centers, (float(g["pose"][3]), float(g["pose"][11])))
            self._raw[path] = hit
        return hit

    def _build(self, nodes: np.ndarray, centers: np.ndarray) -> torch.Tensor:
        n, want = len(nodes), self.node_num
        if n > want:                                        # sg_net.py:252-256: sorted random subset, fresh every call
            keep = np.random.choice(n, want, replace=False)
            keep.sort()
            nodes, centers, n = nodes[keep], centers[keep], want
        block = np.zeros((3 + self.labels, want), dtype=np.float64)
        block[:3, :n] = centers.T
        block[3 + nodes, np.arange(n)] = 1.0                # one-hot rows 3..14; pads stay all-zero (sg_net.py:258-278)
        return torch.from_numpy(block.astype(np.float32))

    def block(self, path: str) -> torch.Tensor:
        """`[15, node_num]` fp32 block of one graph (cached when the graph needs no subsampling)."""
        hit = self._blocks.get(path)
        if hit is not None:
            return hit
        nodes, centers, _ = self._load(path)
        blk = self._build(nodes, centers)
        if len(nodes) <= self.node_num:
            self._blocks[path] = blk
        return blk

    def compact_record(self, path: str) -> torch.Tensor:
        """The graph as a compact record (include/sgpr_b200.h: xyz [3][node_num] fp32, then one label byte per node, 255 for
        pads, padded to 16 bytes) — the 13-byte-per-node form of `block(path)`; the kernels expand it to the same bits."""
        hit = self._compact.get(path)
        if hit is not None:
            return hit
        blk = self.block(path)
        n = self.node_num
        rec = torch.zeros(((13 * n + 15) // 16) * 16, dtype=torch.uint8)
        rec[:12 * n] = blk[:3].contiguous().view(-1).view(torch.uint8)
        sem = blk[3:]
        rec[12 * n:13 * n] = torch.where(sem.sum(dim=0) > 0, sem.argmax(dim=0), torch.full((n,), 255)).to(torch.uint8)
        if len(self._load(path)[0]) <= self.node_num:
            self._compact[path] = rec
        return rec

    def is_static(self, path: str) -> bool:
        """True when the graph's block never changes between calls (no random subsampling)."""
        return len(self._load(path)[0]) <= self.node_num

    def distance(self, path_a: str, path_b: str) -> float:
        """Planar pose distance used as ground truth (utils.py:34-36: pose[3], pose[11])."""
        (xa, za), (xb, zb) = self._load(path_a)[2], self._load(path_b)[2]
        return math.sqrt((xa - xb) ** 2 + (za - zb) ** 2)

    def target(self, path_a: str, path_b: str, p_thresh: float) -> float:
        """sg_net.py:302-309: 1.0 within p_thresh, 0.0 beyond 20 m, otherwise the reference prints and exits."""
        d = self.distance(path_a, path_b)
        if d <= p_thresh:
            return 1.0
        if d >= 20:
            return 0.0
        print("distance error: ", d)
        raise SystemExit(-1)

    def pair_batch(self, pairs: Sequence[Sequence[str]], p_thresh: float, pin: bool = False
                   ) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
        """features_1, features_2 `[B, 15, node_num]` + targets for a list of [path_a, path_b]."""
        b = len(pairs)
        f1 = torch.empty(b, 3 + self.labels, self.node_num, dtype=torch.float32, pin_memory=pin)
        f2 = torch.empty_like(f1, pin_memory=pin)
        targets = np.empty(b, dtype=np.float64)
        for i, (pa, pb) in enumerate(pairs):
            f1[i] = self.block(pa)
            f2[i] = self.block(pb)
            targets[i] = self.target(pa, pb, p_thresh)
        return f1, f2, targets
