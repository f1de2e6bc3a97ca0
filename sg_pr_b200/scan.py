"""All-pairs sequence scan (BASELINE config 4): score every ordered pair of an M-graph sequence, sharded row-block
across the GPUs of one box.

The reference never forms the M x M matrix online — its only all-pairs loop is the offline
data_process/gen_sem_kitti_graph_pairs.py:43-52, and eval_batch.py re-embeds both graphs of every listed pair
(eval_batch.py:30-36 → sg_net.py:503-525).  Every graph's pooled vector is independent of its partner,
so the scan is: embed each graph ONCE (fused kernel, csrc/embed_kernel.cuh), then run only the pair head
(layers_batch.py:70-83 + sg_net.py:131-136) over the row block (csrc/head_kernels.cuh).

Sharding over R ranks (one process per GPU, `torch.distributed`):
    rank r embeds graphs [lo_r, hi_r)  ->  all-gather of the pooled vectors (M x 32 fp32, 512 KB at M = 4000)
    rank r scores rows  [lo_r, hi_r) against all M columns  ->  ONE all-gather of the [rows, M] score blocks
Both collectives are NCCL all-gathers over NVLink (gloo in the CPU tests); there is no reduction and no all-to-all.
The compute callables are injected so that the row-block / gather logic is testable on CPU with world_size 2.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def row_block(m: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank`: blocks differ by at most one row, the first m % world get the extra."""
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _all_gather_rows(local: torch.Tensor, m: int, world: int, group=None) -> torch.Tensor:
    """All-gather row blocks of unequal height (at most one row apart) into [m, ...]: pad to the tallest block."""
    if world == 1:
        return local
    tallest = -(-m // world)
    pad = torch.zeros((tallest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * tallest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        lo, hi = row_block(m, r, world)
        pieces.append(out[r * tallest: r * tallest + (hi - lo)])
    return torch.cat(pieces, dim=0)


def scan_all_pairs(graphs: torch.Tensor, k: int,
                   embed_fn: Callable[[torch.Tensor, int], torch.Tensor],
                   score_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                   rank: int = 0, world: int = 1, group=None, gather_scores: bool = True,
                   graphs_are_local: bool = False):
    """Score matrix S[i, j] = score(graph i as side 1, graph j as side 2) for a sequence of M graphs.

    graphs : [M, 15, N] (every rank holds the sequence) or, with graphs_are_local, this rank's row block only.
    embed_fn(graph_block, k) -> pooled [rows, 32];  score_fn(pooled_rows, pooled_all) -> [rows, M].
    Returns (scores, (lo, hi)): the full [M, M] matrix on every rank when gather_scores, else this rank's rows.
    """
    if graphs_are_local:
        counts = torch.tensor([graphs.shape[0]], dtype=torch.int64, device=graphs.device)
        if world > 1:
            dist.all_reduce(counts, group=group)
        m = int(counts.item())
        lo, hi = row_block(m, rank, world)
        if graphs.shape[0] != hi - lo:
            raise ValueError(f"rank {rank} holds {graphs.shape[0]} graphs, row block [{lo},{hi}) expects {hi - lo}")
        mine = graphs
    else:
        m = graphs.shape[0]
        lo, hi = row_block(m, rank, world)
        mine = graphs[lo:hi]
    pooled_rows = embed_fn(mine.contiguous(), k)
    pooled_all = _all_gather_rows(pooled_rows, m, world, group)
    block = score_fn(pooled_rows, pooled_all)
    if not gather_scores:
        return block, (lo, hi)
    return _all_gather_rows(block, m, world, group), (lo, hi)


class SequenceScanner:
    """`scan_all_pairs` bound to one device's Engine (the B200 path)."""

    def __init__(self, engine, rank: int = 0, world: int = 1, group=None):
        self.engine, self.rank, self.world, self.group = engine, rank, world, group

    def _embed(self, block: torch.Tensor, k: int) -> torch.Tensor:
        if block.shape[0] == 0:
            return torch.empty(0, 32, dtype=torch.float32, device=self.engine.device)
        return self.engine.embed(block.to(self.engine.device, non_blocking=True), k)["pooled"]

    def _score(self, rows: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
        return self.engine.score_matrix(rows, cols)

    def scan(self, graphs: torch.Tensor, k: int, gather_scores: bool = True, graphs_are_local: bool = False):
        return scan_all_pairs(graphs, k, self._embed, self._score, self.rank, self.world, self.group,
                              gather_scores, graphs_are_local)

    def top_matches(self, graphs: torch.Tensor, k: int, per_row: int = 5, exclude_window: int = 50):
        """Loop-closure style query: for each row graph of this rank, the best-scoring earlier frames outside a
        +-exclude_window band (the typical use of the score matrix; not part of the reference)."""
        block, (lo, hi) = self.scan(graphs, k, gather_scores=False)
        m = block.shape[1]
        rows = torch.arange(lo, hi, device=block.device)[:, None]
        cols = torch.arange(m, device=block.device)[None, :]
        masked = block.masked_fill((cols > rows - exclude_window), -1.0)
        vals, idx = masked.topk(min(per_row, m), dim=1)
        return vals, idx, (lo, hi)
