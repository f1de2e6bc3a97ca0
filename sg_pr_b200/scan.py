"""All-pairs sequence scan (BASELINE config 4): score every ordered pair of an M-graph sequence, sharded row-block
across the GPUs of one box.

The reference never forms the M x M matrix online — its only all-pairs loop is the offline
data_process/gen_sem_kitti_graph_pairs.py:43-52, and eval_batch.py re-embeds both graphs of every listed pair
(eval_batch.py:30-36 → sg_net.py:503-525).  Every graph's pooled vector is independent of its partner,
so the scan is: embed each graph ONCE (fused kernel, csrc/embed_kernel.cuh), then run only the pair head
(layers_batch.py:70-83 + sg_net.py:131-136) over the row block (csrc/head_kernels.cuh).

Sharding over R ranks (one process per GPU, `torch.distributed`):
    rank r embeds graphs [lo_r, hi_r)  ->  all-gather of the pooled vectors (M x 32 fp32, 512 KB at M = 4000)
    rank r scores rows  [lo_r, hi_r) against all M columns, writing its block INTO the final [M, M] buffer
                                     ->  ONE in-place all-gather of the score rows (64 MB at M = 4000)
Both collectives are NCCL all-gathers over NVLink (gloo in the CPU tests); there is no reduction and no all-to-all.
(The pooled all-gather could be traded for every rank embedding all M graphs; sharding the embed is the faster choice
from 2 GPUs up — bench.py's `scan` key reports each phase.)
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


def _backend(group=None) -> str:
    return str(dist.get_backend(group)) if dist.is_initialized() else "none"


def _all_gather_rows(local: torch.Tensor, m: int, world: int, group=None, out: Optional[torch.Tensor] = None,
                     rank: Optional[int] = None) -> torch.Tensor:
    """All-gather row blocks into [m, ...].

    m % world == 0 (equal blocks): ONE collective straight into the final buffer, no staging copy.  When `local` already
    IS this rank's slice of `out` (the score block is computed in place there) the NCCL all-gather runs in place.
    Otherwise (blocks one row apart): blocks are padded to the tallest, gathered, and the padding rows dropped."""
    if world == 1:
        if out is not None and out.data_ptr() != local.data_ptr():
            out.copy_(local)
        return local if out is None else out
    tail = tuple(local.shape[1:])
    if m % world == 0:
        if out is None:
            out = torch.empty((m,) + tail, dtype=local.dtype, device=local.device)
        src = local
        if _backend(group) != "nccl" and rank is not None and src.data_ptr() == out[rank * (m // world):].data_ptr():
            src = local.clone()                     # gloo (CPU tests): no aliasing of input and output
        dist.all_gather_into_tensor(out, src.contiguous(), group=group)
        return out
    tallest = -(-m // world)
    pad = torch.zeros((tallest,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * tallest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    if out is None:
        out = torch.empty((m,) + tail, dtype=local.dtype, device=local.device)
    for r in range(world):
        lo, hi = row_block(m, r, world)
        out[lo:hi] = buf[r * tallest: r * tallest + (hi - lo)]
    return out


def scan_all_pairs(graphs: torch.Tensor, k: int,
                   embed_fn: Callable[[torch.Tensor, int], torch.Tensor],
                   score_fn: Callable[..., torch.Tensor],
                   rank: int = 0, world: int = 1, group=None, gather_scores: bool = True,
                   graphs_are_local: bool = False, marks: Optional[Callable[[str], None]] = None,
                   out: Optional[torch.Tensor] = None):
    """Score matrix S[i, j] = score(graph i as side 1, graph j as side 2) for a sequence of M graphs.

    graphs : [M, 15, N] (every rank holds the sequence) or, with graphs_are_local, this rank's row block only.
    embed_fn(graph_block, k) -> pooled [rows, 32];  score_fn(pooled_rows, pooled_all, out) -> [rows, M] written into the
    view `out` of the final matrix (out may be None: allocate).
    `marks(name)`: optional callback at the phase boundaries ("embed", "gather_pooled", "score", "gather_scores") —
    bench.py records CUDA events there.
    `out`: optional [M, M] result buffer to reuse across scans (a fresh 64 MB allocation per scan makes the caching
    allocator wait on the NCCL stream's use of the previous one).
    Returns (scores, (lo, hi)): the full [M, M] matrix on every rank when gather_scores, else this rank's rows.

    Collectives: the pooled vectors (M x 32 fp32, 512 KB at M = 4000) and — the one exchange of real size — the score
    rows, gathered by ONE all-gather directly into the [M, M] result (in place: every rank computes its block inside it).
    """
    mark = marks or (lambda name: None)
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
    mark("embed")
    pooled_all = _all_gather_rows(pooled_rows, m, world, group)
    mark("gather_pooled")
    if not gather_scores:
        block = score_fn(pooled_rows, pooled_all, None)
        mark("score")
        return block, (lo, hi)
    if out is not None and (tuple(out.shape) != (m, m) or out.dtype != pooled_all.dtype or out.device != pooled_all.device
                            or not out.is_contiguous()):
        raise ValueError(f"`out` must be a contiguous [{m}, {m}] {pooled_all.dtype} tensor on {pooled_all.device}")
    full = out if out is not None else torch.empty((m, m), dtype=pooled_all.dtype, device=pooled_all.device)
    block = score_fn(pooled_rows, pooled_all, full[lo:hi])
    if block.data_ptr() != full[lo:hi].data_ptr():      # a score_fn that ignores `out`
        full[lo:hi] = block
    mark("score")
    _all_gather_rows(full[lo:hi], m, world, group, out=full, rank=rank)
    mark("gather_scores")
    return full, (lo, hi)


class _DevicePointer:
    """Wraps a raw device allocation for torch.as_tensor (zero-copy, no ownership)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerResult:
    """The [M, M] score matrix of every rank of one box, each reachable from THIS GPU's kernels: the local one is a
    cudaMalloc'ed buffer (wrapped as a tensor), the peers' are CUDA-IPC mappings of their buffers opened with this GPU
    current (handles exchanged once through the process group) — stores to them travel over NVLink / NVSwitch.  With it the
    scan's exchange step is not a collective: the score kernel stores every value into all `world` matrices
    (sgpr_score_matrix_multi), the stores overlap the tensor-core tiles, and a 1-element all-reduce afterwards orders
    them against the readers."""

    def __init__(self, m: int, engine, rank: int, world: int, group=None):
        self.m, self.rank, self.world, self.engine = m, rank, world, engine
        self._ptr, handle = engine.peer_alloc(m * m * 4)
        self.local = torch.as_tensor(_DevicePointer(self._ptr, (m, m)), device=engine.device)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.ptrs = [self._ptr if r == rank else engine.peer_open(handles[r]) for r in range(world)]
        self.flag = torch.zeros(1, dtype=torch.float32, device=engine.device)

    def close(self):
        """Unmap the peers' buffers and free the local one — collectively: every rank must have stopped scanning."""
        if getattr(self, "ptrs", None):
            for r, p in enumerate(self.ptrs):
                if r != self.rank:
                    self.engine.peer_close(p)
            self.ptrs = None
            torch.cuda.synchronize(self.engine.device)
            dist.barrier()
            self.local = None
            self.engine.peer_free(self._ptr)


class SequenceScanner:
    """`scan_all_pairs` bound to one device's Engine (the B200 path)."""

    def __init__(self, engine, rank: int = 0, world: int = 1, group=None):
        self.engine, self.rank, self.world, self.group = engine, rank, world, group
        self._peers = None

    def _embed(self, block: torch.Tensor, k: int) -> torch.Tensor:
        if block.shape[0] == 0:
            return torch.empty(0, 32, dtype=torch.float32, device=self.engine.device)
        return self.engine.embed(block.to(self.engine.device, non_blocking=True), k)["pooled"]

    def _score(self, rows: torch.Tensor, cols: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if rows.shape[0] == 0:
            return out if out is not None else torch.empty(0, cols.shape[0], dtype=torch.float32, device=self.engine.device)
        return self.engine.score_matrix(rows, cols, out=out)

    def scan(self, graphs: torch.Tensor, k: int, gather_scores: bool = True, graphs_are_local: bool = False, marks=None,
             out: Optional[torch.Tensor] = None, exchange: str = "peer"):
        """exchange="peer" (default with more than one rank on a CUDA box): the exchange fused into the score kernel's
        stores over NVLink peer memory (scan_peer_store; the result lives in a scanner-owned buffer, `out` is ignored).
        exchange="nccl": the score kernel, then one in-place NCCL all-gather of the score rows (scan_all_pairs)."""
        if (exchange == "peer" and self.world > 1 and gather_scores and not graphs_are_local
                and self.engine.device.type == "cuda"):
            return self.scan_peer_store(graphs, k, marks)
        return scan_all_pairs(graphs, k, self._embed, self._score, self.rank, self.world, self.group,
                              gather_scores, graphs_are_local, marks, out)

    def scan_peer_store(self, graphs: torch.Tensor, k: int, marks=None):
        """The multi-GPU scan with compute and exchange in ONE kernel: rank r embeds its row block, the pooled vectors are
        all-gathered (512 KB; this collective also tells every rank that all peers have entered this scan, i.e. are done
        reading the previous result), then the tcgen05 score kernel writes each score of the row block into every rank's
        [M, M] matrix — its own HBM and, over NVLink, the peers' (PeerResult) — while the next tiles are on the tensor
        cores; a stream-ordered 1-element all-reduce then publishes "my stores are complete" to the peers.
        Returns (this rank's full matrix — owned by the scanner, overwritten by the next scan — , (lo, hi))."""
        mark = marks or (lambda name: None)
        m = graphs.shape[0]
        lo, hi = row_block(m, self.rank, self.world)
        if self._peers is None or self._peers.m != m:
            if self._peers is not None:
                self._peers.close()
            self._peers = PeerResult(m, self.engine, self.rank, self.world, self.group)
        pr = self._peers
        pooled_rows = self._embed(graphs[lo:hi].contiguous(), k)
        mark("embed")
        pooled_all = _all_gather_rows(pooled_rows, m, self.world, self.group)
        mark("gather_pooled")
        if hi > lo:
            self.engine.score_matrix_multi(pooled_rows, pooled_all, [p + lo * m * 4 for p in pr.ptrs], m)
        mark("score")
        dist.all_reduce(pr.flag, group=self.group)          # stream-ordered: completes once every rank's kernel has finished
        mark("gather_scores")
        return pr.local, (lo, hi)

    def close(self):
        if self._peers is not None:
            self._peers.close()
            self._peers = None

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
