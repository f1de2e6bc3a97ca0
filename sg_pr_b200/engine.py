"""Python face of the C-ABI (include/sgpr_b200.h): one `Engine` per device.

PyTorch is plumbing here — device memory, the current stream, `.data_ptr()` — the arithmetic of the hot path
(/root/reference/sg_net.py:112-138) happens in csrc/*.cuh.  No method of this class has a CPU or eager-PyTorch
fallback: if the library is missing or no CUDA device is present the constructor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import SgprBn, SgprWeights, c_float_p, check

F3 = 32
IN_CHANNELS = 15
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # current stream of a device as a plain integer
MAX_NODES = 128

_EDGE_LAYERS = (("dgcnn_s_conv1", "dgcnn_s_conv2", "dgcnn_s_conv3"), ("dgcnn_f_conv1", "dgcnn_f_conv2", "dgcnn_f_conv3"))


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def weights_struct(state: Dict[str, "torch.Tensor"], bn_eps: float = 1e-5) -> Tuple[SgprWeights, list]:
    """Build the sgpr_weights struct from a reference state_dict (keys without 'module.', sg_net.py:168-174).
    Returns (struct, keepalive list of the numpy arrays the struct points into)."""
    keep = []

    def ptr(name, shape=None):
        arr = _f32(state[name])
        if shape is not None and tuple(arr.shape) != tuple(shape):
            arr2 = arr.reshape(shape) if arr.size == int(np.prod(shape)) else None
            if arr2 is None:
                raise ValueError(f"{name}: expected shape {shape}, got {arr.shape}")
            arr = np.ascontiguousarray(arr2)
        keep.append(arr)
        return arr.ctypes.data_as(c_float_p)

    def bn(prefix, c):
        return SgprBn(ptr(prefix + ".weight", (c,)), ptr(prefix + ".bias", (c,)),
                      ptr(prefix + ".running_mean", (c,)), ptr(prefix + ".running_var", (c,)))

    f1 = int(state["dgcnn_s_conv1.0.weight"].shape[0])
    f2 = int(state["dgcnn_s_conv2.0.weight"].shape[0])
    f3 = int(state["dgcnn_s_conv3.0.weight"].shape[0])
    t = int(state["tensor_network.bias"].shape[0])
    bnk = int(state["fully_connected_first.weight"].shape[0])
    w = SgprWeights()
    w.filters[0], w.filters[1], w.filters[2] = f1, f2, f3
    w.tensor_neurons, w.bottleneck, w.bn_eps = t, bnk, float(bn_eps)
    if (f1, f2, f3, t, bnk) == (64, 64, 32, 16, 16):
        cin = ((3, 64, 64), (12, 64, 64))
        cout = (64, 64, 32)
        for b, (dst_w, dst_bn) in enumerate(((w.s_conv_w, w.s_bn), (w.f_conv_w, w.f_bn))):
            for l in range(3):
                dst_w[l] = ptr(_EDGE_LAYERS[b][l] + ".0.weight", (cout[l], 2 * cin[b][l]))
                dst_bn[l] = bn(_EDGE_LAYERS[b][l] + ".1", cout[l])
        w.end_conv_w = ptr("dgcnn_conv_end.0.weight", (32, 64))
        w.end_bn = bn("dgcnn_conv_end.1", 32)
        w.att_w = ptr("attention.weight_matrix", (32, 32))
        w.ntn_w = ptr("tensor_network.weight_matrix", (32, 32 * 16))
        w.ntn_v = ptr("tensor_network.weight_matrix_block", (16, 64))
        w.ntn_b = ptr("tensor_network.bias", (16,))
        w.fc1_w = ptr("fully_connected_first.weight", (16, 16))
        w.fc1_b = ptr("fully_connected_first.bias", (16,))
        w.fc2_w = ptr("scoring_layer.weight", (16,))
        w.fc2_b = ptr("scoring_layer.bias", (1,))
    # any other architecture: leave pointers NULL; the library answers SGPR_E_ARCH with a clear message
    return w, keep


def pack_weights_host(state: Dict[str, "torch.Tensor"]):
    """Host-only packing (no GPU): returns (blob float32[packed], head float32[289], offsets dict)."""
    lib = _lib.load()
    w, keep = weights_struct(state)
    n = lib.sgpr_packed_size()
    blob = np.zeros(n, dtype=np.float32)
    head = np.zeros(289, dtype=np.float32)
    offs = (C.c_size_t * 17)()
    check(lib.sgpr_pack_weights_host(C.byref(w), blob.ctypes.data_as(c_float_p), head.ctypes.data_as(c_float_p), offs),
          "sgpr_pack_weights_host")
    names = ["s1", "w_s2", "w_s3", "w_f1", "w_f2", "w_f3", "w_end", "ab_s2", "ab_s3", "ab_f1", "ab_f2", "ab_f3",
             "ab_end", "att_w", "ntn_w", "ntn_v", "ntn_b"]
    del keep
    return blob, head, dict(zip(names, [int(x) for x in offs]))


def compact_graphs(blocks: torch.Tensor, stride: Optional[int] = None) -> torch.Tensor:
    """[B, 15, N] one-hot blocks (CPU) -> uint8 [B, ceil16(13 N)] compact records: xyz rows as they are, then one label
    byte per node (index of the one-hot 1; 255 where the node has none, e.g. zero pads).  Raises if a node's label rows
    are not one-hot / all-zero — such a block has no compact form."""
    b, c, n = blocks.shape
    if c != IN_CHANNELS or blocks.dtype != torch.float32:
        raise ValueError(f"expected float32 [B, {IN_CHANNELS}, N], got {blocks.dtype} {tuple(blocks.shape)}")
    sem = blocks[:, 3:, :]
    ones = (sem == 1.0).sum(dim=1)
    if not bool((((ones == 1) & ((sem != 0).sum(dim=1) == 1)) | ((sem != 0).sum(dim=1) == 0)).all()):
        raise ValueError("label rows are not one-hot: this batch has no compact form")
    label = torch.where(ones == 1, sem.argmax(dim=1), torch.full_like(ones, 255)).to(torch.uint8)
    stride = stride or ((13 * n + 15) // 16) * 16
    out = torch.zeros(b, stride, dtype=torch.uint8)
    out[:, :12 * n] = blocks[:, :3, :].contiguous().view(b, 3 * n).view(torch.uint8).view(b, 12 * n)
    out[:, 12 * n:13 * n] = label
    return out


def _check_graphs(t: torch.Tensor, what: str) -> Tuple[int, int]:
    if t.dim() != 3 or t.shape[1] != IN_CHANNELS:
        raise ValueError(f"{what}: expected [B, {IN_CHANNELS}, N], got {tuple(t.shape)}")
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")
    return int(t.shape[0]), int(t.shape[2])


class Engine:
    """Owns one sgpr_ctx (device-bound packed weights + scratch)."""

    def __init__(self, device: int | torch.device | str = 0, lib=None):
        """`lib`: tests/emu only — the emulator build of the same sources, with host memory standing in for the device.
        The product never passes it: without a CUDA device the constructor raises."""
        self._emulated = lib is not None
        if self._emulated:
            self.device = torch.device("cpu")
            self._lib = lib
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("sg_pr_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
            dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
            if dev.type != "cuda":
                raise RuntimeError(f"sg_pr_b200 engine cannot run on {dev}")
            self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
            self._device_index = int(self.device.index)
            self._lib = _lib.load()
        handle = C.c_void_p()
        check(self._lib.sgpr_create(C.byref(handle), 0 if self._emulated else self.device.index), "sgpr_create", self._lib)
        self._ctx = handle
        self.has_weights = False

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.sgpr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ------------------------------------------------------------------------------------------
    def set_weights(self, state: Dict[str, torch.Tensor], bn_eps: float = 1e-5):
        w, keep = weights_struct(state, bn_eps)
        check(self._lib.sgpr_set_weights(self._ctx, C.byref(w)), "sgpr_set_weights", self._lib)
        del keep
        self.has_weights = True

    def launch_count(self) -> int:
        return int(self._lib.sgpr_launch_count(self._ctx))

    # ---- k-NN tie rule (dgcnn.py:19) ------------------------------------------------------------------------
    KNN_TIES = {"cuda": 0, "cpu": 1}

    def set_knn_ties(self, mode: str):
        """"cuda" (default): ties at the k-th distance go to the lowest node index — ATen's CUDA topk, the reference on
        its native device.  "cpu": the order ATen's CPU topk (std::nth_element) leaves — the reference run on a CPU.
        The two differ only for graphs with fewer than k zero pads (include/sgpr_b200.h, sgpr_set_knn_ties)."""
        if mode not in self.KNN_TIES:
            raise ValueError(f"knn_ties must be one of {sorted(self.KNN_TIES)}, got {mode!r}")
        check(self._lib.sgpr_set_knn_ties(self._ctx, self.KNN_TIES[mode]), "sgpr_set_knn_ties", self._lib)

    def knn_ties(self) -> str:
        return {v: k for k, v in self.KNN_TIES.items()}[int(self._lib.sgpr_get_knn_ties(self._ctx))]

    def _stream(self) -> C.c_void_p:
        if self._emulated:
            return None
        if _RAW_STREAM is not None:              # same stream, without building a torch.cuda.Stream object (~2 us per call)
            return C.c_void_p(_RAW_STREAM(self._device_index))
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t: torch.Tensor, what: str, allow_pinned: bool = False) -> torch.Tensor:
        """A tensor the kernels can address: on the engine's device, or (inputs only) pinned host memory, which UVA
        maps into the device address space — the fused kernel then pulls each graph block over PCIe by itself."""
        if t.device != self.device and not (allow_pinned and t.device.type == "cpu" and t.is_pinned()):
            raise ValueError(f"{what} lives on {t.device}, engine on {self.device}")
        return t.contiguous()

    # ---- the hot path ---------------------------------------------------------------------------------------
    def forward_pairs(self, f1: torch.Tensor, f2: torch.Tensor, k: int, want_att: bool = True, _checked: bool = False):
        """Device (or pinned-host) tensors in, device tensors out: (score [B], att_1 [B,N,1], att_2 [B,N,1]) — sg_net.py:138.
        `_checked`: the caller (SG.forward) has already validated placement / dtype / layout."""
        if not _checked:
            _check_graphs(f1, "features_1")
            if tuple(f2.shape) != tuple(f1.shape):
                raise ValueError(f"features_2 {tuple(f2.shape)} != features_1 {tuple(f1.shape)}")
            f1, f2 = self._dev(f1, "features_1", True), self._dev(f2, "features_2", True)
        b, n = int(f1.shape[0]), int(f1.shape[2])
        score = torch.empty(b, dtype=torch.float32, device=self.device)
        if want_att:
            att1 = torch.empty((b, n, 1), dtype=torch.float32, device=self.device)
            att2 = torch.empty((b, n, 1), dtype=torch.float32, device=self.device)
            p1, p2 = att1.data_ptr(), att2.data_ptr()
        else:
            att1 = att2 = p1 = p2 = None
        rc = self._lib.sgpr_forward_pairs(self._ctx, f1.data_ptr(), f2.data_ptr(), b, n, int(k), score.data_ptr(), p1, p2,
                                          self._stream())
        if rc:
            check(rc, "sgpr_forward_pairs", self._lib)
        return score, att1, att2

    def forward_pairs_host(self, f1: torch.Tensor, f2: torch.Tensor, k: int, want_att: bool = True,
                           out: Optional[Tuple[torch.Tensor, ...]] = None):
        """CPU tensors in, CPU tensors out (H2D + kernel + D2H + sync inside the C call)."""
        b, n = _check_graphs(f1, "features_1")
        if f1.is_cuda or f2.is_cuda:
            raise ValueError("forward_pairs_host takes CPU tensors")
        f1, f2 = f1.contiguous(), f2.contiguous()
        if out is None:
            score = torch.empty(b, dtype=torch.float32)
            att1 = torch.empty(b, n, 1, dtype=torch.float32) if want_att else None
            att2 = torch.empty(b, n, 1, dtype=torch.float32) if want_att else None
        else:
            score, att1, att2 = out
        check(self._lib.sgpr_forward_pairs_host(self._ctx, f1.data_ptr(), f2.data_ptr(), b, n, int(k), score.data_ptr(),
                                                att1.data_ptr() if att1 is not None else None,
                                                att2.data_ptr() if att2 is not None else None),
              "sgpr_forward_pairs_host", self._lib)
        return score, att1, att2

    # ---- compact input (13 bytes per node instead of 60) -------------------------------------------------------------
    def compact_stride(self, n: int) -> int:
        return int(self._lib.sgpr_compact_stride(int(n)))

    def _check_compact(self, g: torch.Tensor, n: int, what: str) -> torch.Tensor:
        if g.dtype != torch.uint8 or g.dim() != 2 or g.shape[1] != self.compact_stride(n):
            raise ValueError(f"{what}: expected uint8 [B, {self.compact_stride(n)}] compact records (see compact_graphs), "
                             f"got {g.dtype} {tuple(g.shape)}")
        return self._dev(g, what, True)

    def forward_pairs_compact(self, g1: torch.Tensor, g2: torch.Tensor, n: int, k: int, want_att: bool = True):
        """Compact records in (device or pinned host), device tensors out — bit-identical to forward_pairs on the expanded
        [B,15,N] blocks."""
        g1, g2 = self._check_compact(g1, n, "graphs_1"), self._check_compact(g2, n, "graphs_2")
        if g1.shape != g2.shape:
            raise ValueError("graphs_1 / graphs_2 differ in shape")
        b = int(g1.shape[0])
        score = torch.empty(b, dtype=torch.float32, device=self.device)
        att1 = torch.empty((b, n, 1), dtype=torch.float32, device=self.device) if want_att else None
        att2 = torch.empty((b, n, 1), dtype=torch.float32, device=self.device) if want_att else None
        check(self._lib.sgpr_forward_pairs_compact(self._ctx, g1.data_ptr(), g2.data_ptr(), b, int(n), int(k), score.data_ptr(),
                                                   att1.data_ptr() if want_att else None,
                                                   att2.data_ptr() if want_att else None, self._stream()),
              "sgpr_forward_pairs_compact", self._lib)
        return score, att1, att2

    def compact_from_blocks(self, blocks: torch.Tensor, out: torch.Tensor) -> bool:
        """Host-side: contiguous float32 CPU blocks [B, 15, N] -> compact records written into `out` (uint8 CPU tensor of at
        least B * stride bytes, typically pinned).  False when the batch has no compact form (label rows not one-hot)."""
        b, n = int(blocks.shape[0]), int(blocks.shape[2])
        rc = self._lib.sgpr_compact_from_blocks(blocks.data_ptr(), b, n, out.data_ptr())
        if rc < 0:
            check(rc, "sgpr_compact_from_blocks", self._lib)
        return rc == 0

    def embed_compact(self, graphs: torch.Tensor, n: int, k: int, want_att: bool = False) -> dict:
        graphs = self._check_compact(graphs, n, "graphs")
        m = int(graphs.shape[0])
        out = {"pooled": torch.empty(m, F3, dtype=torch.float32, device=self.device)}
        if want_att:
            out["att"] = torch.empty(m, n, 1, dtype=torch.float32, device=self.device)
        check(self._lib.sgpr_embed_compact(self._ctx, graphs.data_ptr(), m, int(n), int(k), out["pooled"].data_ptr(),
                                           out["att"].data_ptr() if want_att else None, None, self._stream()),
              "sgpr_embed_compact", self._lib)
        return out

    # ---- embed-once / score-many ------------------------------------------------------------------------------
    def embed(self, graphs: torch.Tensor, k: int, want_att: bool = False, want_emb: bool = False, trace: bool = False):
        """[M,15,N] -> dict(pooled [M,32], att [M,N,1]?, emb [M,N,32]?, knn [M,6,N,k] uint8?, layers [M,6,N,64]?)."""
        m, n = _check_graphs(graphs, "graphs")
        graphs = self._dev(graphs, "graphs", True)
        out = {"pooled": torch.empty(m, F3, dtype=torch.float32, device=self.device)}
        if want_att:
            out["att"] = torch.empty(m, n, 1, dtype=torch.float32, device=self.device)
        if want_emb:
            out["emb"] = torch.empty(m, n, F3, dtype=torch.float32, device=self.device)
        if trace:
            out["knn"] = torch.zeros(m, 6, n, int(k), dtype=torch.uint8, device=self.device)
            out["layers"] = torch.zeros(m, 6, n, 64, dtype=torch.float32, device=self.device)
        p = lambda key: out[key].data_ptr() if key in out else None
        check(self._lib.sgpr_embed_trace(self._ctx, graphs.data_ptr(), m, n, int(k), p("pooled"), p("att"), p("emb"),
                                         p("knn"), p("layers"), self._stream()), "sgpr_embed", self._lib)
        return out

    def score_pairs(self, pooled: torch.Tensor, pair_idx: torch.Tensor) -> torch.Tensor:
        pooled = self._dev(pooled, "pooled")
        idx = self._dev(pair_idx, "pair_idx").to(torch.int32).contiguous()
        p = int(idx.shape[0])
        score = torch.empty(p, dtype=torch.float32, device=self.device)
        check(self._lib.sgpr_score_pairs(self._ctx, pooled.data_ptr(), idx.data_ptr(), p, score.data_ptr(), self._stream()),
              "sgpr_score_pairs", self._lib)
        return score

    def score_matrix(self, pooled_rows: torch.Tensor, pooled_cols: torch.Tensor, out: Optional[torch.Tensor] = None):
        rows, cols = self._dev(pooled_rows, "pooled_rows"), self._dev(pooled_cols, "pooled_cols")
        r, m = int(rows.shape[0]), int(cols.shape[0])
        if out is None:
            out = torch.empty(r, m, dtype=torch.float32, device=self.device)
        if out.shape != (r, m) or out.stride(1) != 1 or out.device != self.device or out.dtype != torch.float32:
            raise ValueError("score_matrix: `out` must be a float32 [R, M] device tensor with unit column stride")
        check(self._lib.sgpr_score_matrix(self._ctx, rows.data_ptr(), r, cols.data_ptr(), m, out.data_ptr(),
                                          int(out.stride(0)) if r > 0 else m, self._stream()), "sgpr_score_matrix", self._lib)
        return out

    def score_matrix_multi(self, pooled_rows: torch.Tensor, pooled_cols: torch.Tensor, out_ptrs, ld: int):
        """score_matrix with every score stored at the same offsets behind EACH device address of `out_ptrs` (row 0 of this
        row block inside this GPU's result and inside the peers' results mapped by peer_open — see scan.PeerResult): the
        multi-GPU scan's exchange fused into the kernel's stores."""
        rows, cols = self._dev(pooled_rows, "pooled_rows"), self._dev(pooled_cols, "pooled_cols")
        r, m = int(rows.shape[0]), int(cols.shape[0])
        ptrs = (C.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
        check(self._lib.sgpr_score_matrix_multi(self._ctx, rows.data_ptr(), r, cols.data_ptr(), m, ptrs, len(out_ptrs), int(ld),
                                                self._stream()), "sgpr_score_matrix_multi", self._lib)

    # ---- peer-visible buffers (CUDA IPC) ---------------------------------------------------------------------------------
    def peer_alloc(self, nbytes: int):
        """(device pointer, 64-byte IPC handle) of a fresh allocation on this engine's GPU."""
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        check(self._lib.sgpr_peer_alloc(self._ctx, int(nbytes), C.byref(ptr), handle), "sgpr_peer_alloc", self._lib)
        return int(ptr.value), handle.raw

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        check(self._lib.sgpr_peer_open(self._ctx, C.create_string_buffer(handle, 64), C.byref(ptr)), "sgpr_peer_open", self._lib)
        return int(ptr.value)

    def peer_close(self, ptr: int):
        self._lib.sgpr_peer_close(self._ctx, C.c_void_p(ptr))

    def peer_free(self, ptr: int):
        self._lib.sgpr_peer_free(self._ctx, C.c_void_p(ptr))
