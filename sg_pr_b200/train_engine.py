"""Host side of the device training step (include/sgpr_b200_train.h): flat state <-> state_dict, one call per step.

What it stands in for is the device half of `SGTrainer.process_batch(batch, training=True)` (/root/reference/sg_net.py:
332-338): train-mode forward, mean BCE, backward and the Adam update, all in csrc/train_kernels.cuh.  No method here
has a CPU or eager-PyTorch implementation behind it.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib

# state_dict shapes of the tensors in the flat layout (sg_net.py:44-76, layers_batch.py:16-18, 58-60)
SHAPES = {
    "dgcnn_s_conv1.0.weight": (64, 6, 1, 1), "dgcnn_s_conv2.0.weight": (64, 128, 1, 1),
    "dgcnn_s_conv3.0.weight": (32, 128, 1, 1), "dgcnn_f_conv1.0.weight": (64, 24, 1, 1),
    "dgcnn_f_conv2.0.weight": (64, 128, 1, 1), "dgcnn_f_conv3.0.weight": (32, 128, 1, 1),
    "dgcnn_conv_end.0.weight": (32, 64, 1), "attention.weight_matrix": (32, 32),
    "tensor_network.weight_matrix": (32, 32, 16), "tensor_network.weight_matrix_block": (16, 64),
    "tensor_network.bias": (16, 1), "fully_connected_first.weight": (16, 16), "fully_connected_first.bias": (16,),
    "scoring_layer.weight": (1, 16), "scoring_layer.bias": (1,),
}
BN_LAYERS = ("dgcnn_s_conv1.1", "dgcnn_s_conv2.1", "dgcnn_s_conv3.1", "dgcnn_f_conv1.1", "dgcnn_f_conv2.1",
             "dgcnn_f_conv3.1", "dgcnn_conv_end.1")


def layout(lib=None) -> Tuple[List[Tuple[str, int, int]], int]:
    """[(name, offset, size)] of the flat state vector and how many of the entries are trainable parameters."""
    lib = lib or _lib.load()
    count, n_params = C.c_int(), C.c_int()
    names = C.POINTER(C.c_char_p)()
    offs, sizes = _lib.c_i64_p(), _lib.c_i64_p()
    _lib.check(lib.sgpr_train_layout(C.byref(count), C.byref(n_params), C.byref(names), C.byref(offs), C.byref(sizes)),
               "sgpr_train_layout", lib)
    return [(names[i].decode(), int(offs[i]), int(sizes[i])) for i in range(count.value)], n_params.value


class TrainEngine:
    """Device-resident parameters + Adam moments; `step()` is one optimiser step on a batch of ordered pairs."""

    def __init__(self, device=0, lib=None):
        self._lib = lib or _lib.load()
        self._emulated = lib is not None           # tests/emu only: buffers are host memory
        if self._emulated:
            self.device = torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("sg_pr_b200 needs a CUDA device (B200, sm_100a); the training step has no CPU path")
            dev = torch.device(f"cuda:{device}" if isinstance(device, int) else device)
            if dev.type != "cuda":
                raise RuntimeError(f"sg_pr_b200 training engine cannot run on {dev}")
            self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.layout, self.n_param_tensors = layout(self._lib)
        self.n_params = int(self._lib.sgpr_train_param_count())
        self.n_state = int(self._lib.sgpr_train_state_count())
        handle = C.c_void_p()
        _lib.check(self._lib.sgpr_train_create(C.byref(handle), 0 if self._emulated else self.device.index),
                   "sgpr_train_create", self._lib)
        self._h = handle
        self._shape = {}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sgpr_train_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ------------------------------------------------------------------------------------------------------
    def set_state(self, state: Dict[str, torch.Tensor], reset_optimizer: bool = True):
        flat = np.zeros(self.n_state, dtype=np.float32)
        for name, off, size in self.layout:
            v = state[name].detach().to("cpu", torch.float32).contiguous().numpy().reshape(-1)
            if v.size != size:
                raise RuntimeError(f"size mismatch for {name}: checkpoint has {v.size} values, the kernels expect {size}")
            flat[off:off + size] = v
            self._shape[name] = tuple(state[name].shape)
        _lib.check(self._lib.sgpr_train_set_state(self._h, flat.ctypes.data_as(C.c_void_p), int(reset_optimizer)),
                   "sgpr_train_set_state", self._lib)

    def _unflatten(self, flat: np.ndarray, entries) -> Dict[str, torch.Tensor]:
        out = {}
        for name, off, size in entries:
            shape = self._shape.get(name) or SHAPES.get(name) or (size,)
            out[name] = torch.from_numpy(flat[off:off + size].copy()).reshape(shape)
        return out

    def get_state(self) -> Dict[str, torch.Tensor]:
        flat = np.empty(self.n_state, dtype=np.float32)
        _lib.check(self._lib.sgpr_train_get_state(self._h, flat.ctypes.data_as(C.c_void_p)), "sgpr_train_get_state", self._lib)
        return self._unflatten(flat, self.layout)

    def grads(self) -> Dict[str, torch.Tensor]:
        flat = np.empty(self.n_params, dtype=np.float32)
        _lib.check(self._lib.sgpr_train_get_grads(self._h, flat.ctypes.data_as(C.c_void_p)), "sgpr_train_get_grads", self._lib)
        return self._unflatten(flat, self.layout[:self.n_param_tensors])

    def set_optimizer(self, lr: float, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8):
        _lib.check(self._lib.sgpr_train_set_optimizer(self._h, lr, weight_decay, betas[0], betas[1], eps),
                   "sgpr_train_set_optimizer", self._lib)

    # ---- the step -----------------------------------------------------------------------------------------------------
    def _buf(self, t: torch.Tensor, what: str) -> torch.Tensor:
        if t.dtype != torch.float32:
            raise TypeError(f"{what} must be float32")
        if t.device != self.device:
            raise RuntimeError(f"{what} must live on {self.device} (got {t.device})")
        return t.contiguous()

    def step(self, f1: torch.Tensor, f2: Optional[torch.Tensor], target: torch.Tensor, k: int, apply: bool = True,
             mirrored: bool = False):
        """f1, f2 [B,15,N], target [B] on the device -> (loss [1], prediction [B]) device tensors (pre-update).
        mirrored=True: the caller guarantees f2[p] == f1[p ^ 1] (process_batch's doubling); f2 is then not read."""
        f1, target = self._buf(f1, "features_1"), self._buf(target, "target")
        f2 = f1 if (mirrored and f2 is None) else self._buf(f2, "features_2")
        if f1.dim() != 3 or f1.shape[1] != 15 or f2.shape != f1.shape or target.shape != (f1.shape[0],):
            raise ValueError(f"expected features [B,15,N] x2 and target [B], got {tuple(f1.shape)}, {tuple(f2.shape)}, "
                             f"{tuple(target.shape)}")
        B, _, N = f1.shape
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        pred = torch.empty(B, dtype=torch.float32, device=self.device)
        stream = None if self._emulated else C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.sgpr_train_step(self._h, f1.data_ptr(), f2.data_ptr(), target.data_ptr(), B, N, int(k),
                                             loss.data_ptr(), pred.data_ptr(), int(bool(apply)) | (2 if mirrored else 0),
                                             stream),
                   "sgpr_train_step", self._lib)
        return loss, pred

    def _stream(self):
        return None if self._emulated else C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- the step split at the loss (caller-driven autograd) -----------------------------------------------------------
    def set_state_flat(self, flat: torch.Tensor):
        """Flat state vector [n_state] already on the engine's device (layout order)."""
        flat = self._buf(flat, "state")
        if flat.numel() != self.n_state:
            raise ValueError(f"state vector has {flat.numel()} values, expected {self.n_state}")
        _lib.check(self._lib.sgpr_train_set_state_dev(self._h, flat.data_ptr(), self._stream()), "sgpr_train_set_state_dev",
                   self._lib)

    def get_state_flat(self) -> torch.Tensor:
        flat = torch.empty(self.n_state, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.sgpr_train_get_state_dev(self._h, flat.data_ptr(), self._stream()), "sgpr_train_get_state_dev",
                   self._lib)
        return flat

    def forward(self, f1: torch.Tensor, f2: Optional[torch.Tensor], k: int, update_running: bool = True,
                mirrored: bool = False, want_att: bool = True):
        """Train-mode forward only -> (prediction [B], att_1 [B,N,1] | None, att_2 [B,N,1] | None).
        The caller keeps f1 / f2 alive until backward(): the first EdgeConv layers' backward re-reads them."""
        f1 = self._buf(f1, "features_1")
        f2 = f1 if (mirrored and f2 is None) else self._buf(f2, "features_2")
        if f1.dim() != 3 or f1.shape[1] != 15 or f2.shape != f1.shape:
            raise ValueError(f"expected features [B,15,N] x2, got {tuple(f1.shape)}, {tuple(f2.shape)}")
        B, _, N = f1.shape
        pred = torch.empty(B, dtype=torch.float32, device=self.device)
        att1 = torch.empty(B, N, 1, dtype=torch.float32, device=self.device) if want_att else None
        att2 = torch.empty(B, N, 1, dtype=torch.float32, device=self.device) if (want_att and not mirrored) else None
        _lib.check(self._lib.sgpr_train_forward(self._h, f1.data_ptr(), f2.data_ptr(), B, N, int(k), pred.data_ptr(),
                                                att1.data_ptr() if att1 is not None else None,
                                                att2.data_ptr() if att2 is not None else None,
                                                int(bool(update_running)) | (2 if mirrored else 0), self._stream()),
                   "sgpr_train_forward", self._lib)
        if want_att and mirrored:
            att2 = att1.view(B // 2, 2, N, 1).flip(1).reshape(B, N, 1)
        return pred, att1, att2

    def backward(self, dpred: torch.Tensor) -> torch.Tensor:
        """d loss / d prediction [B] -> flat gradient vector [n_params] of the last forward()."""
        dpred = self._buf(dpred, "dpred")
        grads = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.sgpr_train_backward(self._h, dpred.data_ptr(), grads.data_ptr(), self._stream()),
                   "sgpr_train_backward", self._lib)
        return grads

    def assemble(self, graphs: torch.Tensor, pair_idx: torch.Tensor, seed: int, step: int, want_draws: bool = False,
                 rows_filled: Optional[int] = None):
        """Device batch assembly + augmentation: graphs [M,15,N] and pair_idx [P,2] int32 on the device -> features_1
        [2P,15,N] of the mirrored batch (row 2p / 2p+1 = augmented graph a / b of listed pair p).  want_draws also
        returns the random draws (draws [2P,12], raw jitter normals [2P,N,3]) for the parity tests.  Indices are the
        caller's responsibility (SGTrainer validates them on the host from its row table); the library rejects
        indices >= M, with M = rows_filled when the table has spare capacity."""
        graphs = self._buf(graphs, "graphs")
        if pair_idx.dtype != torch.int32 or pair_idx.device != self.device or pair_idx.dim() != 2 or pair_idx.shape[1] != 2:
            raise ValueError("pair_idx must be an int32 [P, 2] tensor on the engine's device")
        pair_idx = pair_idx.contiguous()
        M, _, N = graphs.shape
        if rows_filled is not None:                     # a table with spare capacity: only the first rows hold graphs
            M = min(M, int(rows_filled))
        P = int(pair_idx.shape[0])
        if want_draws and P and (int(pair_idx.min()) < 0 or int(pair_idx.max()) >= M):   # debug/parity mode only: two syncs
            raise IndexError("pair_idx out of range")
        out = torch.empty(2 * P, 15, N, dtype=torch.float32, device=self.device)
        draws = torch.zeros(2 * P, 12, dtype=torch.float32, device=self.device) if want_draws else None
        jitter = torch.zeros(2 * P, N, 3, dtype=torch.float32, device=self.device) if want_draws else None
        stream = None if self._emulated else C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.sgpr_train_assemble(self._h, graphs.data_ptr(), M, N, pair_idx.data_ptr(), P, int(seed), int(step),
                                                 out.data_ptr(), draws.data_ptr() if want_draws else None,
                                                 jitter.data_ptr() if want_draws else None, stream),
                   "sgpr_train_assemble", self._lib)
        return (out, draws, jitter) if want_draws else out

    def step_count(self) -> int:
        return int(self._lib.sgpr_train_step_count(self._h))

    def forward_generation(self) -> int:
        """How many train-mode forwards this context has run.  The context keeps the activations of the LAST one only;
        an autograd node records the value after its forward and refuses to run backward if it has moved on."""
        return int(self._lib.sgpr_train_forward_generation(self._h))

    def set_knn_ties(self, mode: str):
        """k-NN tie rule of the train-mode forward: "cuda" (default) or "cpu" — see Engine.set_knn_ties."""
        codes = {"cuda": 0, "cpu": 1}
        if mode not in codes:
            raise ValueError(f"knn_ties must be one of {sorted(codes)}, got {mode!r}")
        _lib.check(self._lib.sgpr_train_set_knn_ties(self._h, codes[mode]), "sgpr_train_set_knn_ties", self._lib)

    def launch_count(self) -> int:
        return int(self._lib.sgpr_train_launch_count(self._h))

    def debug_read(self, what: str, layer: int, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        n = self._lib.sgpr_train_debug_read(self._h, what.encode(), layer, out.ctypes.data_as(C.c_void_p), out.nbytes)
        if n < 0:
            _lib.check(int(n), "sgpr_train_debug_read", self._lib)
        if n != out.nbytes:
            raise RuntimeError(f"debug_read({what}): expected {out.nbytes} bytes, library holds {n}")
        return out
