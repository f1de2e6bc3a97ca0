"""ctypes binding of include/sgpr_b200.h.  Fails loudly when the library is missing: there is no fallback."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGPR_B200_LIB") or os.path.join(PKG, "libsgpr_b200.so")

SGPR_OK = 0
c_float_p = C.POINTER(C.c_float)


class SgprBn(C.Structure):
    _fields_ = [("weight", c_float_p), ("bias", c_float_p), ("running_mean", c_float_p), ("running_var", c_float_p)]


class SgprWeights(C.Structure):
    _fields_ = [
        ("s_conv_w", c_float_p * 3), ("s_bn", SgprBn * 3),
        ("f_conv_w", c_float_p * 3), ("f_bn", SgprBn * 3),
        ("end_conv_w", c_float_p), ("end_bn", SgprBn),
        ("att_w", c_float_p), ("ntn_w", c_float_p), ("ntn_v", c_float_p), ("ntn_b", c_float_p),
        ("fc1_w", c_float_p), ("fc1_b", c_float_p), ("fc2_w", c_float_p), ("fc2_b", c_float_p),
        ("bn_eps", C.c_float), ("filters", C.c_int32 * 3), ("tensor_neurons", C.c_int32), ("bottleneck", C.c_int32),
    ]


# every symbol include/sgpr_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sgpr_abi_version": (C.c_int, []),
    "sgpr_last_error": (C.c_char_p, []),
    "sgpr_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "sgpr_destroy": (C.c_int, [C.c_void_p]),
    "sgpr_set_weights": (C.c_int, [C.c_void_p, C.POINTER(SgprWeights)]),
    "sgpr_forward_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_forward_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "sgpr_embed_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_score_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sgpr_score_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p]),
    "sgpr_score_matrix_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                          C.c_int64, C.c_void_p]),
    "sgpr_enable_peer_access": (C.c_int, [C.c_void_p, C.c_int]),
    "sgpr_peer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "sgpr_peer_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "sgpr_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgpr_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgpr_packed_size": (C.c_size_t, []),
    "sgpr_pack_weights_host": (C.c_int, [C.POINTER(SgprWeights), c_float_p, c_float_p, C.POINTER(C.c_size_t)]),
    "sgpr_launch_count": (C.c_int64, [C.c_void_p]),
    "sgpr_compact_stride": (C.c_size_t, [C.c_int]),
    "sgpr_compact_from_blocks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sgpr_forward_pairs_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_embed_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "sgpr_set_knn_ties": (C.c_int, [C.c_void_p, C.c_int]),
    "sgpr_get_knn_ties": (C.c_int, [C.c_void_p]),
    "sgpr_topk_cpu_rule_host": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
}

# every symbol include/sgpr_b200_train.h declares
c_i64_p = C.POINTER(C.c_int64)
TRAIN_SYMBOLS = {
    "sgpr_train_layout": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_char_p)),
                                    C.POINTER(c_i64_p), C.POINTER(c_i64_p)]),
    "sgpr_train_param_count": (C.c_int64, []),
    "sgpr_train_state_count": (C.c_int64, []),
    "sgpr_train_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "sgpr_train_destroy": (C.c_int, [C.c_void_p]),
    "sgpr_train_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "sgpr_train_get_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgpr_train_set_optimizer": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "sgpr_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sgpr_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "sgpr_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_train_set_state_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_train_get_state_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_train_assemble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgpr_train_get_grads": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgpr_train_forward_generation": (C.c_int64, [C.c_void_p]),
    "sgpr_train_set_knn_ties": (C.c_int, [C.c_void_p, C.c_int]),
    "sgpr_train_get_knn_ties": (C.c_int, [C.c_void_p]),
    "sgpr_train_step_count": (C.c_int64, [C.c_void_p]),
    "sgpr_train_launch_count": (C.c_int64, [C.c_void_p]),
    "sgpr_train_debug_read": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]),
}
SYMBOLS.update(TRAIN_SYMBOLS)


def bind(lib, symbols):
    """Type the given entry points of an already opened library (AttributeError if one is not exported)."""
    for name, (res, args) in symbols.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load():
    """dlopen the in-tree library and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sg_pr_b200.build` (or __graft_entry__.build()). "
            "sg_pr_b200 has no CPU or PyTorch fallback for the SG_PR hot path.")
    lib = bind(C.CDLL(LIB_PATH), SYMBOLS)     # AttributeError if the library does not export a declared symbol
    _lib = lib
    return lib


class SgprError(RuntimeError):
    pass


def check(rc: int, what: str, lib=None):
    if rc != SGPR_OK:
        msg = (lib or load()).sgpr_last_error()
        raise SgprError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
