"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_build/liboracle.so (see oracle/oracle.c)."""
import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2, "min": 3}
ACT = {None: 0, "gelu": 1, "tanh": 2, "relu": 3}
_lib = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr_table(arrs: Sequence[np.ndarray], ctype):
    tab = (ctypes.POINTER(ctype) * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        tab[i] = a.ctypes.data_as(ctypes.POINTER(ctype))
    return tab


def _adj(adjacency):
    srcs = [np.ascontiguousarray(np.asarray(s), dtype=np.int64) for s, _ in adjacency]
    tgts = [np.ascontiguousarray(np.asarray(t), dtype=np.int64) for _, t in adjacency]
    counts = np.array([len(s) for s in srcs], dtype=np.int64)
    return srcs, tgts, counts


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def edge_plan(adjacency, num_nodes: int):
    srcs, tgts, counts = _adj(adjacency)
    E = int(counts.sum())
    out = {
        "row_ptr": np.zeros(num_nodes + 1, np.int32),
        "perm": np.zeros(E, np.int32),
        "pos": np.zeros(E, np.int32),
        "src_sorted": np.zeros(E, np.int32),
        "etype_sorted": np.zeros(E, np.uint8),
    }
    rc = lib().oracle_edge_plan(
        ctypes.c_int64(num_nodes), ctypes.c_int32(len(srcs)), _ptr_table(srcs, ctypes.c_int64),
        _ptr_table(tgts, ctypes.c_int64), counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        out["row_ptr"].ctypes.data_as(ctypes.c_void_p), out["perm"].ctypes.data_as(ctypes.c_void_p),
        out["pos"].ctypes.data_as(ctypes.c_void_p), out["src_sorted"].ctypes.data_as(ctypes.c_void_p),
        out["etype_sorted"].ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise ValueError(f"oracle_edge_plan rc={rc}")
    return out


def scatter(src: np.ndarray, index: np.ndarray, dim_size: int, reduce: str) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    src = np.ascontiguousarray(src, dtype=np.float32)
    index = np.ascontiguousarray(index, dtype=np.int64)
    E, D = src.shape
    out = np.zeros((dim_size, D), np.float32)
    arg = np.zeros((dim_size, D), np.int64) if reduce in ("max", "min") else None
    rc = lib().oracle_scatter_f32(_f(src), index.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(E), ctypes.c_int64(D),
                                  ctypes.c_int64(dim_size), ctypes.c_int(REDUCE[reduce]), _f(out),
                                  arg.ctypes.data_as(ctypes.c_void_p) if arg is not None else None)
    if rc:
        raise ValueError(f"oracle_scatter_f32 rc={rc}")
    return out, arg


def gated_forward(h, adjacency, weights: List[np.ndarray], w_ih, w_hh, b_ih, b_hh, reduce: str) -> np.ndarray:
    h = np.ascontiguousarray(h, np.float32)
    N, H = h.shape
    ws = [np.ascontiguousarray(w, np.float32) for w in weights]
    D = ws[0].shape[0]
    srcs, tgts, counts = _adj(adjacency)
    w_ih, w_hh, b_ih, b_hh = (np.ascontiguousarray(a, np.float32) for a in (w_ih, w_hh, b_ih, b_hh))
    out = np.zeros((N, H), np.float32)
    rc = lib().oracle_gated_forward_f32(
        _f(h), ctypes.c_int64(N), ctypes.c_int64(H), ctypes.c_int64(D), ctypes.c_int32(len(ws)),
        _ptr_table(srcs, ctypes.c_int64), _ptr_table(tgts, ctypes.c_int64),
        counts.ctypes.data_as(ctypes.c_void_p), _ptr_table(ws, ctypes.c_float), _f(w_ih), _f(w_hh), _f(b_ih), _f(b_hh),
        ctypes.c_int(REDUCE[reduce]), _f(out))
    if rc:
        raise ValueError(f"oracle_gated_forward_f32 rc={rc}")
    return out


def mlp_forward(h, adjacency, weights: List[np.ndarray], reduce: str, *, use_target=True, message_activation="gelu",
                ln_weight=None, ln_bias=None, ln_eps=1e-5, dense_weight=None, dense_bias=None,
                dense_activation="tanh") -> np.ndarray:
    h = np.ascontiguousarray(h, np.float32)
    N, H = h.shape
    ws = [np.ascontiguousarray(w, np.float32) for w in weights]
    D = ws[0].shape[0]
    srcs, tgts, counts = _adj(adjacency)
    c = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)  # noqa: E731
    ln_weight, ln_bias, dense_weight, dense_bias = c(ln_weight), c(ln_bias), c(dense_weight), c(dense_bias)
    Hout = dense_weight.shape[0] if dense_weight is not None else D
    out = np.zeros((N, Hout), np.float32)
    rc = lib().oracle_mlp_forward_f32(
        _f(h), ctypes.c_int64(N), ctypes.c_int64(H), ctypes.c_int64(D), ctypes.c_int64(Hout), ctypes.c_int32(len(ws)),
        _ptr_table(srcs, ctypes.c_int64), _ptr_table(tgts, ctypes.c_int64), counts.ctypes.data_as(ctypes.c_void_p),
        _ptr_table(ws, ctypes.c_float), ctypes.c_int(int(use_target)), ctypes.c_int(REDUCE[reduce]),
        ctypes.c_int(ACT[message_activation]), _f(ln_weight), _f(ln_bias), ctypes.c_float(ln_eps),
        _f(dense_weight), _f(dense_bias), ctypes.c_int(ACT[dense_activation]), _f(out))
    if rc:
        raise ValueError(f"oracle_mlp_forward_f32 rc={rc}")
    return out
