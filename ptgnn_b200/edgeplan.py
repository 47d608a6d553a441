"""Edge plan: the per-minibatch integer bookkeeping shared by every message-passing layer.

The reference re-derives the same information inside every layer call: it concatenates the per-type target lists
(`/root/reference/ptgnn/neuralmodels/gnn/messagepassing/gatedmessagepassing.py:46`,
`mlpmessagepassing.py:102-109`) and lets ``torch_scatter.scatter`` group rows by target
(`abstractmessagepassing.py:44-50`).  Here that grouping is computed ONCE per minibatch on the GPU
(``ptgnn_b200_plan_build``: int64->int32, degree histogram, scan, stable radix sort by target) and reused by all
L layers (the reference's weight-shared stacks call the same layer 7-8 times on the same adjacency).
"""
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N

Adjacency = Sequence[Tuple[torch.Tensor, torch.Tensor]]


class EdgePlan:
    """Device-resident CSR-by-target plan (all int32 unless noted).  Field meanings: include/ptgnn_b200.h."""

    __slots__ = (
        "num_nodes", "num_source_nodes", "num_edges", "num_types", "type_off", "type_off_c", "row_ptr", "perm", "pos", "src_sorted",
        "etype_sorted", "src32", "tgt32", "status", "device", "_keepalive", "_validated",
    )

    def __init__(self, adjacency_lists: Adjacency, num_nodes: int, validate: bool = False,
                 num_source_nodes: Optional[int] = None):
        """``num_nodes`` = number of TARGET rows (CSR rows).  ``num_source_nodes`` (default: the same) bounds the source
        ids; it differs only for node-range shards, where targets are local rows and sources index the gathered states."""
        if len(adjacency_lists) > 128:
            raise NotImplementedError("more than 128 edge types")
        if len(adjacency_lists) == 0:
            raise ValueError("at least one edge type is required")
        device = adjacency_lists[0][0].device
        srcs = [N.require_cuda(s, f"adjacency_lists[{i}][0]", torch.int64) for i, (s, _) in enumerate(adjacency_lists)]
        tgts = [N.require_cuda(t, f"adjacency_lists[{i}][1]", torch.int64) for i, (_, t) in enumerate(adjacency_lists)]
        counts = [int(s.shape[0]) for s in srcs]
        for s, t in zip(srcs, tgts):
            if s.dim() != 1 or t.shape != s.shape:
                raise ValueError("adjacency lists must be pairs of equal-length 1-D tensors")
        E = sum(counts)
        self.num_nodes, self.num_edges, self.num_types, self.device = int(num_nodes), E, len(counts), device
        self.num_source_nodes = int(num_source_nodes) if num_source_nodes is not None else int(num_nodes)
        self.type_off = [0]
        for c in counts:
            self.type_off.append(self.type_off[-1] + c)
        self.type_off_c = N.i64_array(self.type_off)

        def i32(n):
            return torch.empty(n, dtype=torch.int32, device=device)

        self.row_ptr = i32(num_nodes + 1)
        self.perm, self.pos, self.src_sorted, self.src32, self.tgt32 = i32(E), i32(E), i32(E), i32(E), i32(E)
        self.etype_sorted = torch.empty(E, dtype=torch.uint8, device=device)
        self.status = i32(1)
        lib = N.lib()
        ws_bytes = lib.ptgnn_b200_plan_workspace_bytes(num_nodes, E)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            rc = lib.ptgnn_b200_plan_build(
                num_nodes, self.num_source_nodes, len(counts), N.ptr_table(srcs), N.ptr_table(tgts), N.i64_array(counts),
                N.ptr(self.row_ptr), N.ptr(self.perm), N.ptr(self.pos), N.ptr(self.src_sorted), N.ptr(self.etype_sorted),
                N.ptr(self.src32), N.ptr(self.tgt32), N.ptr(self.status), N.ptr(ws), ws_bytes, N.current_stream(device),
            )
        N.check(rc, "ptgnn_b200_plan_build")
        self._keepalive = (srcs, tgts)
        self._validated = False
        if validate:
            self.validate()

    def validate(self) -> None:
        """Synchronises and raises IndexError if any edge index was outside [0, num_nodes)."""
        if not self._validated:
            bad = int(self.status.item())
            if bad:
                raise IndexError(f"{bad} edge indices outside [0, {self.num_nodes}) (targets) / [0, {self.num_source_nodes}) (sources)")
            self._validated = True


# ---- small identity-keyed cache so that the L layers of one forward share one plan ----------------
_CACHE: "OrderedDict[tuple, EdgePlan]" = OrderedDict()
_CACHE_SIZE = 4


def _key(adjacency_lists: Adjacency, num_nodes: int, num_source_nodes: Optional[int] = None) -> tuple:
    parts: List[int] = [num_nodes, -1 if num_source_nodes is None else num_source_nodes]
    for s, t in adjacency_lists:
        parts += [s.data_ptr(), s.shape[0], s._version, t.data_ptr(), t._version]
    return tuple(parts)


def plan_for(adjacency_lists: Adjacency, num_nodes: int, plan: Optional[EdgePlan] = None,
             num_source_nodes: Optional[int] = None) -> EdgePlan:
    """Returns the plan for these adjacency tensors, building it on a cache miss.  Entries keep their index tensors
    alive, so a (data_ptr, version) key cannot alias different contents."""
    if plan is not None:
        return plan
    key = _key(adjacency_lists, num_nodes, num_source_nodes)
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit
    built = EdgePlan(adjacency_lists, num_nodes, num_source_nodes=num_source_nodes)
    _CACHE[key] = built
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return built


def clear_plan_cache() -> None:
    _CACHE.clear()
