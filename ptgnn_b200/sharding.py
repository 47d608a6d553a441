"""Multi-GPU sharding of the message-passing path; one process per GPU, ``torch.distributed``.

Two strategies, both leave the per-layer kernels untouched:

* **Graph-granular** (`partition_by_graph`, `shard_graphs`): a minibatch is block-diagonal -- graphs are contiguous node
  ranges and every edge is intra-graph (`/root/reference/ptgnn/neuralmodels/gnn/graphneuralnetwork.py:397-438`) -- so
  cutting at graph boundaries gives shards with an EMPTY halo: each rank runs the single-GPU path on its own graphs and
  no collective touches the data path.  This is also how the reference itself scales (DDP over file shards,
  `baseneuralmodel/distributedtrainer.py:297-306`).
* **Node-range** (`row_shard`, `RowShardedLayerLoop`): when ONE connected graph must be split, rank r owns the target
  rows `[lo_r, hi_r)` and all edges pointing into them (edge cut by target => the aggregation is local, no
  reduce-scatter); sources may live anywhere, so every layer starts with ONE all-gather of the state shards
  (`dist.all_gather_into_tensor`, NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Adjacency = List[Tuple[torch.Tensor, torch.Tensor]]


# ---------------------------------------------------------------------------------------------------
# graph-granular sharding (no collective)
# ---------------------------------------------------------------------------------------------------
@dataclass
class GraphShard:
    node_lo: int
    node_hi: int
    graph_lo: int
    graph_hi: int
    adjacency_lists: Adjacency          # node ids renumbered to [0, node_hi - node_lo)
    node_to_graph_idx: torch.Tensor     # graph ids renumbered to [0, graph_hi - graph_lo)

    @property
    def num_nodes(self) -> int:
        return self.node_hi - self.node_lo

    @property
    def num_graphs(self) -> int:
        return self.graph_hi - self.graph_lo


def partition_by_graph(node_to_graph_idx: torch.Tensor, adjacency_lists: Adjacency, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous graph ranges ``[(g_lo, g_hi)] * world_size`` balanced by per-graph cost = nodes + in-edges
    (the layer-level edge count after backward/self expansion is proportional to it)."""
    n2g = node_to_graph_idx.cpu()
    num_graphs = int(n2g[-1]) + 1 if n2g.numel() else 0
    cost = torch.bincount(n2g, minlength=num_graphs).to(torch.float64)
    for _, tgt in adjacency_lists:
        if tgt.numel():
            cost += 2.0 * torch.bincount(n2g[tgt.cpu()], minlength=num_graphs).to(torch.float64)
    prefix = torch.cumsum(cost, 0)
    total = float(prefix[-1]) if num_graphs else 0.0
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        g = int(torch.searchsorted(prefix, torch.tensor(target, dtype=torch.float64)).item())
        cuts.append(min(max(g, cuts[-1]), num_graphs))
    cuts.append(num_graphs)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def shard_graphs(node_to_graph_idx: torch.Tensor, adjacency_lists: Adjacency, graph_range: Tuple[int, int]) -> GraphShard:
    """Extracts the graphs ``[g_lo, g_hi)`` as a self-contained batch (ids renumbered from 0)."""
    g_lo, g_hi = graph_range
    n2g = node_to_graph_idx
    bounds = torch.searchsorted(n2g.contiguous(), torch.tensor([g_lo, g_hi], dtype=n2g.dtype, device=n2g.device))
    lo, hi = int(bounds[0]), int(bounds[1])
    adj: Adjacency = []
    for src, tgt in adjacency_lists:
        keep = (tgt >= lo) & (tgt < hi)
        s, t = src[keep], tgt[keep]
        if s.numel() and (int(s.min()) < lo or int(s.max()) >= hi):
            raise ValueError("an edge crosses the shard boundary: the batch is not block-diagonal; use row_shard()")
        adj.append((s - lo, t - lo))
    return GraphShard(lo, hi, g_lo, g_hi, adj, n2g[lo:hi] - g_lo)


# ---------------------------------------------------------------------------------------------------
# node-range sharding (one all-gather per layer)
# ---------------------------------------------------------------------------------------------------
@dataclass
class RowShard:
    lo: int
    hi: int
    rows_per_rank: int                  # equal chunk size used by the all-gather (last rank is padded)
    num_nodes: int                      # global node count
    adjacency_lists: Adjacency          # (src GLOBAL id, tgt LOCAL id = tgt - lo), edges with lo <= tgt < hi

    @property
    def num_local(self) -> int:
        return self.hi - self.lo


def row_shard(num_nodes: int, adjacency_lists: Adjacency, world_size: int, rank: int) -> RowShard:
    rows = (num_nodes + world_size - 1) // world_size
    lo, hi = min(rank * rows, num_nodes), min((rank + 1) * rows, num_nodes)
    adj: Adjacency = []
    for src, tgt in adjacency_lists:
        keep = (tgt >= lo) & (tgt < hi)
        adj.append((src[keep].contiguous(), (tgt[keep] - lo).contiguous()))
    return RowShard(lo, hi, rows, num_nodes, adj)


LayerFn = Callable[[torch.Tensor, torch.Tensor, Adjacency], torch.Tensor]


class RowShardedLayerLoop:
    """Runs L layers on a node-range shard.  ``layer_fns[l](own_states, gathered_states, shard_adjacency)`` computes the
    new states of the OWNED rows; on the GPU it is ``lambda h, g, adj: layer(h, adj, gather_states=g)`` with a
    ptgnn_b200 layer, in the CPU (gloo) tests it is the oracle restricted to the shard."""

    def __init__(self, shard: RowShard, group: Optional[dist.ProcessGroup] = None):
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_gather_states(self, own: torch.Tensor) -> torch.Tensor:
        s = self.shard
        if self.world == 1:
            return own
        padded = own
        if own.shape[0] != s.rows_per_rank:   # last rank: pad to the common chunk size
            padded = torch.zeros(s.rows_per_rank, own.shape[1], dtype=own.dtype, device=own.device)
            padded[: own.shape[0]] = own
        full = torch.empty(self.world * s.rows_per_rank, own.shape[1], dtype=own.dtype, device=own.device)
        dist.all_gather_into_tensor(full, padded.contiguous(), group=self.group)
        return full[: s.num_nodes]

    def run(self, own_states: torch.Tensor, layer_fns: Sequence[LayerFn]) -> torch.Tensor:
        h = own_states
        for fn in layer_fns:
            gathered = self.all_gather_states(h)          # the ONE collective of the layer
            h = fn(h, gathered, self.shard.adjacency_lists)
        return h
