"""Seeded synthetic minibatches shaped like the reference's workloads.

The reference batches many graphs into one block-diagonal graph: nodes of graph ``g`` occupy a contiguous
id range, every edge is intra-graph, per-type edge lists are graph-major and NOT sorted by target
(`/root/reference/ptgnn/neuralmodels/gnn/graphneuralnetwork.py:386-438,445-493`).  These generators emit
exactly that layout (int64 index tensors, CPU) so the same data feeds the CUDA path and the CPU oracle.
"""
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

Adjacency = List[Tuple[torch.Tensor, torch.Tensor]]


@dataclass
class GraphBatch:
    """One flattened minibatch of graphs (raw edge types only; the container adds backward/self edges)."""

    num_nodes: int
    num_graphs: int
    adjacency_lists: Adjacency  # per raw edge type: (src [E_t] int64, tgt [E_t] int64)
    node_to_graph_idx: torch.Tensor  # [N] int64, non-decreasing

    @property
    def num_raw_edges(self) -> int:
        return sum(int(a[0].shape[0]) for a in self.adjacency_lists)

    def layer_level_edges(self, introduce_backwards_edges: bool = True, add_self_edges: bool = True) -> int:
        """``num_edges`` as the reference counts it (graphneuralnetwork.py:198-199)."""
        e = self.num_raw_edges * (2 if introduce_backwards_edges else 1)
        return e + (self.num_nodes if add_self_edges else 0)


def _split(total: int, fractions: Sequence[float]) -> List[int]:
    counts = [int(total * f) for f in fractions]
    counts[0] += total - sum(counts)
    return counts


def block_diagonal_batch(
    num_graphs: int,
    nodes_per_graph: int,
    raw_edges: int,
    type_fractions: Sequence[float],
    seed: int,
    local_fraction: float = 0.5,
    local_span: int = 32,
) -> GraphBatch:
    """``num_graphs`` graphs of ``nodes_per_graph`` nodes; per type, edges are split evenly over graphs;
    within a graph ``local_fraction`` of the edges are "local" (tgt = src + U[1, local_span], clipped to the
    graph) and the rest uniform random intra-graph pairs."""
    gen = torch.Generator().manual_seed(seed)
    per_type = _split(raw_edges, type_fractions)
    adjacency: Adjacency = []
    for n_t in per_type:
        per_graph = _split(n_t, [1.0 / num_graphs] * num_graphs)
        srcs, tgts = [], []
        for g, n_e in enumerate(per_graph):
            base = g * nodes_per_graph
            src = torch.randint(0, nodes_per_graph, (n_e,), generator=gen)
            n_local = int(n_e * local_fraction)
            delta = torch.randint(1, local_span + 1, (n_local,), generator=gen)
            tgt_local = torch.clamp(src[:n_local] + delta, max=nodes_per_graph - 1)
            tgt_unif = torch.randint(0, nodes_per_graph, (n_e - n_local,), generator=gen)
            tgt = torch.cat([tgt_local, tgt_unif])
            shuffle = torch.randperm(n_e, generator=gen)
            srcs.append(src[shuffle] + base)
            tgts.append(tgt[shuffle] + base)
        adjacency.append((torch.cat(srcs).to(torch.int64), torch.cat(tgts).to(torch.int64)))
    n = num_graphs * nodes_per_graph
    node_to_graph = torch.arange(num_graphs, dtype=torch.int64).repeat_interleave(nodes_per_graph)
    return GraphBatch(n, num_graphs, adjacency, node_to_graph)


def single_random_graph(num_nodes: int, num_edges: int, num_types: int, seed: int) -> GraphBatch:
    """One connected-ish uniform random graph (NOT block diagonal): config 5 of BASELINE.json."""
    gen = torch.Generator().manual_seed(seed)
    per_type = _split(num_edges, [1.0 / num_types] * num_types)
    adjacency = [
        (
            torch.randint(0, num_nodes, (n_t,), generator=gen, dtype=torch.int64),
            torch.randint(0, num_nodes, (n_t,), generator=gen, dtype=torch.int64),
        )
        for n_t in per_type
    ]
    return GraphBatch(num_nodes, 1, adjacency, torch.zeros(num_nodes, dtype=torch.int64))


# ---- named configurations of BASELINE.json ------------------------------------------------------
GRAPH2CLASS_FRACTIONS = (0.30, 0.25, 0.15, 0.10, 0.08, 0.06, 0.04, 0.02)


def graph2class_batch(num_graphs: int = 80, seed: int = 1234) -> GraphBatch:
    """Config 2: 80 x 2,560 nodes = 204,800 nodes, 8 raw types, 450,560 raw edges
    => T = 17, E = 1,105,920 after backward + self edges (hidden 128, 8 GatedMP layers)."""
    return block_diagonal_batch(num_graphs, 2560, 5632 * num_graphs, GRAPH2CLASS_FRACTIONS, seed)


def varmisuse_batch(num_graphs: int = 40, seed: int = 1235) -> GraphBatch:
    """Config 3: 40 x 2,000 nodes, 11 raw types (geometric split), 200,000 raw edges => T = 23, E = 480,000."""
    weights = [0.75**i for i in range(11)]
    fractions = [w / sum(weights) for w in weights]
    return block_diagonal_batch(num_graphs, 2000, 5000 * num_graphs, fractions, seed)


def ppi_batch(seed: int = 0) -> GraphBatch:
    """Config 1: 2 graphs x 1,500 nodes, one raw type, 22,500 uniform intra-graph edges each => T=3, E=93,000."""
    return block_diagonal_batch(2, 1500, 45000, (1.0,), seed, local_fraction=0.0)
