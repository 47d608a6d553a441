"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the graph-structure half of the reference's minibatch assembly.

Only tests/ may import this.  Follows, line by line, /root/reference/ptgnn/neuralmodels/gnn/graphneuralnetwork.py:
  initialize_minibatch   :372-384
  extend_minibatch_with  :386-438   (node / edge embedder calls left out: they are not part of the graph structure)
  finalize_minibatch     :445-493
Pinned by tests/golden/minibatch.npz, which tests/golden/generate_golden.py produces by calling the UNMODIFIED reference methods
(with stub embedders) on the same graphs (tests/test_oracle_batching.py).
"""
from collections import defaultdict
from typing import Any, Dict, List

import numpy as np


def initialize_minibatch(num_edge_types: int) -> Dict[str, Any]:          # :372-384
    return {
        "adjacency_lists": [([], []) for _ in range(num_edge_types)],
        "num_nodes_per_graph": [],
        "reference_node_graph_idx": defaultdict(list),
        "reference_node_ids": defaultdict(list),
        "num_nodes_in_mb": 0,
    }


def extend_minibatch_with(graph, partial: Dict[str, Any], stop_after_num_nodes: int) -> bool:      # :386-438
    graph_idx = len(partial["num_nodes_per_graph"])                        # :397
    so_far = partial["num_nodes_in_mb"]                                    # :401
    for (src, tgt), (mb_src, mb_tgt) in zip(graph.adjacency_lists, partial["adjacency_lists"]):
        mb_src.append(np.asarray(src) + so_far)                            # :419-421
        mb_tgt.append(np.asarray(tgt) + so_far)                            # :422-424
    for name, ref_nodes in graph.reference_nodes.items():                  # :431-436
        partial["reference_node_graph_idx"][name].extend(graph_idx for _ in range(len(ref_nodes)))
        partial["reference_node_ids"][name].append(np.asarray(ref_nodes) + so_far)
    partial["num_nodes_per_graph"].append(graph.num_nodes)                 # :438-439
    partial["num_nodes_in_mb"] = so_far + graph.num_nodes
    return partial["num_nodes_in_mb"] < stop_after_num_nodes


def finalize_minibatch(acc: Dict[str, Any]) -> Dict[str, Any]:             # :445-493 (numpy int64 instead of device tensors)
    def cat(parts: List[np.ndarray]) -> np.ndarray:
        return np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, dtype=np.int64)

    node_to_graph = [i for i, size in enumerate(acc["num_nodes_per_graph"]) for _ in range(size)]      # :441-443
    return {
        "adjacency_lists": [(cat(s), cat(t)) for s, t in acc["adjacency_lists"]],                       # :463-469
        "node_to_graph_idx": np.asarray(node_to_graph, dtype=np.int64),                                  # :471-480
        "reference_node_graph_idx": {k: np.asarray(v, dtype=np.int64) for k, v in acc["reference_node_graph_idx"].items()},
        "reference_node_ids": {k: np.concatenate(v).astype(np.int32).astype(np.int64) for k, v in acc["reference_node_ids"].items()},
        "num_graphs": len(acc["num_nodes_per_graph"]),                                                   # :492
    }
