"""``EGCMessagePassingLayer`` (EGC-S with per-edge-type bases) on the native operators -- SURVEY.md §8 row f-4.

Same constructor, attribute names (hence ``state_dict`` keys) and forward contract as
`/root/reference/ptgnn/neuralmodels/gnn/messagepassing/egcmessagepassing.py:8-99`.  The layer is the hot path's
gather -> per-type Linear -> scatter with a wider message (``num_bases * output_state_dimension`` columns, :75-83), followed by a
per-node combination of the aggregated bases with weights from one more Linear (:64-66, :90): composed from the stand-alone native
kernels (``edge_messages``, ``segment_reduce``, ``linear``); the final weighted sum over the bases is a node-sized pointwise op.
Forward only (eval mode, or training mode with ``dropout_rate == 0``, under ``no_grad``); fp32 states.
"""
from typing import Dict, List, Tuple

import torch
from torch import nn

from . import _native as N
from . import composed as C
from .messagepassing import AbstractMessagePassingLayer, _check_states, _reduce_code, _refuse_autograd


class EGCMessagePassingLayer(AbstractMessagePassingLayer):
    def __init__(self, input_state_dimension: int, output_state_dimension: int, num_edge_types: int, message_aggregation_function: str,
                 num_bases: int = 4, num_heads: int = 8, dropout_rate: float = 0.0):
        super().__init__()
        if output_state_dimension % num_heads != 0:
            raise AssertionError("output_state_dimension must be a multiple of num_heads")
        self._dims = (int(input_state_dimension), int(output_state_dimension), int(num_heads), int(num_bases))
        self._reduce_name = message_aggregation_function
        self._drop_p = float(dropout_rate)
        # the two parameter holders keep the reference's (name-mangled) attribute names and construction order: identical
        # state_dict keys, identical parameters for the same seed
        self.__bases = nn.ModuleList(nn.Linear(input_state_dimension, num_bases * output_state_dimension, bias=False) for _ in range(num_edge_types))
        self.__weight_coeffs = nn.Linear(input_state_dimension, num_heads * num_bases)

    def forward(self, node_states: torch.Tensor, adjacency_lists: List[Tuple[torch.Tensor, torch.Tensor]],
                node_to_graph_idx: torch.Tensor = None, reference_node_ids: Dict[str, torch.Tensor] = None,
                reference_node_graph_idx: Dict[str, torch.Tensor] = None, edge_features: List[torch.Tensor] = None) -> torch.Tensor:
        in_dim, out, heads, bases = self._dims
        assert len(adjacency_lists) == len(self.__bases)
        if self.training and self._drop_p > 0:
            raise NotImplementedError("EGCMessagePassingLayer: training-mode dropout has no native path")
        _refuse_autograd(self, node_states)
        if node_states.dtype != torch.float32:
            raise NotImplementedError("EGCMessagePassingLayer: fp32 states only")
        _check_states(node_states, in_dim, "EGCMessagePassingLayer")
        reduce = _reduce_code(self._reduce_name)
        h = N.require_cuda(node_states, "node_states", torch.float32)
        n = h.shape[0]
        plan = self._plan(adjacency_lists, n, None)
        coeff = self.__weight_coeffs
        node_weights = C.linear(h, coeff.weight, coeff.bias).reshape(n, heads, bases, 1)                                   # :64-66
        messages = C.edge_messages(plan, h, None, [b.weight for b in self.__bases], False)        # [E, bases * out]           :75-83
        aggregated = C.segment_reduce(messages, plan, reduce).reshape(n, heads, bases, out // heads)                        # :85-89
        return (aggregated * node_weights).sum(dim=-2).reshape(n, out)                                                      # :90

    @property
    def input_state_dimension(self) -> int:
        return self._dims[0]

    @property
    def output_state_dimension(self) -> int:
        return self._dims[1]
