"""Skip-connection pseudo-layers that sit in a ``GraphNeuralNetwork`` layer list between real message-passing layers.

Same constructor / property / ``pass_through_dummy_layer()`` contract as the reference's
`/root/reference/ptgnn/neuralmodels/gnn/messagepassing/residuallayers.py:8-137` so that the factories of
`implementations/typilus/train.py:39-99`, `ppi/train.py:36-57`, `varmisuse/train.py:42-107` can be written against ptgnn_b200
alone.  They move no edges: a *tap* layer remembers the states flowing through it, the matching *join* layer later combines the
remembered states with the current ones (mean / concatenation / bias-free Linear over the concatenation).  Plain device-side
tensor ops -- nothing here is on the kernel path; the reference's own classes work just as well in a ptgnn_b200 container.
"""
from typing import Optional

import torch
from torch import nn

from .messagepassing import AbstractMessagePassingLayer


class _ResidualOriginLayer(AbstractMessagePassingLayer):
    """The tap: hands its input to `join` and returns it unchanged."""

    def __init__(self, input_dim: int, target_layer: "_JoinLayer"):
        super().__init__()
        self.__target_layer = target_layer  # registered as a submodule, like the reference: same state_dict keys
        self.__dim = input_dim

    def forward(self, node_states, adjacency_lists=None, node_to_graph_idx=None, reference_node_ids=None,
                reference_node_graph_idx=None, edge_features=None) -> torch.Tensor:
        self.__target_layer._original_input = node_states
        return node_states

    @property
    def input_state_dimension(self) -> int:
        return self.__dim

    @property
    def output_state_dimension(self) -> int:
        return self.__dim


class _JoinLayer(AbstractMessagePassingLayer):
    def __init__(self, tap_dim: int):
        super().__init__()
        self._original_input: Optional[torch.Tensor] = None
        self._tap_dim = tap_dim

    def pass_through_dummy_layer(self) -> _ResidualOriginLayer:
        return _ResidualOriginLayer(self._tap_dim, target_layer=self)

    def _take(self) -> torch.Tensor:
        assert self._original_input is not None, "Initial Pass Through Layer was not used."
        remembered, self._original_input = self._original_input, None
        return remembered

    def _combine(self, remembered: torch.Tensor, current: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, node_states, adjacency_lists=None, node_to_graph_idx=None, reference_node_ids=None,
                reference_node_graph_idx=None, edge_features=None) -> torch.Tensor:
        return self._combine(self._take(), node_states)


class MeanResidualLayer(_JoinLayer):
    """out = (remembered + current) / 2, computed as the reference does (stack + mean)."""

    def __init__(self, input_dim: int):
        super().__init__(input_dim)

    def _combine(self, remembered, current):
        return torch.stack((remembered, current), dim=-1).mean(dim=-1)

    @property
    def input_state_dimension(self) -> int:
        return self._tap_dim

    @property
    def output_state_dimension(self) -> int:
        return self._tap_dim


class ConcatResidualLayer(_JoinLayer):
    """out = [remembered ; current]  (doubles the state width)."""

    def __init__(self, input_dim: int):
        super().__init__(input_dim)

    def _combine(self, remembered, current):
        return torch.cat((remembered, current), dim=-1)

    @property
    def input_state_dimension(self) -> int:
        return self._tap_dim

    @property
    def output_state_dimension(self) -> int:
        return 2 * self._tap_dim


class LinearResidualLayer(_JoinLayer):
    """out = Dropout(W [remembered ; current]), W bias-free."""

    def __init__(self, state_dimension1: int, state_dimension2: int, target_state_size: int, dropout_rate: float = 0.0):
        super().__init__(state_dimension1)
        self.__input_dim2 = state_dimension2
        self.__linear_combination = nn.Linear(state_dimension1 + state_dimension2, target_state_size, bias=False)
        self.__dropout = nn.Dropout(p=dropout_rate)

    def _combine(self, remembered, current):
        return self.__dropout(self.__linear_combination(torch.cat((remembered, current), dim=-1)))

    @property
    def input_state_dimension(self) -> int:
        return self.__input_dim2

    @property
    def output_state_dimension(self) -> int:
        return self.__linear_combination.out_features
