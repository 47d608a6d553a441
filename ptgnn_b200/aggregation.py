"""Module aggregators for ``MlpMessagePassingLayer(message_aggregation_function=<module>)``.

``PnaMessageAggregation``: Principal Neighbourhood Aggregation (https://arxiv.org/abs/2004.05718) with the reference's conventions
(`/root/reference/ptgnn/neuralmodels/gnn/messagepassing/pna_aggregation.py:13-59`): per target the sum, mean (sum / (degree + 1e-5)),
max, min and a std built from ``relu(m^2 - mean[target]^2) + 1e-10``, concatenated and repeated under the three degree scalers
(identity, ``log(degree + 1) / delta``, its damped inverse) -> ``15 x message_dim`` features.  The five segmented reductions run on
the native segmented-reduce kernel (``ptgnn_b200.torch_scatter_shim``); the element-wise glue is device-side torch ops.
"""
import torch

from . import torch_scatter_shim as ts
from .messagepassing import AbstractMessageAggregation


class PnaMessageAggregation(AbstractMessageAggregation):
    def __init__(self, delta: float = 1):
        super().__init__()
        self._delta = delta

    def forward(self, messages: torch.Tensor, message_targets: torch.Tensor, num_nodes):
        degree = ts.scatter(torch.ones_like(message_targets), index=message_targets, dim_size=num_nodes, reduce="sum")
        dtype = messages.dtype
        m = messages.to(torch.float32)
        total = ts.scatter(m, index=message_targets, dim=0, dim_size=num_nodes, reduce="sum")
        mean = total / (degree.unsqueeze(-1) + 1e-5)
        largest = ts.scatter(m, index=message_targets, dim=0, dim_size=num_nodes, reduce="max")
        smallest = ts.scatter(m, index=message_targets, dim=0, dim_size=num_nodes, reduce="min")
        spread = torch.relu(m.pow(2) - mean[message_targets].pow(2)) + 1e-10
        std = torch.sqrt(ts.scatter(spread, index=message_targets, dim=0, dim_size=num_nodes, reduce="sum"))
        stats = torch.cat([total, mean, largest, smallest, std], dim=-1).to(dtype)
        amplify = torch.log(degree.float() + 1).unsqueeze(-1) / self._delta
        attenuate = 1 / (amplify + 1e-3)
        return torch.cat([stats, stats * amplify, stats * attenuate], dim=-1)

    def output_state_size(self, message_input_size: int) -> int:
        return message_input_size * 5 * 3
