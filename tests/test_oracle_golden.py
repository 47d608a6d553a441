"""Pins the CPU oracle (oracle/ptgnn_oracle.py and the plain-C oracle/oracle.c) against what the REFERENCE ITSELF
computed (tests/golden/*.npz, produced by tests/golden/generate_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from helpers import (GOLDEN_GATED, GOLDEN_MLP_KW, assert_close, gated_oracle_args, golden_adjacency, golden_state_dict,
                     load_golden, mlp_oracle_call_kwargs)
from oracle import coracle
from oracle import ptgnn_oracle as O


def _feats(adj):
    return [torch.empty(a[0].shape[0], 0) for a in adj]


@pytest.mark.parametrize("name", GOLDEN_GATED)
def test_gated_oracle_matches_reference(name):
    g = load_golden(name)
    adj, sd = golden_adjacency(g), golden_state_dict(g)
    h, agg = torch.from_numpy(g["h"]), str(g["agg"])
    out = O.gated_layer_forward(h, adj, _feats(adj), aggregation_fn=agg, **gated_oracle_args(sd))
    assert torch.equal(out, torch.from_numpy(g["out"])), "python oracle must reproduce the reference bit-exactly"
    a = gated_oracle_args(sd)
    c_out = coracle.gated_forward(g["h"], [(s.numpy(), t.numpy()) for s, t in adj], [w.numpy() for w in a["edge_weights"]],
                                  a["gru_w_ih"].numpy(), a["gru_w_hh"].numpy(), a["gru_b_ih"].numpy(), a["gru_b_hh"].numpy(), agg)
    assert_close(torch.from_numpy(c_out), torch.from_numpy(g["out"]), what=f"C oracle {name}")


@pytest.mark.parametrize("name", sorted(GOLDEN_MLP_KW))
def test_mlp_oracle_matches_reference(name):
    g = load_golden(name)
    adj, sd = golden_adjacency(g), golden_state_dict(g)
    h = torch.from_numpy(g["h"])
    kw = mlp_oracle_call_kwargs(name, sd)
    out = O.mlp_layer_forward(h, adj, _feats(adj), **kw)
    assert torch.equal(out, torch.from_numpy(g["out"])), "python oracle must reproduce the reference bit-exactly"
    c_out = coracle.mlp_forward(
        g["h"], [(s.numpy(), t.numpy()) for s, t in adj], [w[0].numpy() for w in kw["edge_mlp_weights"]], kw["aggregation_fn"],
        use_target=kw["use_target_state_as_message_input"], message_activation=kw["message_activation"],
        ln_weight=None if "ln_weight" not in kw else kw["ln_weight"].numpy(), ln_bias=None if "ln_bias" not in kw else kw["ln_bias"].numpy(),
        dense_weight=None if "dense_weight" not in kw else kw["dense_weight"].numpy(),
        dense_bias=None if "dense_bias" not in kw else kw["dense_bias"].numpy(), dense_activation=kw["dense_activation"])
    assert_close(torch.from_numpy(c_out), torch.from_numpy(g["out"]), what=f"C oracle {name}")


def test_container_oracle_matches_reference():
    g = load_golden("gnn_container")
    raw = golden_adjacency(g)
    n = g["h"].shape[0]
    adj = O.expand_adjacency(raw, n, True, True)
    assert len(adj) == int(g["num_expanded_types"])
    assert sum(a[0].shape[0] for a in adj) == int(g["num_edges"]) and n == int(g["num_nodes"])
    shared = dict(kind="gated", aggregation_fn="sum", **gated_oracle_args(golden_state_dict(g, "shared::")))
    last = dict(kind="gated", aggregation_fn="max", **gated_oracle_args(golden_state_dict(g, "last::")))
    states = O.gnn_forward(torch.from_numpy(g["h"]), adj, [shared, shared, last])
    assert torch.equal(states[-1], torch.from_numpy(g["out"]))
