"""CUDA layers (called through the C ABI) vs the reference's own outputs (golden fixtures) and vs the oracle."""
import pytest
import torch

from helpers import (GOLDEN_GATED, GOLDEN_MLP_KW, assert_close, gated_oracle_args, golden_adjacency, golden_state_dict,
                     load_golden, mlp_oracle_call_kwargs, random_adjacency)
from oracle import ptgnn_oracle as O

pytestmark = pytest.mark.gpu


def _run(layer, h, adj):
    layer = layer.cuda().eval()
    with torch.no_grad():
        return layer(node_states=h.cuda(), adjacency_lists=[(s.cuda(), t.cuda()) for s, t in adj],
                     node_to_graph_idx=torch.zeros(h.shape[0], dtype=torch.int64).cuda(), reference_node_ids={},
                     reference_node_graph_idx={}, edge_features=[torch.empty(a[0].shape[0], 0).cuda() for a in adj]).cpu()


@pytest.mark.parametrize("name", GOLDEN_GATED)
def test_gated_vs_reference_golden(name):
    import ptgnn_b200 as P

    g = load_golden(name)
    adj, sd = golden_adjacency(g), golden_state_dict(g)
    H = g["h"].shape[1]
    layer = P.GatedMessagePassingLayer(H, H, len(adj), str(g["agg"]))
    layer.load_state_dict(sd)
    assert_close(_run(layer, torch.from_numpy(g["h"]), adj), torch.from_numpy(g["out"]), what=name)


@pytest.mark.parametrize("name", sorted(GOLDEN_MLP_KW))
def test_mlp_vs_reference_golden(name):
    import ptgnn_b200 as P

    g = load_golden(name)
    adj, sd = golden_adjacency(g), golden_state_dict(g)
    layer = P.MlpMessagePassingLayer(num_edge_types=len(adj), **GOLDEN_MLP_KW[name])
    layer.load_state_dict(sd)
    assert_close(_run(layer, torch.from_numpy(g["h"]), adj), torch.from_numpy(g["out"]), what=name)


@pytest.mark.parametrize("agg", ["sum", "max"])
@pytest.mark.parametrize("n,H,counts", [
    (1000, 64, [3000, 1500, 0, 700]),
    (3001, 128, [9000, 9000, 5000, 1, 130, 0, 2000]),
    (700, 256, [4000, 300]),
    (129, 32, [127]),
])
def test_gated_vs_oracle_random(agg, n, H, counts):
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(n + H)
    torch.manual_seed(n)
    adj = random_adjacency(gen, n, counts)
    h = torch.randn(n, H, generator=gen)
    layer = P.GatedMessagePassingLayer(H, H, len(counts), agg)
    ref = O.gated_layer_forward(h, adj, [torch.empty(c, 0) for c in counts], aggregation_fn=agg,
                                **gated_oracle_args({k: v.clone() for k, v in layer.state_dict().items()}))
    assert_close(_run(layer, h, adj), ref, what=f"gated {agg} N={n} H={H}")


@pytest.mark.parametrize("agg", ["sum", "max", "mean"])
@pytest.mark.parametrize("n,Hin,D,Hout,counts", [
    (1000, 64, 64, 64, [3000, 1500, 0, 700]),
    (2500, 128, 128, 128, [9000, 9000, 5000, 1]),
    (900, 128, 128, 64, [4000, 300]),      # Typilus-style 2H -> H layer with D = 2H
    (400, 32, 96, 160, [1500]),             # odd tile shapes: D, Hout not multiples of 64
])
def test_mlp_vs_oracle_random(agg, n, Hin, D, Hout, counts):
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(n + D)
    torch.manual_seed(n + 1)
    adj = random_adjacency(gen, n, counts)
    h = torch.randn(n, Hin, generator=gen)
    layer = P.MlpMessagePassingLayer(Hin, Hout, D, len(counts), agg)
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    p = "_MlpMessagePassingLayer__"
    ref = O.mlp_layer_forward(
        h, adj, [torch.empty(c, 0) for c in counts],
        [[sd[f"{p}edge_message_transformation_layers.{t}._MLP__mlp_modules.1.weight"]] for t in range(len(counts))], agg,
        ln_weight=sd[p + "state_update.0.weight"], ln_bias=sd[p + "state_update.0.bias"],
        dense_weight=sd[p + "state_update.1.weight"], dense_bias=sd[p + "state_update.1.bias"])
    assert_close(_run(layer, h, adj), ref, what=f"mlp {agg} N={n}")


def test_layer_is_deterministic_and_does_not_modify_input():
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(4)
    adj = random_adjacency(gen, 2000, [8000, 8000])
    h = torch.randn(2000, 128, generator=gen)
    layer = P.GatedMessagePassingLayer(128, 128, 2, "sum").cuda().eval()
    hc = h.cuda()
    adj_c = [(s.cuda(), t.cuda()) for s, t in adj]
    with torch.no_grad():
        a = layer(hc, adj_c)
        b = layer(hc, adj_c)
    assert torch.equal(a, b) and torch.equal(hc.cpu(), h) and a.data_ptr() != hc.data_ptr()


def test_layers_with_no_edges_at_all():
    """Every edge type empty: aggregation is all zeros (torch_scatter: empty -> 0), the node update still runs."""
    import ptgnn_b200 as P

    torch.manual_seed(3)
    n, H = 300, 64
    h = torch.randn(n, H, generator=torch.Generator().manual_seed(1))
    adj = [(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)) for _ in range(2)]
    gated = P.GatedMessagePassingLayer(H, H, 2, "max")
    ref = O.gated_layer_forward(h, adj, [torch.empty(0, 0)] * 2, aggregation_fn="max",
                                **gated_oracle_args({k: v.clone() for k, v in gated.state_dict().items()}))
    assert_close(_run(gated, h, adj), ref, what="gated, E = 0")
    mlp = P.MlpMessagePassingLayer(H, H, H, 2, "sum")
    sd = {k: v.clone() for k, v in mlp.state_dict().items()}
    p = "_MlpMessagePassingLayer__"
    ref = O.mlp_layer_forward(h, adj, [torch.empty(0, 0)] * 2,
                              [[sd[f"{p}edge_message_transformation_layers.{t}._MLP__mlp_modules.1.weight"]] for t in range(2)], "sum",
                              ln_weight=sd[p + "state_update.0.weight"], ln_bias=sd[p + "state_update.0.bias"],
                              dense_weight=sd[p + "state_update.1.weight"], dense_bias=sd[p + "state_update.1.bias"])
    assert_close(_run(mlp, h, adj), ref, what="mlp, E = 0")


def test_single_node_and_single_edge():
    import ptgnn_b200 as P

    torch.manual_seed(4)
    H = 32
    h = torch.randn(1, H)
    adj = [(torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64))]
    layer = P.GatedMessagePassingLayer(H, H, 1, "sum")
    ref = O.gated_layer_forward(h, adj, [torch.empty(1, 0)], aggregation_fn="sum",
                                **gated_oracle_args({k: v.clone() for k, v in layer.state_dict().items()}))
    assert_close(_run(layer, h, adj), ref, what="N = 1")
