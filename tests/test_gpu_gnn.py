"""GraphNeuralNetwork container on the GPU: reference golden run, PPI-shaped config 1, and full-size properties."""
import pytest
import torch

from helpers import assert_close, gated_oracle_args, golden_adjacency, golden_state_dict, load_golden
from oracle import ptgnn_oracle as O

pytestmark = pytest.mark.gpu


class _Embed(torch.nn.Module):
    def forward(self, x):
        return x


def _forward(gnn, h, raw, num_graphs=1, node_to_graph=None):
    gnn = gnn.cuda().eval()
    n = h.shape[0]
    n2g = (torch.zeros(n, dtype=torch.int64) if node_to_graph is None else node_to_graph).cuda()
    adj = [(s.cuda(), t.cuda()) for s, t in raw]
    with torch.no_grad():
        out = gnn(node_data={"x": h.cuda()}, adjacency_lists=adj, edge_feature_data=[], node_to_graph_idx=n2g,
                  reference_node_ids={}, reference_node_graph_idx={}, num_graphs=num_graphs)
    assert len(adj) == len(raw), "the caller's adjacency list must not be mutated"
    return out


def test_container_vs_reference_golden():
    import ptgnn_b200 as P

    g = load_golden("gnn_container")
    raw = golden_adjacency(g)
    T = int(g["num_expanded_types"])
    shared, last = P.GatedMessagePassingLayer(32, 32, T, "sum"), P.GatedMessagePassingLayer(32, 32, T, "max")
    shared.load_state_dict(golden_state_dict(g, "shared::"))
    last.load_state_dict(golden_state_dict(g, "last::"))
    gnn = P.GraphNeuralNetwork([shared, shared, last], _Embed(), introduce_backwards_edges=True, add_self_edges=True)
    out = _forward(gnn, torch.from_numpy(g["h"]), raw, num_graphs=3)
    assert_close(out.output_node_representations.cpu(), torch.from_numpy(g["out"]), what="container")
    # integer bookkeeping is bit-exact with the reference's metrics
    assert gnn.report_metrics() == {"num_graphs": int(g["num_graphs"]), "num_nodes": int(g["num_nodes"]),
                                    "num_edges": int(g["num_edges"])}
    assert out.num_graphs == 3 and torch.equal(out.input_node_representations.cpu(), torch.from_numpy(g["h"]))


def test_config1_ppi_shaped_batch():
    """BASELINE.json configs[0]: PPI-shaped minibatch, 1 GGNN layer, hidden 64, backward + self edges (T=3, E=93,000)."""
    import ptgnn_b200 as P
    from ptgnn_b200.synthetic import ppi_batch

    b = ppi_batch(seed=0)
    torch.manual_seed(0)
    layer = P.GatedMessagePassingLayer(64, 64, 3, "sum")
    gnn = P.GraphNeuralNetwork([layer], _Embed(), introduce_backwards_edges=True, add_self_edges=True)
    h = torch.randn(b.num_nodes, 64, generator=torch.Generator().manual_seed(0))
    out = _forward(gnn, h, b.adjacency_lists, num_graphs=b.num_graphs, node_to_graph=b.node_to_graph_idx)
    adj = O.expand_adjacency(b.adjacency_lists, b.num_nodes, True, True)
    sd = {k: v.cpu() for k, v in layer.state_dict().items()}
    ref = O.gnn_forward(h, adj, [dict(kind="gated", aggregation_fn="sum", **gated_oracle_args(sd))])[-1]
    assert_close(out.output_node_representations.cpu(), ref, what="config 1")
    assert gnn.report_metrics() == {"num_graphs": 2, "num_nodes": 3000, "num_edges": 93000}


@pytest.mark.parametrize("agg", ["sum", "max"])
def test_config2_full_size_properties(agg):
    """BASELINE.json configs[1] at full size (N=204,800, E=1,105,920, T=17, H=128): size-independent properties --
    (a) a random subset of target rows equals the oracle run on just the edges into those rows,
    (b) permuting edges inside each type leaves `max` bit-identical and `sum` within tolerance,
    (c) run-to-run determinism."""
    import ptgnn_b200 as P
    from ptgnn_b200.synthetic import graph2class_batch

    b = graph2class_batch()
    assert b.num_nodes == 204800 and b.layer_level_edges() == 1105920
    torch.manual_seed(0)
    layer = P.GatedMessagePassingLayer(128, 128, 17, agg)
    gnn = P.GraphNeuralNetwork([layer], _Embed(), True, True)
    gen = torch.Generator().manual_seed(11)
    h = torch.randn(b.num_nodes, 128, generator=gen)
    out = _forward(gnn, h, b.adjacency_lists, b.num_graphs, b.node_to_graph_idx).output_node_representations.cpu()
    out2 = _forward(gnn, h, b.adjacency_lists, b.num_graphs, b.node_to_graph_idx).output_node_representations.cpu()
    assert torch.equal(out, out2)

    # (a) sampled rows vs oracle
    adj = O.expand_adjacency(b.adjacency_lists, b.num_nodes, True, True)
    rows = torch.randperm(b.num_nodes, generator=gen)[:512]
    keep = torch.zeros(b.num_nodes, dtype=torch.bool)
    keep[rows] = True
    sub = [(s[keep[t]], t[keep[t]]) for s, t in adj]
    sd = {k: v.cpu() for k, v in layer.state_dict().items()}
    ref = O.gated_layer_forward(h, sub, [torch.empty(a[0].shape[0], 0) for a in sub], aggregation_fn=agg,
                                **gated_oracle_args(sd))
    assert_close(out[rows], ref[rows], what=f"config 2 sampled rows ({agg})")

    # (b) edge order inside a type is irrelevant
    shuffled = []
    for s, t in b.adjacency_lists:
        p = torch.randperm(s.shape[0], generator=gen)
        shuffled.append((s[p], t[p]))
    out3 = _forward(gnn, h, shuffled, b.num_graphs, b.node_to_graph_idx).output_node_representations.cpu()
    if agg == "max":
        assert torch.equal(out, out3)
    else:
        assert_close(out3, out, what="edge-order invariance")
