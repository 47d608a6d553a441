"""Device-side minibatch finalisation through the C ABI: bit-exact against the reference's own output
(tests/golden/minibatch.npz) and against the oracle on a config-2-sized batch; the result feeds the layer loop."""
import types

import numpy as np
import pytest
import torch

from helpers import minibatch_graphs
from oracle import batching_oracle as B
from test_oracle_batching import golden_minibatch

pytestmark = pytest.mark.gpu


def _ns(adj, refs, n):
    return types.SimpleNamespace(adjacency_lists=adj, reference_nodes=refs, num_nodes=n)


def _same(out, ref):
    assert out["num_graphs"] == ref["num_graphs"]
    assert torch.equal(out["node_to_graph_idx"].cpu(), torch.from_numpy(ref["node_to_graph_idx"]))
    assert len(out["adjacency_lists"]) == len(ref["adjacency_lists"])
    for (s, t), (rs, rt) in zip(out["adjacency_lists"], ref["adjacency_lists"]):
        assert s.dtype == torch.int64 and s.is_cuda
        assert torch.equal(s.cpu(), torch.from_numpy(rs)) and torch.equal(t.cpu(), torch.from_numpy(rt))
    assert sorted(out["reference_node_ids"]) == sorted(ref["reference_node_ids"])
    for k in ref["reference_node_ids"]:
        assert torch.equal(out["reference_node_ids"][k].cpu(), torch.from_numpy(ref["reference_node_ids"][k]))
        assert torch.equal(out["reference_node_graph_idx"][k].cpu(), torch.from_numpy(ref["reference_node_graph_idx"][k]))


def test_finalize_matches_the_reference_golden():
    import ptgnn_b200 as P

    g = golden_minibatch()
    asm = P.MinibatchAssembler(4, 10 ** 9)
    mb = asm.initialize_minibatch()
    for adj, refs, n in minibatch_graphs(num_types=4):
        assert asm.extend_minibatch_with(_ns(adj, refs, n), mb)
    out = asm.finalize_minibatch(mb, "cuda")
    ref = {"num_graphs": int(g["num_graphs"]), "node_to_graph_idx": g["node_to_graph_idx"],
           "adjacency_lists": [(g[f"src{t}"], g[f"tgt{t}"]) for t in range(4)],
           "reference_node_ids": {k.split("::", 1)[1]: g[k] for k in g.files if k.startswith("ref_ids::")},
           "reference_node_graph_idx": {k.split("::", 1)[1]: g[k] for k in g.files if k.startswith("ref_graph::")}}
    _same(out, ref)


@pytest.mark.parametrize("num_graphs,nodes", [(1, 7), (300, 700), (6000, 30)])
def test_finalize_vs_oracle_large_and_ragged(num_graphs, nodes):
    """Up to ~200k nodes / 450k edges (BASELINE config 2's batch), more graphs than the kernels' shared-memory pointer cache (4096),
    empty graphs' edge lists, a graph with zero reference nodes."""
    import ptgnn_b200 as P

    rng = np.random.RandomState(num_graphs)
    T = 8
    asm = P.MinibatchAssembler(T, 10 ** 9)
    mb, ref_mb = asm.initialize_minibatch(), B.initialize_minibatch(T)
    for g in range(num_graphs):
        n = int(rng.randint(1, 2 * nodes))
        adj = []
        for t in range(T):
            e = 0 if rng.rand() < 0.2 else int(rng.randint(0, n // 2 + 2))
            adj.append((rng.randint(0, n, e).astype(np.int32), rng.randint(0, n, e).astype(np.int32)))
        refs = {"a": rng.randint(0, n, int(rng.randint(0, 3))).astype(np.int32)}
        graph = _ns(adj, refs, n)
        asm.extend_minibatch_with(graph, mb)
        B.extend_minibatch_with(graph, ref_mb, 10 ** 9)
    out = asm.finalize_minibatch(mb, "cuda")
    torch.cuda.synchronize()
    _same(out, B.finalize_minibatch(ref_mb))


def test_assembled_minibatch_runs_through_the_layer_loop():
    import ptgnn_b200 as P

    asm = P.MinibatchAssembler(4, 10 ** 9)
    mb = asm.initialize_minibatch()
    for adj, refs, n in minibatch_graphs(num_types=4):
        asm.extend_minibatch_with(_ns(adj, refs, n), mb)
    out = asm.finalize_minibatch(mb, "cuda")
    torch.manual_seed(0)
    H, T = 64, 2 * 4 + 1
    class _Embed(torch.nn.Module):
        def forward(self, x):
            return x

    gnn = P.GraphNeuralNetwork([P.GatedMessagePassingLayer(H, H, T, "sum")], _Embed(), True, True).cuda().eval()
    n = int(out["node_to_graph_idx"].shape[0])
    with torch.no_grad():
        res = gnn(node_data={"x": torch.randn(n, H).cuda()}, adjacency_lists=out["adjacency_lists"], edge_feature_data=[],
                  node_to_graph_idx=out["node_to_graph_idx"], reference_node_ids=out["reference_node_ids"],
                  reference_node_graph_idx=out["reference_node_graph_idx"], num_graphs=out["num_graphs"])
    assert res.output_node_representations.shape == (n, H) and res.num_graphs == out["num_graphs"]
    assert torch.isfinite(res.output_node_representations).all()
