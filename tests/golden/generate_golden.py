"""Generates the committed golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/generate_golden.py          # writes tests/golden/*.npz

The reference classes are imported read-only from /root/reference through ``oracle/refimport.py`` (which stubs the
two absent third-party wheels, see oracle/refstubs/).  Every fixture stores the seeded inputs, the reference
module's ``state_dict`` and the reference's outputs, so the GPU box -- which has no /root/reference -- can check
both the oracle and the CUDA path against what the reference itself computed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refimport import import_reference  # noqa: E402

import_reference()
from ptgnn.neuralmodels.gnn import GraphNeuralNetwork  # noqa: E402
from ptgnn.neuralmodels.gnn.messagepassing import GatedMessagePassingLayer, MlpMessagePassingLayer  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def random_graph(gen, n, counts):
    adj = []
    for c in counts:
        adj.append((torch.randint(0, n, (c,), generator=gen), torch.randint(0, n, (c,), generator=gen)))
    return adj


def pack(prefix, adj):
    d = {}
    for t, (s, g) in enumerate(adj):
        d[f"{prefix}src{t}"] = s.numpy()
        d[f"{prefix}tgt{t}"] = g.numpy()
    return d


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def state(module):
    return {"sd::" + k: v.detach().numpy() for k, v in module.state_dict().items()}


def run_layer(layer, h, adj):
    layer.eval()
    feats = [torch.empty(a[0].shape[0], 0) for a in adj]
    with torch.no_grad():
        return layer(node_states=h, adjacency_lists=adj, node_to_graph_idx=torch.zeros(h.shape[0], dtype=torch.int64),
                     reference_node_ids={}, reference_node_graph_idx={}, edge_features=feats)


def main():
    # ---- GatedMessagePassingLayer: sum / max / mean / min, empty type, isolated targets, hub target ----
    for i, (agg, n, H, counts) in enumerate(
        [("sum", 257, 32, [400, 0, 173, 31]), ("max", 257, 32, [400, 0, 173, 31]), ("mean", 130, 64, [300, 300]),
         ("min", 130, 64, [300, 300])]
    ):
        gen = torch.Generator().manual_seed(100 + i)
        torch.manual_seed(200 + i)
        adj = random_graph(gen, n - 20, counts)  # the last 20 nodes receive no edge (empty -> 0 rule)
        adj[0] = (adj[0][0], torch.cat([torch.full((60,), 7), adj[0][1][60:]]))  # hub: node 7 gets 60 type-0 edges
        h = torch.randn(n, H, generator=gen)
        layer = GatedMessagePassingLayer(H, H, len(counts), agg)
        out = run_layer(layer, h, adj)
        save(f"gated_{agg}", h=h.numpy(), out=out.numpy(), agg=agg, **pack("", adj), **state(layer))

    # ---- MlpMessagePassingLayer: default (GELU+LN+dense+tanh) max & sum; 2H -> H with D = 2H; no-target variant ----
    for i, (name, kw, n, counts) in enumerate(
        [
            ("mlp_max", dict(input_state_dimension=32, output_state_dimension=32, message_dimension=32, message_aggregation_function="max"), 200, [350, 120, 0]),
            ("mlp_sum", dict(input_state_dimension=32, output_state_dimension=32, message_dimension=32, message_aggregation_function="sum"), 200, [350, 120, 0]),
            ("mlp_wide", dict(input_state_dimension=64, output_state_dimension=32, message_dimension=64, message_aggregation_function="max"), 150, [280, 90]),
            ("mlp_notarget", dict(input_state_dimension=32, output_state_dimension=64, message_dimension=32, message_aggregation_function="mean",
                                  use_target_state_as_message_input=False), 150, [280, 90]),
            ("mlp_bare", dict(input_state_dimension=32, output_state_dimension=32, message_dimension=32, message_aggregation_function="min",
                              message_activation=None, use_layer_norm=False, use_dense_layer=False), 150, [280, 90]),
        ]
    ):
        gen = torch.Generator().manual_seed(300 + i)
        torch.manual_seed(400 + i)
        adj = random_graph(gen, n - 10, counts)
        h = torch.randn(n, kw["input_state_dimension"], generator=gen)
        layer = MlpMessagePassingLayer(num_edge_types=len(counts), **kw)
        # make LayerNorm / bias non-trivial
        with torch.no_grad():
            for k, v in layer.state_dict().items():
                if k.endswith("__state_update.0.weight") and v.dim() == 1:
                    v.copy_(1 + 0.1 * torch.randn(v.shape, generator=gen))
                if k.endswith(".bias"):
                    v.copy_(0.1 * torch.randn(v.shape, generator=gen))
        out = run_layer(layer, h, adj)
        save(name, h=h.numpy(), out=out.numpy(), **pack("", adj), **state(layer))

    # ---- GraphNeuralNetwork container: backward + self edges, 2 gated layers (one shared twice) + metrics ----
    gen = torch.Generator().manual_seed(500)
    torch.manual_seed(501)
    n, H = 180, 32
    raw = random_graph(gen, n, [260, 0, 75])
    T = 2 * len(raw) + 1
    shared = GatedMessagePassingLayer(H, H, T, "sum")
    last = GatedMessagePassingLayer(H, H, T, "max")

    class Embed(torch.nn.Module):
        def forward(self, x):
            return x

    gnn = GraphNeuralNetwork([shared, shared, last], Embed(), introduce_backwards_edges=True, add_self_edges=True)
    gnn.eval()
    h = torch.randn(n, H, generator=gen)
    adj_arg = list(raw)  # the reference mutates this list in place
    with torch.no_grad():
        res = gnn(node_data={"x": h}, adjacency_lists=adj_arg, edge_feature_data=[], node_to_graph_idx=torch.zeros(n, dtype=torch.int64),
                  reference_node_ids={}, reference_node_graph_idx={}, num_graphs=3)
    metrics = gnn.report_metrics()
    save("gnn_container", h=h.numpy(), out=res.output_node_representations.numpy(), num_expanded_types=len(adj_arg),
         num_graphs=metrics["num_graphs"], num_nodes=metrics["num_nodes"], num_edges=metrics["num_edges"], **pack("", raw),
         **{"shared::" + k: v.numpy() for k, v in shared.state_dict().items()},
         **{"last::" + k: v.numpy() for k, v in last.state_dict().items()})


def main_round2():
    """Round-2 fixtures: (1) the reference's OWN bf16 path -- fp32 modules under torch.autocast("cpu", bfloat16), the AMP
    arithmetic of trainer.py:221 -- next to its fp32 output on the same inputs (the gap between the two is what any bf16
    bar has to be read against); (2) a Typilus-GGNN-style stack with the reference's ConcatResidualLayer between gated layers
    (typilus/train.py:39-65) through the reference container."""
    from ptgnn.neuralmodels.gnn.messagepassing.residuallayers import ConcatResidualLayer

    for i, (kind, agg) in enumerate([("gated", "sum"), ("gated", "max"), ("mlp", "sum"), ("mlp", "max")]):
        gen = torch.Generator().manual_seed(700 + i)
        torch.manual_seed(800 + i)
        n, H, counts = 384, 128, [900, 0, 500, 260]
        adj = random_graph(gen, n - 16, counts)
        h = torch.randn(n, H, generator=gen)
        if kind == "gated":
            layer = GatedMessagePassingLayer(H, H, len(counts), agg)
        else:
            layer = MlpMessagePassingLayer(H, H, H, len(counts), agg)
        out32 = run_layer(layer, h, adj)
        hb = h.to(torch.bfloat16)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out_ac = run_layer(layer, hb, adj)
        out32_on_rounded = run_layer(layer, hb.float(), adj)
        gap = (out_ac.float() - out32_on_rounded).abs()
        print(f"{kind}_{agg}_bf16ac: autocast output dtype {out_ac.dtype}; autocast vs fp32 (same rounded inputs): max {gap.max():.3e} "
              f"mean {gap.mean():.3e} within 1e-2: {(gap <= 1e-2 * out32_on_rounded.abs().clamp(min=1)).float().mean():.4f}")
        save(f"{kind}_{agg}_bf16ac", h=h.numpy(), out_autocast=out_ac.float().numpy(), out_fp32=out32.numpy(),
             out_fp32_rounded_inputs=out32_on_rounded.numpy(), agg=agg, **pack("", adj), **state(layer))

    gen = torch.Generator().manual_seed(900)
    torch.manual_seed(901)
    n, H = 300, 64
    raw = random_graph(gen, n, [500, 0, 140])
    T = 2 * len(raw) + 1
    shared = GatedMessagePassingLayer(H, H, T, "max")
    last = GatedMessagePassingLayer(2 * H, H, T, "max")
    r1 = ConcatResidualLayer(H)

    class Embed(torch.nn.Module):
        def forward(self, x):
            return x

    gnn = GraphNeuralNetwork([r1.pass_through_dummy_layer(), shared, shared, shared, r1, last], Embed(),
                             introduce_backwards_edges=True, add_self_edges=True)
    gnn.eval()
    h = torch.randn(n, H, generator=gen)
    with torch.no_grad():
        res = gnn(node_data={"x": h}, adjacency_lists=list(raw), edge_feature_data=[], node_to_graph_idx=torch.zeros(n, dtype=torch.int64),
                  reference_node_ids={}, reference_node_graph_idx={}, num_graphs=2)
    save("gnn_residual", h=h.numpy(), out=res.output_node_representations.numpy(), **pack("", raw),
         **{"shared::" + k: v.numpy() for k, v in shared.state_dict().items()},
         **{"last::" + k: v.numpy() for k, v in last.state_dict().items()})


def main_minibatch():
    """Runs the UNMODIFIED GraphNeuralNetworkModel.initialize_minibatch / extend_minibatch_with / finalize_minibatch
    (graphneuralnetwork.py:372-493) on synthetic graphs.  The methods are called on a stand-in `self` that carries exactly the
    attributes they read (stub node embedder, no edge embedder): the model's constructor needs embedder models and metadata that are
    out of this path's scope, the three methods do not."""
    import types

    from ptgnn.neuralmodels.gnn.graphneuralnetwork import GraphNeuralNetworkModel
    from ptgnn.neuralmodels.gnn.structs import TensorizedGraphData

    class _NodeEmbedderStub:
        def initialize_minibatch(self):
            return {}

        def extend_minibatch_with(self, item, mb):
            return True

        def finalize_minibatch(self, mb, device):
            return {}

    num_types = 4
    stub = types.SimpleNamespace(stop_extending_minibatch_after_num_nodes=10 ** 9)
    setattr(stub, "_GraphNeuralNetworkModel__node_embedding_model", _NodeEmbedderStub())
    setattr(stub, "_GraphNeuralNetworkModel__edge_embedding_model", None)
    setattr(stub, "_GraphNeuralNetworkModel__edge_types", {f"e{t}": t for t in range(num_types)})
    setattr(stub, "_GraphNeuralNetworkModel__create_node_to_graph_idx",
            getattr(GraphNeuralNetworkModel, "_GraphNeuralNetworkModel__create_node_to_graph_idx"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import minibatch_graphs

    graphs = minibatch_graphs(num_types=num_types)
    mb = GraphNeuralNetworkModel.initialize_minibatch(stub)
    for adj, refs, n in graphs:
        GraphNeuralNetworkModel.extend_minibatch_with(
            stub, TensorizedGraphData(adjacency_lists=adj, node_tensorized_data=[], edge_features=None, reference_nodes=refs, num_nodes=n), mb)
    out = GraphNeuralNetworkModel.finalize_minibatch(stub, mb, "cpu")
    arrays = {"num_graphs": out["num_graphs"], "node_to_graph_idx": out["node_to_graph_idx"].numpy()}
    for t, (s, g) in enumerate(out["adjacency_lists"]):
        arrays[f"src{t}"], arrays[f"tgt{t}"] = s.numpy(), g.numpy()
    for k, v in out["reference_node_ids"].items():
        arrays[f"ref_ids::{k}"] = v.numpy()
    for k, v in out["reference_node_graph_idx"].items():
        arrays[f"ref_graph::{k}"] = v.numpy()
    save("minibatch", **arrays)


def main_egc():
    """EGCMessagePassingLayer (egcmessagepassing.py) run unmodified: sum and max, 3 edge types (one empty), H = 64 -> 64."""
    from ptgnn.neuralmodels.gnn.messagepassing.egcmessagepassing import EGCMessagePassingLayer

    for agg in ("sum", "max"):
        gen = torch.Generator().manual_seed(31)
        torch.manual_seed(31)
        n, counts = 300, [900, 0, 250]
        adj = random_graph(gen, n, counts)
        h = torch.randn(n, 64, generator=gen)
        layer = EGCMessagePassingLayer(64, 64, len(counts), agg, num_bases=4, num_heads=8)
        out = run_layer(layer, h, adj)
        save(f"egc_{agg}", h=h.numpy(), out=out.numpy(), **pack("", adj), **state(layer))


if __name__ == "__main__":
    if "--minibatch-only" in sys.argv:
        main_minibatch()
        sys.exit(0)
    if "--egc-only" in sys.argv:
        main_egc()
        sys.exit(0)
    if "--round2-only" not in sys.argv:
        main()
    main_round2()
    main_minibatch()
    main_egc()
