"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares, the
module API mirrors the reference's, and the product refuses to run without CUDA (no CPU fallback)."""
import inspect
import os
import re

import pytest
import torch

import ptgnn_b200 as P
from ptgnn_b200 import _native as N
from helpers import GOLDEN_MLP_KW, golden_state_dict, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ptgnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptgnn_b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 12
    handle = N.lib()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/ptgnn_b200.h but not exported"
    assert sorted(N.SIGNATURES) == declared, "ctypes SIGNATURES must cover exactly the header's entry points"
    assert handle.ptgnn_b200_abi_version() == 2


def test_workspace_size_queries_run_without_a_gpu():
    handle = N.lib()
    assert handle.ptgnn_b200_plan_workspace_bytes(1000, 5000) > 4 * 5000 * 4
    assert handle.ptgnn_b200_gated_workspace_bytes(1000, 5000, 17, 128, 128) >= 5000 * 128 * 4 + 1000 * 128 * 4
    assert handle.ptgnn_b200_mlp_workspace_bytes(1000, 5000, 17, 128, 128, 128, 1) >= 5000 * 128 * 4
    assert handle.ptgnn_b200_scatter_workspace_bytes(1000, 5000) > handle.ptgnn_b200_plan_workspace_bytes(1000, 5000)


def test_no_cpu_fallback():
    layer = P.GatedMessagePassingLayer(32, 32, 1, "sum").eval()
    adj = [(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64))]
    with torch.no_grad(), pytest.raises(N.NativeLibraryError):
        layer(node_states=torch.zeros(4, 32), adjacency_lists=adj, node_to_graph_idx=None, reference_node_ids={},
              reference_node_graph_idx={}, edge_features=[torch.empty(3, 0)])
    with pytest.raises(N.NativeLibraryError):
        P.scatter(torch.zeros(3, 4), torch.zeros(3, dtype=torch.int64), dim=0, dim_size=2, reduce="sum")


def test_constructor_signatures_match_reference_contract():
    gated = list(inspect.signature(P.GatedMessagePassingLayer.__init__).parameters)
    assert gated == ["self", "state_dimension", "message_dimension", "num_edge_types", "message_aggregation_function",
                     "dropout_rate", "edge_feature_dimension"]
    mlp = list(inspect.signature(P.MlpMessagePassingLayer.__init__).parameters)
    assert mlp == ["self", "input_state_dimension", "output_state_dimension", "message_dimension", "num_edge_types",
                   "message_aggregation_function", "message_activation", "use_target_state_as_message_input",
                   "mlp_hidden_layers", "use_layer_norm", "use_dense_layer", "dropout_rate", "dense_activation",
                   "features_dimension"]
    fwd = list(inspect.signature(P.AbstractMessagePassingLayer.forward).parameters)
    assert fwd == ["self", "node_states", "adjacency_lists", "node_to_graph_idx", "reference_node_ids",
                   "reference_node_graph_idx", "edge_features"]
    gnn = list(inspect.signature(P.GraphNeuralNetwork.__init__).parameters)
    assert gnn == ["self", "message_passing_layers", "node_embedder", "introduce_backwards_edges", "add_self_edges",
                   "edge_dropout_rate", "edge_feature_embedder"]


def test_reference_checkpoints_load():
    g = load_golden("gated_sum")
    layer = P.GatedMessagePassingLayer(32, 32, 4, "sum")
    layer.load_state_dict(golden_state_dict(g), strict=True)
    assert layer.input_state_dimension == 32 and layer.output_state_dimension == 32
    for name, kw in GOLDEN_MLP_KW.items():
        sd = golden_state_dict(load_golden(name))
        T = sum(1 for k in sd if k.endswith("_MLP__mlp_modules.1.weight"))
        P.MlpMessagePassingLayer(num_edge_types=T, **kw).load_state_dict(sd, strict=True)


def test_unsupported_configurations_fail_loudly():
    adj = [(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64))]
    # training (tests/test_gpu_backward.py) exists for both layer classes -- on CUDA tensors only, fp32 only
    with pytest.raises(N.NativeLibraryError):
        P.MlpMessagePassingLayer(32, 32, 32, 1, "sum")(torch.zeros(4, 32), adj)
    with pytest.raises(N.NativeLibraryError):
        P.GatedMessagePassingLayer(32, 32, 1, "sum")(torch.zeros(4, 32), adj)
    with pytest.raises(NotImplementedError):  # ... not for bf16 states
        P.GatedMessagePassingLayer(32, 32, 1, "sum")(torch.zeros(4, 32, dtype=torch.bfloat16), adj)
    with pytest.raises(NotImplementedError):  # ... nor for message MLPs with hidden layers
        P.MlpMessagePassingLayer(32, 32, 32, 1, "sum", mlp_hidden_layers=1)(torch.zeros(4, 32), adj)
    with torch.no_grad():
        with pytest.raises(N.NativeLibraryError):  # training-mode dropout runs (per-edge mask on the gathered rows) -- on CUDA tensors
            P.GatedMessagePassingLayer(32, 32, 1, "sum", dropout_rate=0.5).train()(torch.zeros(4, 32), adj)
        with pytest.raises(N.NativeLibraryError):  # hidden MLP layers are supported (composed path) -- on CUDA tensors only
            P.MlpMessagePassingLayer(32, 32, 32, 1, "sum", mlp_hidden_layers=1).eval()(torch.zeros(4, 32), adj)
        with pytest.raises(NotImplementedError):  # unknown aggregation
            P.GatedMessagePassingLayer(32, 32, 1, "median").eval()(torch.zeros(4, 32), adj)
        with pytest.raises(ValueError):  # width mismatch is a shape error, not a read past the weight buffers
            P.GatedMessagePassingLayer(32, 32, 1, "sum").eval()(torch.zeros(4, 64), adj)
        with pytest.raises(ValueError):
            P.MlpMessagePassingLayer(32, 32, 32, 1, "sum").eval()(torch.zeros(4, 32, 1), adj)


def test_container_metrics_protocol():
    layer = P.GatedMessagePassingLayer(32, 32, 3, "sum")
    gnn = P.GraphNeuralNetwork([layer, layer], torch.nn.Identity(), True, True)
    assert gnn.report_metrics() == {"num_graphs": 0, "num_nodes": 0, "num_edges": 0}
    assert gnn.input_node_state_dim == 32 and gnn.output_node_state_dim == 32
    assert len(gnn.message_passing_layers) == 2
    # weight sharing: the same module registered twice contributes its parameters once
    assert len(list(gnn.parameters())) == len(list(layer.parameters()))
    raw = [(torch.tensor([0, 1]), torch.tensor([1, 2]))]
    expanded = gnn.expand_adjacency(raw, 4, torch.device("cpu"))
    assert len(raw) == 1 and len(expanded) == 3  # caller's list is NOT mutated
    assert torch.equal(expanded[1][0], raw[0][1]) and torch.equal(expanded[2][0], torch.arange(4))


def test_weight_cache_bookkeeping_without_gpu(monkeypatch):
    """The derived-weight cache is keyed on (data_ptr, version) of every parameter: reuse only while nothing changed, never
    in training mode, never carried through deepcopy / pickling.  (The kernels behind it: tests/test_gpu_weight_cache.py.)"""
    import copy

    import ptgnn_b200 as P

    class _Stream:
        cuda_stream = 0

    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    layer = P.GatedMessagePassingLayer(64, 64, 2, "sum").eval()
    params = list(layer.parameters())
    dev = torch.device("cpu")
    buf, valid = layer._weight_cache("f32", 1024, params, dev)
    assert buf is not None and not valid
    layer._weight_cache_filled("f32", dev)
    buf2, valid = layer._weight_cache("f32", 1024, params, dev)
    assert valid and buf2 is buf
    layer._weight_cache_filled("f32", dev)
    with torch.no_grad():
        params[0].mul_(2.0)                                   # what load_state_dict / an optimiser step does
    assert layer._weight_cache("f32", 1024, params, dev)[1] is False
    layer._weight_cache_filled("f32", dev)
    assert layer._weight_cache("f32", 1024, params, dev)[1] is True
    layer._weight_cache_filled("f32", dev)
    layer.train()
    assert layer._weight_cache("f32", 1024, params, dev)[1] is False       # never trusted in training mode
    layer.eval()
    assert layer._weight_cache("f32", 0, params, dev) == (None, False)     # nothing to cache for these dims
    assert copy.deepcopy(layer)._derived_weights == {}
    assert not any("derived" in k for k in layer.state_dict())
    layer.invalidate_weight_cache()
    assert layer._weight_cache("f32", 1024, params, dev)[1] is False


def test_bench_clock_summary_decodes_nvml_reason_bits():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    assert s.summary()["sm_mhz"] is None
    s.sm_max, s.source = 1965, "nvml"
    s.samples = [(1965, 300.0, 0x0), (1950, 320.0, 0x4), (1965, 310.0, 0x1), (1700, 330.0, 0x40)]
    out = s.summary()
    assert out["sm_mhz"] == 1965 and out["sm_max_mhz"] == 1965 and out["samples"] == 4
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"] and out["power_w_max"] == 330.0


def test_state_chain_bookkeeping_without_gpu(monkeypatch):
    """edgeplan.state_chain: per-thread slot keyed on tensor identity + version; nests; PTGNN_B200_CHAIN=0 hides it; other threads
    never see it.  (What the layers do with it: tests/test_gpu_round2.py::test_state_chain_*.)"""
    import threading

    from ptgnn_b200 import edgeplan as EP

    assert EP.current_state_chain() is None
    out, packed = torch.zeros(4, 8), torch.zeros(4, 32, dtype=torch.uint8)
    with EP.state_chain() as chain:
        assert EP.current_state_chain() is chain and chain.lookup(out) is None
        chain.store(out, packed)
        assert chain.lookup(out) is packed
        assert chain.lookup(out.clone()) is None                     # identity, not equality
        seen = []
        t = threading.Thread(target=lambda: seen.append(EP.current_state_chain()))
        t.start(); t.join()
        assert seen == [None]                                        # per thread
        with EP.state_chain() as inner:
            assert EP.current_state_chain() is inner and inner.lookup(out) is None
        assert EP.current_state_chain() is chain
        out.add_(1.0)                                                # in-place edit: the packed copy is stale
        assert chain.lookup(out) is None
        chain.store(out, None)                                       # a layer that produced no packed form clears the slot
        assert chain.lookup(out) is None
        monkeypatch.setenv("PTGNN_B200_CHAIN", "0")
        assert EP.current_state_chain() is None
        monkeypatch.delenv("PTGNN_B200_CHAIN")
    assert EP.current_state_chain() is None


def test_minibatch_assembler_is_exported_and_egc_has_the_reference_signature():
    egc = list(inspect.signature(P.EGCMessagePassingLayer.__init__).parameters)
    assert egc == ["self", "input_state_dimension", "output_state_dimension", "num_edge_types", "message_aggregation_function",
                   "num_bases", "num_heads", "dropout_rate"]
    asm = P.MinibatchAssembler(3, stop_extending_minibatch_after_num_nodes=7)
    mb = asm.initialize_minibatch()
    assert sorted(mb) == ["adjacency_lists", "num_nodes_in_mb", "num_nodes_per_graph", "reference_node_ids"] and len(mb["adjacency_lists"]) == 3
