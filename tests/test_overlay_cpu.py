"""`ptgnn_b200.overlay.install()`: the reference's own packages build ptgnn_b200 layers WITHOUT being edited.

Needs the reference tree (PTGNN_REFERENCE_ROOT, default /root/reference) -> skipped on the GPU box.  Runs in a fresh interpreter
because the point is import order: overlay first, then `ptgnn.implementations.*`.  (The ppi and varmisuse train modules cannot be
imported under Python 3.12 even without the overlay -- `GraphData[...]` generic-arity TypeError inside the reference -- so the
Typilus and Graph2Seq factories are the ones exercised.)"""
import os
import subprocess
import sys

import pytest

from oracle.refimport import REFERENCE_ROOT, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {stubs!r}); sys.path.insert(0, {ref!r})
import ptgnn_b200 as P
import ptgnn_b200.overlay as ov
report = ov.install(force_torch_scatter=True)
assert report["layers"] and report["container"] and report["metrics"], report
import torch_scatter
assert torch_scatter.__name__ == "ptgnn_b200.torch_scatter_shim" and hasattr(torch_scatter, "scatter_log_softmax")
from torch_scatter.composite import scatter_logsumexp      # the form grucopydecoder.py:10 uses

# --- the reference's factories, imported AFTER the overlay, unchanged ---
import ptgnn.implementations.typilus.train as typilus_train
import ptgnn.implementations.graph2seq.train as g2s_train   # imports GatedMessagePassingLayer by module path too
from ptgnn.baseneuralmodel import ModuleWithMetrics
from ptgnn.neuralmodels.gnn import GraphNeuralNetworkModel
import ptgnn.neuralmodels.gnn.graphneuralnetwork as ref_gnn_mod
assert typilus_train.GatedMessagePassingLayer is P.GatedMessagePassingLayer
assert typilus_train.MlpMessagePassingLayer is P.MlpMessagePassingLayer
assert ref_gnn_mod.GraphNeuralNetwork is P.GraphNeuralNetwork
from ptgnn.neuralmodels.gnn.messagepassing.residuallayers import ConcatResidualLayer as RefConcat
assert issubclass(RefConcat, P.AbstractMessagePassingLayer)     # the reference's residual layers now derive from our base

# --- build the Typilus Graph2Class model exactly as the reference does ---
import random
random.seed(0)
def sample(i):
    n = 12
    nodes = [random.choice(["foo_bar", "baz", "x", "getValue", "int", "self"]) for _ in range(n)]
    edges = {{"NEXT": {{str(j): [j + 1] for j in range(n - 1)}}, "CHILD": {{"0": [3, 4], "5": [6]}}, "OCCURRENCE_OF": {{}}}}
    return {{"nodes": nodes, "edges": edges, "token-sequence": list(range(n)),
            "supernodes": {{"2": {{"name": "a", "annotation": random.choice(["int", "str"])}}, "7": {{"name": "b", "annotation": "int"}}}}}}
data = [sample(i) for i in range(8)]
model = typilus_train.create_graph2class_gnn_model(hidden_state_size=64)
model.compute_metadata(iter(data), parallelize=False)
nn_module = model.build_neural_module()
gnns = [m for m in nn_module.modules() if isinstance(m, P.GraphNeuralNetwork)]
assert len(gnns) == 1, "the reference built our container"
gnn = gnns[0]
kinds = [type(l).__module__ + "." + type(l).__name__ for l in gnn.message_passing_layers]
assert kinds.count("ptgnn_b200.messagepassing.MlpMessagePassingLayer") == 8, kinds
assert any(k.endswith("residuallayers.ConcatResidualLayer") and k.startswith("ptgnn.") for k in kinds), kinds
# --- metrics protocol under a reference parent (modulewithmetrics.py:44-57) ---
assert isinstance(gnn, ModuleWithMetrics)
gnn._GraphNeuralNetwork__num_edges = 7; gnn._GraphNeuralNetwork__num_nodes = 3; gnn._GraphNeuralNetwork__num_graphs = 1
nn_module._Graph2ClassModule__num_samples = 1      # the parent's own metric divides by its sample count
rep = nn_module.report_metrics()
assert rep["num_edges"] == 7 and rep["num_nodes"] == 3 and rep["num_graphs"] == 1, rep
nn_module.reset_metrics()
assert gnn._module_metrics() == {{"num_graphs": 0, "num_nodes": 0, "num_edges": 0}}
# --- minibatch tensors flow through the reference's tensorize/finalize into our container's signature (CPU: no kernel call) ---
mb = model.initialize_minibatch()
for d in data[:3]:
    model.extend_minibatch_with(model.tensorize(d), mb)
final = model.finalize_minibatch(mb, "cpu")
import torch
try:
    with torch.no_grad():
        nn_module.eval()(**final)
except P._native.NativeLibraryError as e:      # reaches the first ptgnn_b200 layer, which refuses CPU tensors (no fallback)
    assert "no CPU fallback" in str(e)
else:
    raise AssertionError("expected the ptgnn_b200 layer to refuse CPU tensors")
ov.uninstall()
assert ref_gnn_mod.GraphNeuralNetwork is not P.GraphNeuralNetwork
print("OVERLAY-OK", kinds[:3])
'''


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_reference_implementations_build_ptgnn_b200_layers_unchanged():
    code = SCRIPT.format(root=ROOT, stubs=os.path.join(ROOT, "oracle", "refstubs"), ref=REFERENCE_ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "OVERLAY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_overlay_without_reference_installs_only_the_shim():
    code = ("import sys; sys.path.insert(0, %r)\nimport ptgnn_b200.overlay as ov\nr = ov.install(force_torch_scatter=True)\n"
            "import torch_scatter\nassert torch_scatter.__name__ == 'ptgnn_b200.torch_scatter_shim'\n"
            "assert r['layers'] is False and 'note' in r, r\nprint('SHIM-OK')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "SHIM-OK" in r.stdout, r.stdout + r.stderr[-3000:]
