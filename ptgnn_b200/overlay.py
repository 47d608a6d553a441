"""Makes the reference's own packages run on the ptgnn_b200 kernels WITHOUT editing them.

The reference constructs its GNN by class name (`/root/reference/ptgnn/neuralmodels/gnn/graphneuralnetwork.py:298-305`), its
factories import the layer classes by module path (`implementations/typilus/train.py:24-33`, `ppi/train.py:31-32`,
`varmisuse/train.py:17-25`) and its modules import ``torch_scatter`` by name (`abstractmessagepassing.py:4`).  ``install()``

1. registers ``ptgnn_b200.torch_scatter_shim`` as ``torch_scatter`` (if the real wheel is absent, or ``force_torch_scatter``),
2. pre-seeds ``sys.modules`` so that ``ptgnn.neuralmodels.gnn.messagepassing.{abstract,gated,mlp}messagepassing`` resolve to the
   ptgnn_b200 classes (everything else -- residual layers, PNA, GraphNorm, embedders, trainers -- stays the reference's), and
   re-binds those names in reference modules that were imported earlier,
3. replaces ``GraphNeuralNetwork`` inside the reference's ``graphneuralnetwork`` module (``GraphNeuralNetworkModel`` stays),
4. registers the ptgnn_b200 container as a (virtual) ``ModuleWithMetrics`` so that a reference parent's ``report_metrics()`` /
   ``reset_metrics()`` see its ``num_graphs / num_nodes / num_edges`` (`baseneuralmodel/modulewithmetrics.py:44-57`).

After ``install()``: ``import ptgnn.implementations.ppi.train`` etc. build ptgnn_b200 layers, unchanged.  ``uninstall()``
restores the reference's classes.  The reference must be importable as ``ptgnn`` for steps 2-4 (it is not on the GPU test box;
``install()`` then only does step 1 and says so in its report).
"""
import importlib
import sys
import types
from typing import Dict

from . import gnn as _gnn
from . import messagepassing as _mp

_LAYER_MODULES = {
    "ptgnn.neuralmodels.gnn.messagepassing.abstractmessagepassing": ("AbstractMessagePassingLayer", "AbstractMessageAggregation"),
    "ptgnn.neuralmodels.gnn.messagepassing.gatedmessagepassing": ("GatedMessagePassingLayer",),
    "ptgnn.neuralmodels.gnn.messagepassing.mlpmessagepassing": ("MlpMessagePassingLayer",),
}
_saved: Dict[str, object] = {}


def _reference_importable() -> bool:
    try:
        return importlib.util.find_spec("ptgnn") is not None
    except (ImportError, ValueError):
        return False


def _stub_module(name: str, names) -> types.ModuleType:
    m = types.ModuleType(name, f"ptgnn_b200 overlay of the reference module {name}")
    for n in names:
        setattr(m, n, getattr(_mp, n))
    m.__ptgnn_b200_overlay__ = True
    return m


def install(force_torch_scatter: bool = False) -> Dict[str, object]:
    report: Dict[str, object] = {"torch_scatter": "real", "layers": False, "container": False, "metrics": False}
    # 1. torch_scatter
    have_real = False
    if not force_torch_scatter:
        try:
            have_real = importlib.util.find_spec("torch_scatter") is not None and not getattr(sys.modules.get("torch_scatter"), "__ptgnn_b200_overlay__", False)
        except (ImportError, ValueError):
            have_real = False
    if not have_real:
        from . import torch_scatter_shim as shim

        shim.__ptgnn_b200_overlay__ = True
        sys.modules["torch_scatter"] = shim
        sys.modules["torch_scatter.composite"] = shim.composite
        report["torch_scatter"] = "ptgnn_b200.torch_scatter_shim"
    if not _reference_importable():
        report["note"] = "reference package `ptgnn` not importable: only the torch_scatter shim was installed"
        return report
    # 2. layer modules: replace already-imported reference classes, pre-seed the ones not imported yet
    replaced = {}
    for mod_name, names in _LAYER_MODULES.items():
        old = sys.modules.get(mod_name)
        if old is not None and not getattr(old, "__ptgnn_b200_overlay__", False):
            _saved[mod_name] = old
            for n in names:
                if hasattr(old, n):
                    replaced[getattr(old, n)] = getattr(_mp, n)
        sys.modules[mod_name] = _stub_module(mod_name, names)
    for mod_name, mod in list(sys.modules.items()):      # `from ... import GatedMessagePassingLayer` bindings made earlier
        if mod is None or not mod_name.startswith("ptgnn.") or getattr(mod, "__ptgnn_b200_overlay__", False):
            continue
        for attr, val in list(vars(mod).items()):
            try:
                new = replaced.get(val)
            except TypeError:        # unhashable module attribute
                continue
            if new is not None:
                _saved.setdefault(f"{mod_name}:{attr}", val)
                setattr(mod, attr, new)
    report["layers"] = True
    # 3. the container class inside the reference's graphneuralnetwork module (GraphNeuralNetworkModel looks it up at call time)
    ref_gnn = importlib.import_module("ptgnn.neuralmodels.gnn.graphneuralnetwork")
    if ref_gnn.GraphNeuralNetwork is not _gnn.GraphNeuralNetwork:
        _saved["container"] = ref_gnn.GraphNeuralNetwork
        ref_gnn.GraphNeuralNetwork = _gnn.GraphNeuralNetwork
    pkg = sys.modules.get("ptgnn.neuralmodels.gnn")
    if pkg is not None and getattr(pkg, "GraphNeuralNetwork", None) is _saved.get("container"):
        pkg.GraphNeuralNetwork = _gnn.GraphNeuralNetwork
    report["container"] = True
    # 4. metrics protocol
    report["metrics"] = _gnn.register_with_reference_metrics()
    return report


def uninstall() -> None:
    for key, val in list(_saved.items()):
        if key == "container":
            ref_gnn = sys.modules.get("ptgnn.neuralmodels.gnn.graphneuralnetwork")
            if ref_gnn is not None:
                ref_gnn.GraphNeuralNetwork = val
            pkg = sys.modules.get("ptgnn.neuralmodels.gnn")
            if pkg is not None and getattr(pkg, "GraphNeuralNetwork", None) is _gnn.GraphNeuralNetwork:
                pkg.GraphNeuralNetwork = val
        elif ":" in key:
            mod_name, attr = key.split(":", 1)
            if mod_name in sys.modules:
                setattr(sys.modules[mod_name], attr, val)
        else:
            sys.modules[key] = val
    for mod_name in _LAYER_MODULES:
        if getattr(sys.modules.get(mod_name), "__ptgnn_b200_overlay__", False):
            del sys.modules[mod_name]
    for name in ("torch_scatter", "torch_scatter.composite"):
        if getattr(sys.modules.get(name), "__ptgnn_b200_overlay__", False) or name.endswith("composite"):
            mod = sys.modules.get(name)
            if mod is not None and mod.__name__.startswith("ptgnn_b200."):
                del sys.modules[name]
    _saved.clear()
