"""Debug helper (GPU box): per-layer error vs the oracle for the smoke configuration, each layer fed identical inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptgnn_b200 as P
from oracle import ptgnn_oracle as O
from ptgnn_b200.synthetic import block_diagonal_batch

b = block_diagonal_batch(4, 256, 3000, (0.5, 0.3, 0.2), seed=1)
torch.manual_seed(0)
T = 2 * len(b.adjacency_lists) + 1
h = torch.randn(b.num_nodes, 64, generator=torch.Generator().manual_seed(2))
adj = O.expand_adjacency(b.adjacency_lists, b.num_nodes, True, True)
adj_c = [(s.cuda(), t.cuda()) for s, t in adj]
feats = [torch.empty(a[0].shape[0], 0) for a in adj]

def err(a, r):
    e = (a.double() - r.double()).abs()
    return f"max_abs={e.max().item():.3e} mean_abs={e.mean().item():.3e} argmax={tuple(int(x) for x in divmod(int(e.argmax()), a.shape[1]))}"

for agg in ("sum", "max"):
    gated = P.GatedMessagePassingLayer(64, 64, T, agg)
    sd = {k: v.clone() for k, v in gated.state_dict().items()}; p = "_GatedMessagePassingLayer__"
    ref = O.gated_layer_forward(h, adj, feats, [sd[f"{p}edge_message_transformation_layers.{t}.weight"] for t in range(T)],
                                sd[p + "state_update.weight_ih"], sd[p + "state_update.weight_hh"],
                                sd[p + "state_update.bias_ih"], sd[p + "state_update.bias_hh"], agg)
    with torch.no_grad():
        got = gated.cuda().eval()(h.cuda(), adj_c).cpu()
    print(f"gated {agg}: {err(got, ref)}")

for name, kw in [("mlp max default", {}), ("mlp sum default", dict(agg="sum")),
                 ("mlp max no-LN", dict(use_layer_norm=False)), ("mlp max no-dense", dict(use_dense_layer=False)),
                 ("mlp max bare", dict(use_layer_norm=False, use_dense_layer=False, message_activation=None)),
                 ("mlp max no-act", dict(message_activation=None))]:
    kw = dict(kw); agg = kw.pop("agg", "max")
    mlp = P.MlpMessagePassingLayer(64, 64, 64, T, agg, **kw)
    sd = {k: v.clone() for k, v in mlp.state_dict().items()}; p = "_MlpMessagePassingLayer__"
    i = 0; okw = {}
    if kw.get("use_layer_norm", True):
        okw.update(ln_weight=sd[f"{p}state_update.{i}.weight"], ln_bias=sd[f"{p}state_update.{i}.bias"]); i += 1
    if kw.get("use_dense_layer", True):
        okw.update(dense_weight=sd[f"{p}state_update.{i}.weight"], dense_bias=sd[f"{p}state_update.{i}.bias"])
    ref = O.mlp_layer_forward(h, adj, feats, [[sd[f"{p}edge_message_transformation_layers.{t}._MLP__mlp_modules.1.weight"]] for t in range(T)],
                              agg, message_activation=None if ("message_activation" in kw) else "gelu", **okw)
    with torch.no_grad():
        got = mlp.cuda().eval()(h.cuda(), adj_c).cpu()
    print(f"{name}: {err(got, ref)}")
