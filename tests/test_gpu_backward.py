"""Backward of GatedMessagePassingLayer against torch.autograd through the CPU oracle.
Reference: `loss.backward()` through the layers, /root/reference/ptgnn/baseneuralmodel/trainer.py:221-236."""
import pytest
import torch

from helpers import gated_oracle_args, random_adjacency
from oracle import ptgnn_oracle as O

pytestmark = pytest.mark.gpu


def _close(got: torch.Tensor, ref: torch.Tensor, what: str, tol: float = 1e-4):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} != {tuple(ref.shape)}"
    scale = max(1.0, float(ref.abs().max()))
    err = float((got - ref).abs().max()) / scale
    rel = float((got - ref).norm() / ref.norm().clamp(min=1e-30))
    assert err <= tol and rel <= tol, f"{what}: max error {err:.3e} (scaled), rel L2 {rel:.3e}"


def _states_without_near_ties(agg, adj, n, H, gen, h0, message_fn):
    """Which edge wins a (target, feature) under max / min is decided on messages that carry ~1e-6 of rounding error in the kernels:
    two candidates within that distance may legitimately swap, which moves one gradient element between two rows.  Returns states
    drawn so that no runner-up is within 5e-6 of its winner (at ~40k (target, feature) pairs about every second draw qualifies,
    which is why max / min run on the smaller cases only)."""
    if agg not in ("max", "min"):
        return h0
    sign = 1.0 if agg == "max" else -1.0
    tgt = torch.cat([t for _s, t in adj])
    for _ in range(24):
        with torch.no_grad():
            msgs = sign * torch.cat([message_fn(h0, s, t, i) for i, (s, t) in enumerate(adj)])
            top = O.scatter(msgs, tgt, n, "max")
            masked = torch.where(msgs >= top[tgt], torch.full_like(msgs, -3e38), msgs)
            second = O.scatter(masked, tgt, n, "max")
            gap = torch.where(second < -1e38, torch.full_like(top, 1.0), top - second)
            gap[torch.bincount(tgt, minlength=n) < 2] = 1.0          # no or one candidate: nothing to swap
        if float(gap.min()) > 5e-6:
            return h0
        h0 = torch.randn(n, H, generator=gen)
    pytest.fail("could not draw states without near-ties")


@pytest.mark.parametrize("agg,n,counts,H", [
    ("sum", 700, [2500, 0, 900, 40], 64), ("mean", 700, [2500, 0, 900, 40], 64), ("max", 700, [2500, 0, 900, 40], 64),
    ("min", 700, [2500, 0, 900, 40], 64), ("sum", 3000, [9000, 5000, 130, 1], 128), ("mean", 3000, [9000, 5000, 130, 1], 128),
    ("max", 300, [1200, 500, 30, 1], 128), ("min", 300, [1200, 500, 30, 1], 128)])
def test_gated_backward_vs_oracle_autograd(agg, n, counts, H):
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(3)
    torch.manual_seed(3)
    adj = random_adjacency(gen, n, counts)
    layer = P.GatedMessagePassingLayer(H, H, len(counts), agg)
    h0 = torch.randn(n, H, generator=gen)
    linears = [m for m in layer.modules() if isinstance(m, torch.nn.Linear)]
    h0 = _states_without_near_ties(agg, adj, n, H, gen, h0, lambda x, s, t, i: torch.nn.functional.linear(x[s], linears[i].weight))
    probe = torch.randn(n, H, generator=gen)                     # loss = <out, probe>: a generic upstream gradient

    # oracle: the same arithmetic as the reference layer, autograd on the CPU
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    args = gated_oracle_args(sd)
    leaves = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else [w.clone().requires_grad_(True) for w in v]) for k, v in args.items()}
    h_ref = h0.clone().requires_grad_(True)
    out_ref = O.gated_layer_forward(h_ref, adj, [torch.empty(c, 0) for c in counts], aggregation_fn=agg, **leaves)
    (out_ref * probe).sum().backward()

    layer = layer.cuda().train()
    h = h0.cuda().requires_grad_(True)
    adj_d = [(s.cuda(), t.cuda()) for s, t in adj]
    out = layer(h, adj_d)
    assert out.requires_grad
    _close(out, out_ref, "forward", 2e-5)
    (out * probe.cuda()).sum().backward()
    _close(h.grad, h_ref.grad, f"{agg}: d node_states")
    names = dict(layer.named_parameters())
    p = "_GatedMessagePassingLayer__"
    for t in range(len(counts)):
        _close(names[f"{p}edge_message_transformation_layers.{t}.weight"].grad, leaves["edge_weights"][t].grad, f"{agg}: dW_{t}")
    for ours, theirs in (("weight_ih", "gru_w_ih"), ("weight_hh", "gru_w_hh"), ("bias_ih", "gru_b_ih"), ("bias_hh", "gru_b_hh")):
        _close(names[f"{p}state_update.{ours}"].grad, leaves[theirs].grad, f"{agg}: d {ours}")


def test_training_steps_through_the_container():
    """Two optimiser steps on a 3-layer stack (one layer shared, as in the reference's GGNN configurations): the loss goes down, the
    derived-weight caches follow the parameter updates, and the result matches the same steps taken with the oracle on the CPU."""
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(9)
    torch.manual_seed(9)
    n, counts, H = 1500, [4000, 1500, 60], 128
    adj = random_adjacency(gen, n, counts)
    shared, last = P.GatedMessagePassingLayer(H, H, len(counts), "sum"), P.GatedMessagePassingLayer(H, H, len(counts), "max")
    gnn = P.GraphNeuralNetwork([shared, shared, last], torch.nn.Identity(), False, False)
    h0 = torch.randn(n, H, generator=gen)
    target = torch.randn(n, H, generator=gen)
    ref_params = [p.detach().clone().requires_grad_(True) for p in gnn.parameters()]
    names = [k for k, _ in gnn.named_parameters()]

    def oracle_forward(params):
        sd = dict(zip(names, params))
        x = h0
        for prefix, agg in (("0", "sum"), ("0", "sum"), ("2", "max")):
            lsd = {k.split(".", 2)[2]: v for k, v in sd.items() if k.startswith(f"_GraphNeuralNetwork__message_passing_layers.{prefix}.")}
            x = O.gated_layer_forward(x, adj, [torch.empty(c, 0) for c in counts], aggregation_fn=agg, **gated_oracle_args(lsd))
        return x

    assert any(k.startswith("_GraphNeuralNetwork__message_passing_layers.2.") for k in names)
    opt_ref = torch.optim.SGD(ref_params, lr=0.05)
    ref_losses = []
    for _ in range(2):
        opt_ref.zero_grad()
        loss = ((oracle_forward(ref_params) - target) ** 2).mean()
        loss.backward()
        opt_ref.step()
        ref_losses.append(float(loss.detach()))

    gnn = gnn.cuda().train()
    adj_d = [(s.cuda(), t.cuda()) for s, t in adj]
    opt = torch.optim.SGD(gnn.parameters(), lr=0.05)
    losses = []
    for _ in range(2):
        opt.zero_grad()
        out = gnn.gnn(h0.cuda(), adj_d, None, None, {}, {})
        loss = ((out - target.cuda()) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[1] < losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (losses, ref_losses)
    for (k, p), q in zip(gnn.named_parameters(), ref_params):
        _close(p, q, f"parameter {k} after two steps", 1e-4)


# ---- MlpMessagePassingLayer: the implementations' default layer (varmisuse/train.py:43-74, typilus/train.py:69-99) -----------------
@pytest.mark.parametrize("agg,use_target,n,counts,H", [
    ("sum", True, 700, [2500, 0, 900, 40], 64), ("mean", False, 700, [2500, 0, 900, 40], 64),
    ("max", True, 300, [1200, 500, 30, 1], 128), ("min", False, 300, [1200, 500, 30, 1], 128),
    ("sum", True, 2000, [9000, 3000, 200], 128)])
def test_mlp_backward_vs_oracle_autograd(agg, use_target, n, counts, H):
    import ptgnn_b200 as P
    from helpers import mlp_oracle_args

    gen = torch.Generator().manual_seed(13)
    torch.manual_seed(13)
    adj = random_adjacency(gen, n, counts)
    layer = P.MlpMessagePassingLayer(H, H, H, len(counts), agg, use_target_state_as_message_input=use_target)
    h0 = torch.randn(n, H, generator=gen)
    mlp_w = [m.single_linear.weight for m in layer.modules() if isinstance(m, P.MLP)]
    h0 = _states_without_near_ties(agg, adj, n, H, gen, h0, lambda x, s, t, i: torch.nn.functional.linear(
        torch.cat([x[s], x[t]], -1) if use_target else x[s], mlp_w[i]))
    probe = torch.randn(n, H, generator=gen)
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    args = mlp_oracle_args(sd)

    def leafify(v):
        if torch.is_tensor(v):
            return v.clone().requires_grad_(True)
        return [leafify(x) for x in v]

    leaves = {k: leafify(v) for k, v in args.items()}
    h_ref = h0.clone().requires_grad_(True)
    out_ref = O.mlp_layer_forward(h_ref, adj, [torch.empty(c, 0) for c in counts], leaves["edge_mlp_weights"], agg,
                                  use_target_state_as_message_input=use_target,
                                  **{k: v for k, v in leaves.items() if k != "edge_mlp_weights"})
    (out_ref * probe).sum().backward()

    layer = layer.cuda().train()
    h = h0.cuda().requires_grad_(True)
    out = layer(h, [(s.cuda(), t.cuda()) for s, t in adj])
    _close(out, out_ref, "forward", 5e-5)
    (out * probe.cuda()).sum().backward()
    # max / min: a near-tied winner may legitimately swap (see the gated test); these seeds have none closer than the kernels' rounding
    _close(h.grad, h_ref.grad, f"{agg}: d node_states", 2e-4)
    names = dict(layer.named_parameters())
    p = "_MlpMessagePassingLayer__"
    for t in range(len(counts)):
        _close(names[f"{p}edge_message_transformation_layers.{t}._MLP__mlp_modules.1.weight"].grad, leaves["edge_mlp_weights"][t][0].grad,
               f"{agg}: dW_{t}", 2e-4)
    _close(names[f"{p}state_update.0.weight"].grad, leaves["ln_weight"].grad, "d LayerNorm.weight", 2e-4)
    _close(names[f"{p}state_update.0.bias"].grad, leaves["ln_bias"].grad, "d LayerNorm.bias", 2e-4)
    _close(names[f"{p}state_update.1.weight"].grad, leaves["dense_weight"].grad, "d dense.weight", 2e-4)
    _close(names[f"{p}state_update.1.bias"].grad, leaves["dense_bias"].grad, "d dense.bias", 2e-4)


def test_mlp_output_dropout_is_a_mask_on_the_output():
    """nn.Dropout at the end of the state update (mlpmessagepassing.py:65): every output element is either 0 or the eval-mode value
    / (1 - p), gradients flow through the kept elements only."""
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(2)
    torch.manual_seed(2)
    n, counts, H, p = 500, [2000, 300], 128, 0.25
    adj = [(s.cuda(), t.cuda()) for s, t in random_adjacency(gen, n, counts)]
    layer = P.MlpMessagePassingLayer(H, H, H, len(counts), "max", dropout_rate=p).cuda()
    h = torch.randn(n, H, generator=gen).cuda()
    with torch.no_grad():
        ref = layer.eval()(h, adj)
        drop = layer.train()(h, adj)
    kept = drop != 0
    assert 0.70 < kept.float().mean().item() < 0.80
    assert torch.allclose(drop[kept], ref[kept] / (1 - p), rtol=1e-6, atol=1e-7)
    hg = h.clone().requires_grad_(True)
    out = layer(hg, adj)
    out.sum().backward()
    assert torch.isfinite(hg.grad).all() and hg.grad.abs().sum() > 0


# ---- gated layer where the gathered rows must exist: per-edge dropout (training mode), edge features under autograd ---------------
def test_gated_edge_features_backward_and_per_edge_dropout():
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(4)
    torch.manual_seed(4)
    n, counts, H, Fd = 600, [1800, 0, 500], 64, 8
    adj = random_adjacency(gen, n, counts)
    feats = [torch.randn(c, Fd, generator=gen) for c in counts]
    layer = P.GatedMessagePassingLayer(H, H, len(counts), "sum", edge_feature_dimension=Fd)
    h0 = torch.randn(n, H, generator=gen)
    probe = torch.randn(n, H, generator=gen)
    args = gated_oracle_args({k: v.clone() for k, v in layer.state_dict().items()})
    leaves = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else [w.clone().requires_grad_(True) for w in v]) for k, v in args.items()}
    h_ref = h0.clone().requires_grad_(True)
    f_ref = [f.clone().requires_grad_(True) for f in feats]
    out_ref = O.gated_layer_forward(h_ref, adj, f_ref, aggregation_fn="sum", **leaves)
    (out_ref * probe).sum().backward()

    layer = layer.cuda().train()
    h = h0.cuda().requires_grad_(True)
    f_dev = [f.cuda().requires_grad_(True) for f in feats]
    adj_d = [(s.cuda(), t.cuda()) for s, t in adj]
    out = layer(h, adj_d, edge_features=f_dev)
    _close(out, out_ref, "forward with edge features", 2e-5)
    (out * probe.cuda()).sum().backward()
    _close(h.grad, h_ref.grad, "d node_states (edge features)")
    _close(f_dev[0].grad, f_ref[0].grad, "d edge features")
    names = dict(layer.named_parameters())
    _close(names["_GatedMessagePassingLayer__edge_message_transformation_layers.2.weight"].grad, leaves["edge_weights"][2].grad, "dW_2 [D, H + F]")

    # per-edge dropout: with p = 1 - 1e-9 ... not testable for equality; check the mask statistics through a linear probe instead:
    # E[dropout(x)] = x, so the mean over many draws of the aggregate converges to the eval-mode aggregate
    drop = P.GatedMessagePassingLayer(H, H, len(counts), "sum", dropout_rate=0.5).cuda()
    with torch.no_grad():
        ev = drop.eval()(h0.cuda(), adj_d)
        tr = drop.train()(h0.cuda(), adj_d)
    assert tr.shape == ev.shape and torch.isfinite(tr).all()
    assert not torch.equal(tr, ev)                                       # a mask was applied ...
    assert not torch.equal(tr, drop(h0.cuda(), adj_d))                    # ... and it is drawn anew per call
    hg = h0.cuda().requires_grad_(True)
    drop(hg, adj_d).sum().backward()
    assert torch.isfinite(hg.grad).all() and hg.grad.abs().sum() > 0
