"""EGCMessagePassingLayer: oracle pinned to the unmodified reference class (tests/golden/egc_*.npz), the native
layer against both; same state_dict keys as the reference."""
import pytest
import torch

from helpers import assert_close, golden_adjacency, golden_state_dict, load_golden, random_adjacency
from oracle import ptgnn_oracle as O

P_ = "_EGCMessagePassingLayer__"


def _oracle_from_sd(h, adj, sd, agg, heads=8, bases=4):
    T = sum(1 for k in sd if k.startswith(P_ + "bases."))
    return O.egc_layer_forward(h, adj, [sd[f"{P_}bases.{t}.weight"] for t in range(T)], sd[P_ + "weight_coeffs.weight"],
                               sd[P_ + "weight_coeffs.bias"], agg, heads, bases)


@pytest.mark.parametrize("agg", ["sum", "max"])
def test_egc_oracle_reproduces_the_reference(agg):
    g = load_golden(f"egc_{agg}")
    out = _oracle_from_sd(torch.from_numpy(g["h"]), golden_adjacency(g), golden_state_dict(g), agg)
    assert torch.equal(out, torch.from_numpy(g["out"]))


def test_egc_state_dict_keys_and_seeded_init_match_the_reference():
    import ptgnn_b200 as P

    g = load_golden("egc_sum")
    torch.manual_seed(31)
    torch.Generator().manual_seed(31)
    layer = P.EGCMessagePassingLayer(64, 64, 3, "sum", num_bases=4, num_heads=8)
    sd = golden_state_dict(g)
    assert sorted(layer.state_dict()) == sorted(sd)
    layer.load_state_dict(sd, strict=True)
    assert layer.input_state_dimension == 64 and layer.output_state_dimension == 64


@pytest.mark.gpu
@pytest.mark.parametrize("agg", ["sum", "max"])
def test_egc_native_vs_reference_golden(agg):
    import ptgnn_b200 as P

    g = load_golden(f"egc_{agg}")
    layer = P.EGCMessagePassingLayer(64, 64, 3, agg, num_bases=4, num_heads=8)
    layer.load_state_dict(golden_state_dict(g), strict=True)
    layer = layer.cuda().eval()
    adj = [(s.cuda(), t.cuda()) for s, t in golden_adjacency(g)]
    with torch.no_grad():
        out = layer(torch.from_numpy(g["h"]).cuda(), adj)
    assert_close(out, torch.from_numpy(g["out"]), tol=2e-5, what=f"EGC {agg} vs the reference's output")


@pytest.mark.gpu
@pytest.mark.parametrize("agg,n,counts,H,out_dim,heads,bases", [("mean", 2000, [7000, 1500, 0, 90], 128, 128, 8, 4),
                                                                 ("min", 500, [1500, 400], 64, 96, 4, 2)])
def test_egc_native_vs_oracle(agg, n, counts, H, out_dim, heads, bases):
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(8)
    torch.manual_seed(8)
    adj = random_adjacency(gen, n, counts)
    layer = P.EGCMessagePassingLayer(H, out_dim, len(counts), agg, num_bases=bases, num_heads=heads)
    h = torch.randn(n, H, generator=gen)
    ref = _oracle_from_sd(h, adj, {k: v.clone() for k, v in layer.state_dict().items()}, agg, heads, bases)
    layer = layer.cuda().eval()
    with torch.no_grad():
        out = layer(h.cuda(), [(s.cuda(), t.cuda()) for s, t in adj])
    assert_close(out, ref, tol=2e-5, what=f"EGC {agg}")
