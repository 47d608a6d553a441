"""Diagnostic: MLP.forward (native dense kernels) vs CPU F.linear, fresh allocator vs poisoned (garbage-filled) allocator blocks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import ptgnn_b200 as P  # noqa: E402
from ptgnn_b200 import composed as C  # noqa: E402


def run(tag):
    torch.manual_seed(0)
    for kw in (dict(hidden_layers=[48, 20], use_biases=True), dict(hidden_layers=1), dict(hidden_layers=0),
               dict(hidden_layers=[7], use_biases=True, activation=torch.nn.Tanh())):
        mlp = P.MLP(36, 24, **kw).eval()
        x = torch.randn(500, 36)
        ref, cur = x, x.cuda()
        lins = mlp.linears
        errs = []
        for i, l in enumerate(lins):
            ref = F.linear(ref, l.weight, l.bias)
            act = mlp.activation if i + 1 < len(lins) else None
            if act is not None:
                ref = act(ref)
            with torch.no_grad():
                got = C.linear(ref_in.cuda() if (ref_in := None) is not None else cur, l.weight.cuda(), None if l.bias is None else l.bias.cuda(), act)
            errs.append(((got.cpu() - ref.detach()).abs() / ref.detach().abs().clamp(min=1)).max().item())
            cur = ref.detach().cuda()          # teacher forcing: every layer on the reference's input
        print(tag, kw, ["%.2e" % e for e in errs])


run("fresh   ")
junk = [torch.full((1 << 22,), float("nan"), device="cuda") for _ in range(8)] + [torch.full((1 << 20,), 3e30, device="cuda") for _ in range(32)]
del junk
run("poisoned")
