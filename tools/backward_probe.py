"""Diagnostics for the max / min backward: native arg-max edge ids vs the oracle's, and where gradients differ."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import ptgnn_b200 as P  # noqa: E402
from ptgnn_b200 import _native as N  # noqa: E402
from ptgnn_b200 import composed as C  # noqa: E402
from helpers import gated_oracle_args, random_adjacency  # noqa: E402
from oracle import ptgnn_oracle as O  # noqa: E402

gen = torch.Generator().manual_seed(3)
torch.manual_seed(3)
n, counts, H = 700, [2500, 0, 900, 40], 64
adj = random_adjacency(gen, n, counts)
layer = P.GatedMessagePassingLayer(H, H, len(counts), "max")
h0 = torch.randn(n, H, generator=gen)
args = gated_oracle_args({k: v.clone() for k, v in layer.state_dict().items()})
msgs = torch.cat([torch.nn.functional.linear(h0[s], w) for (s, t), w in zip(adj, args["edge_weights"])])
tgts = torch.cat([t for s, t in adj])
out_ref, arg_ref = O.scatter_with_arg(msgs, tgts, n, "max")
adj_d = [(s.cuda(), t.cuda()) for s, t in adj]
plan = P.plan_for(adj_d, n)
W = [w.cuda() for w in args["edge_weights"]]
msg = C.edge_messages(plan, h0.cuda(), None, W, False)
agg, arg = C.segment_reduce(msg, plan, N.REDUCE["max"], return_arg=True)
print("messages max abs diff", (msg.cpu() - msgs).abs().max().item())
print("agg max abs diff", (agg.cpu() - out_ref).abs().max().item())
bad = (arg.cpu() != arg_ref)
print("arg mismatches", int(bad.sum()), "of", arg.numel())
if bad.any():
    i = torch.nonzero(bad)[:8]
    for v, d in i.tolist():
        a, b = int(arg[v, d]), int(arg_ref[v, d])
        print(f"  target {v} feature {d}: ours edge {a} (msg {msgs[a, d] if a < msgs.shape[0] else None}, tgt {int(tgts[a]) if a < len(tgts) else None}), "
              f"oracle edge {b} (msg {msgs[b, d] if b < msgs.shape[0] else None}, tgt {int(tgts[b]) if b < len(tgts) else None})")
