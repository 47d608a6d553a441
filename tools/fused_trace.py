"""Timeline of the fused aggregation kernel (CTA 0) at config 2: PTGNN_FUSED_TRACE=1 python tools/fused_trace.py [f32|bf16]
Prints per-role stage durations (cycles, clock64 of the SM) and dumps the raw events to gpurun_out/fused_trace_<dtype>.txt."""
import ctypes
import os
import sys

os.environ["PTGNN_FUSED_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ptgnn_b200 as P  # noqa: E402
from ptgnn_b200 import _native as N  # noqa: E402
from ptgnn_b200.synthetic import graph2class_batch  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
b = graph2class_batch()
torch.manual_seed(0)
layer = P.GatedMessagePassingLayer(128, 128, 17, "sum").cuda().eval()
gnn = P.GraphNeuralNetwork([layer], torch.nn.Identity(), True, True).cuda().eval()
h = torch.randn(b.num_nodes, 128).cuda()
if dtype == "bf16":
    h = h.to(torch.bfloat16)
adj = gnn.expand_adjacency([(s.cuda(), t.cuda()) for s, t in b.adjacency_lists], b.num_nodes, "cuda")
with torch.no_grad():
    for _ in range(3):
        layer(h, adj)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (5 * 2048))()
lib = N.lib()
lib.ptgnn_b200_debug_fused_trace.argtypes = [ctypes.c_void_p]
assert lib.ptgnn_b200_debug_fused_trace(buf) == 1
roles = ["gather", "mma", "epi0", "epi1", "wload"]
ev = {}
for r, name in enumerate(roles):
    ev[name] = [((v >> 24), (v >> 8) & 0xFFFF, v & 0xFF) for v in buf[r * 2048:(r + 1) * 2048] if v]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
t0 = min(e[0][0] for e in ev.values() if e)
with open(os.path.join(ROOT, "gpurun_out", f"fused_trace_{dtype}.txt"), "w") as f:
    for name in roles:
        for clk, step, tag in ev[name]:
            f.write(f"{name} {clk - t0} {step} {tag}\n")


def stats(name, tag_a, tag_b, label):
    by = {}
    for clk, step, tag in ev[name]:
        by.setdefault(step, {})[tag] = clk
    d = [v[tag_b] - v[tag_a] for v in by.values() if tag_a in v and tag_b in v]
    if d:
        d.sort()
        print(f"  {label:44s} n={len(d):4d} mean {sum(d) / len(d):8.0f}  p50 {d[len(d) // 2]:7d}  p90 {d[int(len(d) * 0.9)]:7d}  max {d[-1]:7d}")


def period(name, tag, label):
    c = sorted(clk for clk, _, t in ev[name] if t == tag)
    if len(c) > 2:
        print(f"  {label:44s} n={len(c):4d} mean period {(c[-1] - c[0]) / (len(c) - 1):8.0f}   span {c[-1] - c[0]}")


print(f"== {dtype}: events recorded", {k: len(v) for k, v in ev.items()})
print("gather (per step):")
stats("gather", 1, 2, "wait x_empty")
stats("gather", 2, 3, "issue LDGSTS")
period("gather", 4, "x_full arrivals")
print("mma (per step):")
stats("mma", 10, 11, "wait acc_empty")
stats("mma", 11, 12, "wait x_full")
stats("mma", 12, 13, "wait w_full")
stats("mma", 13, 14, "issue MMAs + commits")
period("mma", 14, "steps")
for e in ("epi0", "epi1"):
    print(e, "(per sub-group):")
    stats(e, 20, 21, "wait acc_full")
    stats(e, 21, 22, "drain + combine")
    period(e, 22, "sub-groups")
    c23 = sorted(clk for clk, _, t in ev[e] if t == 23)
    c24 = sorted(clk for clk, _, t in ev[e] if t == 24)
    if c23 and c24:
        print(f"  block write-out + re-init: mean {sum(b - a for a, b in zip(c23, c24)) / min(len(c23), len(c24)):.0f} cycles, {len(c23)} blocks")
print("weight loader (per group):")
stats("wload", 30, 31, "issue LDGs")
stats("wload", 31, 32, "wait w_empty (+ LDG latency)")
stats("wload", 32, 33, "STTM + arrive")
period("wload", 33, "groups")
# cross-role latencies: x_full arrive (gather tag 4, step s) -> mma got x (tag 12, step s); mma done (14) -> epi got acc (21)
g4 = {s: c for c, s, t in ev["gather"] if t == 4}
m12 = {s: c for c, s, t in ev["mma"] if t == 12}
m14 = {s: c for c, s, t in ev["mma"] if t == 14}
d = sorted(m12[s] - g4[s] for s in g4 if s in m12)
if d:
    print(f"x_full arrive -> MMA sees it: p50 {d[len(d) // 2]} p90 {d[int(len(d) * .9)]}")
g2 = {s: c for c, s, t in ev["gather"] if t == 2}
d = sorted(g4[s] - g2[s] for s in g4 if s in g2)
if d:
    print(f"gather: slot granted -> data landed (arrive): p50 {d[len(d) // 2]} p90 {d[int(len(d) * .9)]} mean {sum(d) / len(d):.0f}")
