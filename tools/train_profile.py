"""Where a training step (forward + backward, config 2, 8 gated layers) spends its GPU time: torch.profiler kernel table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import ptgnn_b200 as P  # noqa: E402
from ptgnn_b200.synthetic import graph2class_batch  # noqa: E402

layers_n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
b = graph2class_batch()
torch.manual_seed(0)
layers = [P.GatedMessagePassingLayer(128, 128, 17, "sum") for _ in range(layers_n)]
gnn = P.GraphNeuralNetwork(layers, torch.nn.Identity(), True, True).cuda().train()
h = torch.randn(b.num_nodes, 128).cuda().requires_grad_(True)
adj = gnn.expand_adjacency([(s.cuda(), t.cuda()) for s, t in b.adjacency_lists], b.num_nodes, "cuda")


def step():
    for p in gnn.parameters():
        p.grad = None
    h.grad = None
    gnn.gnn(h, adj, None, None, {}, {}).mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print(f"train step: {e0.elapsed_time(e1) / 3:.2f} ms ({layers_n} layers)")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
