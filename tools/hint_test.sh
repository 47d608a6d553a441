#!/bin/bash
for hm in 0 1 2 4 7; do
  echo "== PTGNN_L2_HINTS=$hm"
  PTGNN_L2_HINTS=$hm timeout -s KILL 200 python - <<'PY' 2>&1 | tail -2
import torch, sys, os
sys.path.insert(0, os.getcwd())
import bench
from ptgnn_b200 import _native as N
batch = bench.make_batch("graph2class")
gnn = bench.build_model(17, "sum").cuda()
h = torch.randn(batch.num_nodes, 128).cuda()
adj = [(s.cuda(), t.cuda()) for s, t in batch.adjacency_lists]
ident = torch.arange(batch.num_nodes, device="cuda")
ex = list(adj) + [(t, s) for s, t in adj] + [(ident, ident)]
with torch.no_grad():
    for _ in range(2): gnn.gnn(h, ex, None, None, {}, {})
    N.kernel_timing(True); N.read_kernel_timing()
    for _ in range(3): gnn.gnn(h, ex, None, None, {}, {})
    kt = N.read_kernel_timing()
print({k: round(v[0] / max(v[1], 1), 4) for k, v in kt.items() if v[1]})
PY
done
