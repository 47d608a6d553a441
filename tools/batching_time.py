"""Minibatch finalisation at BASELINE config 2's shape (80 graphs x 2,560 nodes, 8 edge types, 450,560 raw edges): the device-side
MinibatchAssembler against the reference's procedure (oracle/batching_oracle.py: numpy adds per graph + the per-node Python loop),
both starting from the same per-graph LOCAL arrays.  python tools/batching_time.py"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ptgnn_b200 as P  # noqa: E402
from oracle import batching_oracle as B  # noqa: E402  (tools/: measurement only)

rng = np.random.RandomState(0)
G, n, T = 80, 2560, 8
fractions = np.array([0.30, 0.25, 0.15, 0.10, 0.08, 0.06, 0.04, 0.02])
graphs = []
for g in range(G):
    adj = []
    for t in range(T):
        e = int(5632 * fractions[t])
        adj.append((rng.randint(0, n, e).astype(np.int32), rng.randint(0, n, e).astype(np.int32)))
    graphs.append(types.SimpleNamespace(adjacency_lists=adj, reference_nodes={"ref": rng.randint(0, n, 4).astype(np.int32)}, num_nodes=n))


def device_run():
    asm = P.MinibatchAssembler(T, 10 ** 9)
    mb = asm.initialize_minibatch()
    for g in graphs:
        asm.extend_minibatch_with(g, mb)
    out = asm.finalize_minibatch(mb, "cuda")
    torch.cuda.synchronize()
    return out


def reference_run():
    mb = B.initialize_minibatch(T)
    for g in graphs:
        B.extend_minibatch_with(g, mb, 10 ** 9)
    out = B.finalize_minibatch(mb)
    # the reference then builds int64 device tensors from the numpy arrays (graphneuralnetwork.py:463-491)
    adj = [(torch.tensor(s, dtype=torch.int64, device="cuda"), torch.tensor(t, dtype=torch.int64, device="cuda")) for s, t in out["adjacency_lists"]]
    n2g = torch.tensor(out["node_to_graph_idx"], dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    return adj, n2g


for fn, name in ((device_run, "MinibatchAssembler (device-side finalisation)"), (reference_run, "reference procedure (numpy + per-node Python loop + H2D)")):
    fn()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {dt * 1e3:.2f} ms per minibatch of {G * n} nodes, {sum(len(a[0]) for g in graphs for a in g.adjacency_lists)} raw edges")
