"""Debug: where does the end-to-end step time go (H2D, kernels, D2H, overlap)?"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
import ptgnn_b200 as P
dev = torch.device("cuda", 0)
batch = bench.make_batch("graph2class")
gnn = bench.build_model(17, "sum").to(dev)
n = batch.num_nodes
h_host = torch.randn(n, 128).pin_memory()
adj_host = [(s.pin_memory(), t.pin_memory()) for s, t in batch.adjacency_lists]
out_host = torch.empty(n, 128).pin_memory()
h_dev = h_host.to(dev); adj_dev = [(s.to(dev), t.to(dev)) for s, t in adj_host]
n2g = batch.node_to_graph_idx.to(dev)
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
def h2d():
    h_dev.copy_(h_host, non_blocking=True)
    for (ds, dt), (hs, ht) in zip(adj_dev, adj_host): ds.copy_(hs, non_blocking=True); dt.copy_(ht, non_blocking=True)
def compute():
    P.clear_plan_cache()
    with torch.no_grad():
        return gnn(node_data={"input": h_dev}, adjacency_lists=list(adj_dev), edge_feature_data=[], node_to_graph_idx=n2g,
                   reference_node_ids={}, reference_node_graph_idx={}, num_graphs=batch.num_graphs).output_node_representations
res = compute()
def d2h(): out_host.copy_(res, non_blocking=True)
print("h2d ms", timeit(h2d), "compute ms", timeit(compute), "d2h ms", timeit(d2h))
t0 = time.perf_counter()
for _ in range(10): compute()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue ms/step", (t1 - t0) * 100, "total ms/step", (t2 - t0) * 100)
# overlap test: copies on side streams while computing
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def overlapped():
    with torch.cuda.stream(s1): h2d()
    with torch.cuda.stream(s2): d2h()
    compute()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print("compute || h2d || d2h ms", timeit(overlapped))
def overlapped_h2d():
    with torch.cuda.stream(s1): h2d()
    compute(); torch.cuda.current_stream().wait_stream(s1)
print("compute || h2d ms", timeit(overlapped_h2d))
def overlapped_d2h():
    with torch.cuda.stream(s2): d2h()
    compute(); torch.cuda.current_stream().wait_stream(s2)
print("compute || d2h ms", timeit(overlapped_d2h))
