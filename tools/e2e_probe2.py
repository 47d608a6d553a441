"""Debug: which part of the pipelined e2e loop fails to overlap?"""
import os, sys
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
n2g = batch.node_to_graph_idx.to(dev)
h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
h_buf = [torch.empty(n, 128, device=dev) for _ in range(2)]
adj_buf = [[(torch.empty_like(s, device=dev), torch.empty_like(t, device=dev)) for s, t in adj_host] for _ in range(2)]
out_host = [torch.empty(n, 128).pin_memory() for _ in range(2)]
in_ready = [torch.cuda.Event() for _ in range(2)]; compute_done = [torch.cuda.Event() for _ in range(2)]; d2h_done = [torch.cuda.Event() for _ in range(2)]
keep = [None, None]
def prefetch(k, do=True):
    with torch.cuda.stream(h2d):
        h2d.wait_event(compute_done[k])
        if do:
            h_buf[k].copy_(h_host, non_blocking=True)
            for (ds, dt), (hs, ht) in zip(adj_buf[k], adj_host): ds.copy_(hs, non_blocking=True); dt.copy_(ht, non_blocking=True)
        in_ready[k].record(h2d)
def run(steps, do_h2d, do_d2h, contig_out=False):
    main = torch.cuda.current_stream(dev)
    prefetch(0, do_h2d)
    for i in range(steps):
        k = i % 2
        main.wait_event(in_ready[k])
        P.clear_plan_cache()
        with torch.no_grad():
            out = gnn(node_data={"input": h_buf[k]}, adjacency_lists=list(adj_buf[k]), edge_feature_data=[], node_to_graph_idx=n2g,
                      reference_node_ids={}, reference_node_graph_idx={}, num_graphs=batch.num_graphs).output_node_representations
        compute_done[k].record(main)
        if do_d2h:
            with torch.cuda.stream(d2h):
                d2h.wait_event(compute_done[k])
                out_host[k].copy_(out, non_blocking=True)
                d2h_done[k].record(d2h)
            out.record_stream(d2h)
        keep[k] = out
        prefetch(1 - k, do_h2d)
    main.wait_stream(d2h); main.wait_stream(h2d)
def timeit(**kw):
    run(4, **kw); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(20, **kw); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 20
prefetch(0); prefetch(1); torch.cuda.synchronize()
print("no copies      ", timeit(do_h2d=False, do_d2h=False))
print("h2d only       ", timeit(do_h2d=True, do_d2h=False))
print("d2h only       ", timeit(do_h2d=False, do_d2h=True))
print("h2d + d2h      ", timeit(do_h2d=True, do_d2h=True))
