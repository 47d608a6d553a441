"""Host-side multi-GPU logic on CPU: graph-granular partitioning, node-range shards, and the per-layer all-gather loop
run with world_size = 2 over the `gloo` backend (the compute inside each rank is the CPU oracle)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from helpers import gated_oracle_args
from oracle import ptgnn_oracle as O
from ptgnn_b200 import sharding
from ptgnn_b200.synthetic import block_diagonal_batch, single_random_graph


def _weights(T, H, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(edge_weights=[torch.randn(H, H, generator=g) * 0.1 for _ in range(T)],
                gru_w_ih=torch.randn(3 * H, H, generator=g) * 0.1, gru_w_hh=torch.randn(3 * H, H, generator=g) * 0.1,
                gru_b_ih=torch.randn(3 * H, generator=g) * 0.01, gru_b_hh=torch.randn(3 * H, generator=g) * 0.01)


def _oracle_shard_layer(w, agg):
    """One gated layer on a node-range shard: sources index the gathered states, targets the owned rows."""
    def fn(own, gathered, adj):
        msgs = torch.cat([F.linear(F.embedding(s, gathered), wt) for (s, _), wt in zip(adj, w["edge_weights"])])
        a = O.aggregate_messages(msgs, torch.cat([t for _, t in adj]), own.shape[0], agg)
        return O.gru_cell(a, own, w["gru_w_ih"], w["gru_w_hh"], w["gru_b_ih"], w["gru_b_hh"])
    return fn


def test_partition_by_graph_covers_and_balances():
    b = block_diagonal_batch(10, 100, 3000, (0.6, 0.4), seed=3)
    ranges = sharding.partition_by_graph(b.node_to_graph_idx, b.adjacency_lists, 4)
    assert ranges[0][0] == 0 and ranges[-1][1] == 10
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(3))
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(sizes) <= 1                       # equal-cost graphs -> near-equal split
    shards = [sharding.shard_graphs(b.node_to_graph_idx, b.adjacency_lists, r) for r in ranges]
    assert sum(s.num_nodes for s in shards) == b.num_nodes
    for t in range(2):
        assert sum(s.adjacency_lists[t][0].shape[0] for s in shards) == b.adjacency_lists[t][0].shape[0]
    for s in shards:                                           # renumbered ids are local
        for src, tgt in s.adjacency_lists:
            if src.numel():
                assert 0 <= int(src.min()) and int(src.max()) < s.num_nodes and int(tgt.max()) < s.num_nodes
        assert int(s.node_to_graph_idx.min()) == 0 and int(s.node_to_graph_idx.max()) == s.num_graphs - 1


def test_graph_shards_reproduce_the_unsharded_result():
    b = block_diagonal_batch(6, 64, 1500, (0.5, 0.3, 0.2), seed=5)
    H, T = 16, 3
    w = _weights(T, H, 1)
    h = torch.randn(b.num_nodes, H, generator=torch.Generator().manual_seed(2))
    feats = [torch.empty(a[0].shape[0], 0) for a in b.adjacency_lists]
    whole = O.gated_layer_forward(h, b.adjacency_lists, feats, aggregation_fn="sum", **w)
    parts = []
    for r in sharding.partition_by_graph(b.node_to_graph_idx, b.adjacency_lists, 3):
        s = sharding.shard_graphs(b.node_to_graph_idx, b.adjacency_lists, r)
        f = [torch.empty(a[0].shape[0], 0) for a in s.adjacency_lists]
        parts.append(O.gated_layer_forward(h[s.node_lo:s.node_hi], s.adjacency_lists, f, aggregation_fn="sum", **w))
    assert torch.equal(torch.cat(parts), whole)                # empty halo: bit-identical


def test_shard_graphs_rejects_cross_shard_edges():
    g = single_random_graph(100, 400, 2, seed=1)
    with pytest.raises(ValueError):
        sharding.shard_graphs(torch.arange(100) // 50, g.adjacency_lists, (0, 1))


def test_row_shard_bookkeeping():
    g = single_random_graph(1001, 5000, 3, seed=7)
    shards = [sharding.row_shard(g.num_nodes, g.adjacency_lists, 4, r) for r in range(4)]
    assert [s.lo for s in shards] == [0, 251, 502, 753] and shards[-1].hi == 1001
    for t in range(3):
        assert sum(s.adjacency_lists[t][0].shape[0] for s in shards) == g.adjacency_lists[t][0].shape[0]
    for s in shards:
        for src, tgt in s.adjacency_lists:
            assert int(tgt.min()) >= 0 and int(tgt.max()) < s.num_local and int(src.max()) < g.num_nodes


def _row_sharded_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = single_random_graph(301, 2500, 3, seed=11)        # ONE connected graph, N not divisible by the world size
        H = 16
        ws = [_weights(3, H, 20 + l) for l in range(3)]
        h = torch.randn(g.num_nodes, H, generator=torch.Generator().manual_seed(3))
        shard = sharding.row_shard(g.num_nodes, g.adjacency_lists, world, rank)
        loop = sharding.RowShardedLayerLoop(shard)
        out = loop.run(h[shard.lo:shard.hi].clone(), [_oracle_shard_layer(w, "sum" if i % 2 == 0 else "max") for i, w in enumerate(ws)])
        full = loop.all_gather_states(out)
        if rank == 0:
            ref = h
            feats = [torch.empty(a[0].shape[0], 0) for a in g.adjacency_lists]
            for i, w in enumerate(ws):
                ref = O.gated_layer_forward(ref, g.adjacency_lists, feats, aggregation_fn="sum" if i % 2 == 0 else "max", **w)
            results.put((torch.equal(full, ref), float((full - ref).abs().max())))
    finally:
        dist.destroy_process_group()


def test_row_sharded_loop_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_row_sharded_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same, err = results.get(timeout=10)
    assert same, f"sharded result differs from the unsharded oracle (max abs {err:.3e})"
