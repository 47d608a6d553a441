"""Sharded execution on the GPU: graph shards and node-range shards must reproduce the unsharded result bit-for-bit
(same messages, same per-target accumulation order).  The 2-GPU NCCL test runs only when two devices are visible."""
import os

import pytest
import torch

from helpers import random_adjacency

pytestmark = pytest.mark.gpu


def _layers(T, H):
    import ptgnn_b200 as P

    torch.manual_seed(0)
    return [P.GatedMessagePassingLayer(H, H, T, "sum").cuda().eval(), P.MlpMessagePassingLayer(H, H, H, T, "max").cuda().eval()]


def test_graph_shards_match_unsharded():
    from ptgnn_b200 import sharding
    from ptgnn_b200.synthetic import block_diagonal_batch

    b = block_diagonal_batch(9, 300, 9000, (0.5, 0.3, 0.2), seed=4)
    H = 64
    layers = _layers(3, H)
    h = torch.randn(b.num_nodes, H, generator=torch.Generator().manual_seed(1)).cuda()
    adj = [(s.cuda(), t.cuda()) for s, t in b.adjacency_lists]
    with torch.no_grad():
        whole = layers[1](layers[0](h, adj), adj)
        parts = []
        for r in sharding.partition_by_graph(b.node_to_graph_idx, b.adjacency_lists, 4):
            s = sharding.shard_graphs(b.node_to_graph_idx.cuda(), adj, r)
            x = h[s.node_lo:s.node_hi].contiguous()
            parts.append(layers[1](layers[0](x, s.adjacency_lists), s.adjacency_lists))
    assert torch.equal(torch.cat(parts), whole)


@pytest.mark.parametrize("world", [2, 3])
def test_row_shards_match_unsharded_single_device(world):
    """All shards executed one after the other on one GPU (the collective is replaced by slicing the full state)."""
    from ptgnn_b200 import sharding
    from ptgnn_b200.synthetic import single_random_graph

    g = single_random_graph(5001, 40000, 3, seed=2)
    H = 64
    layers = _layers(3, H)
    h = torch.randn(g.num_nodes, H, generator=torch.Generator().manual_seed(1)).cuda()
    adj = [(s.cuda(), t.cuda()) for s, t in g.adjacency_lists]
    with torch.no_grad():
        ref1 = layers[0](h, adj)
        ref2 = layers[1](ref1, adj)
        shards = [sharding.row_shard(g.num_nodes, adj, world, r) for r in range(world)]
        out1 = torch.cat([layers[0](h[s.lo:s.hi].contiguous(), s.adjacency_lists, gather_states=h) for s in shards])
        out2 = torch.cat([layers[1](out1[s.lo:s.hi].contiguous(), s.adjacency_lists, gather_states=out1) for s in shards])
    assert torch.equal(out1, ref1) and torch.equal(out2, ref2)


def _nccl_worker(rank, world, port, ok):
    import torch.distributed as dist

    import ptgnn_b200 as P
    from ptgnn_b200 import sharding
    from ptgnn_b200.synthetic import single_random_graph

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = single_random_graph(20001, 150000, 4, seed=5)
        H = 128
        torch.manual_seed(0)
        layers = [P.GatedMessagePassingLayer(H, H, 4, "sum").cuda().eval() for _ in range(3)]
        h = torch.randn(g.num_nodes, H, generator=torch.Generator().manual_seed(1)).cuda()
        adj = [(s.cuda(), t.cuda()) for s, t in g.adjacency_lists]
        shard = sharding.row_shard(g.num_nodes, adj, world, rank)
        loop = sharding.RowShardedLayerLoop(shard)
        with torch.no_grad():
            out = loop.run(h[shard.lo:shard.hi].contiguous(),
                           [lambda own, full, a, L=L: L(own, a, gather_states=full) for L in layers])
            full = loop.all_gather_states(out)
            ref = h
            for L in layers:
                ref = L(ref, adj)
        if rank == 0:
            ok.put(bool(torch.equal(full, ref)))
    finally:
        dist.destroy_process_group()


def test_row_sharded_loop_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ok = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ok.get(timeout=10)
