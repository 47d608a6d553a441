"""GPU edge plan vs the oracle: integer bookkeeping must be BIT-EXACT (north_star)."""
import numpy as np
import pytest
import torch

from helpers import random_adjacency
from oracle import ptgnn_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    (1, [5]), (10, [0, 3]), (50, [30, 0, 12]), (257, [1000, 3, 0, 511]), (4000, [9000, 7000, 1]),
    (70000, [200000, 150000, 1, 0, 99999]),   # > 2 radix passes (17 bits), multi-block scans
    (300, [60000]),                            # heavy duplication: long per-target segments
]


def _gpu_plan(adj, n):
    import ptgnn_b200 as P

    plan = P.EdgePlan([(s.cuda(), t.cuda()) for s, t in adj], n, validate=True)
    return plan


@pytest.mark.parametrize("n,counts", CASES)
def test_plan_bit_exact(n, counts):
    gen = torch.Generator().manual_seed(n + len(counts))
    adj = random_adjacency(gen, n, counts)
    ref = O.edge_plan(adj, n)
    plan = _gpu_plan(adj, n)
    for k in ("row_ptr", "perm", "pos", "src_sorted", "etype_sorted", "src32", "tgt32"):
        got = getattr(plan, k).cpu().numpy()
        assert got.dtype == ref[k].dtype, k
        assert np.array_equal(got, ref[k]), f"{k} differs for N={n}, counts={counts}"
    assert plan.type_off == ref["type_off"].tolist()


def test_plan_hub_and_block_diagonal():
    from ptgnn_b200.synthetic import block_diagonal_batch

    b = block_diagonal_batch(6, 500, 6000, (0.5, 0.3, 0.2), seed=9)
    adj = list(b.adjacency_lists)
    adj.append((torch.arange(b.num_nodes), torch.zeros(b.num_nodes, dtype=torch.int64)))  # node 0 = hub of degree N
    ref = O.edge_plan(adj, b.num_nodes)
    plan = _gpu_plan(adj, b.num_nodes)
    for k in ("row_ptr", "perm", "pos", "src_sorted", "etype_sorted"):
        assert np.array_equal(getattr(plan, k).cpu().numpy(), ref[k]), k


def test_plan_empty_edge_set():
    plan = _gpu_plan([(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64))], 17)
    assert plan.num_edges == 0 and torch.equal(plan.row_ptr.cpu(), torch.zeros(18, dtype=torch.int32))


def test_plan_reports_out_of_range_indices():
    import ptgnn_b200 as P

    adj = [(torch.tensor([0, 9]).cuda(), torch.tensor([1, -1]).cuda())]
    with pytest.raises(IndexError):
        P.EdgePlan(adj, 5, validate=True)


def test_plan_is_deterministic_and_cached():
    import ptgnn_b200 as P

    gen = torch.Generator().manual_seed(0)
    adj = [(s.cuda(), t.cuda()) for s, t in random_adjacency(gen, 5000, [40000, 20000])]
    a, b = P.EdgePlan(adj, 5000), P.EdgePlan(adj, 5000)
    assert torch.equal(a.perm, b.perm) and torch.equal(a.row_ptr, b.row_ptr)
    P.clear_plan_cache()
    c1 = P.plan_for(adj, 5000)
    assert P.plan_for(adj, 5000) is c1
    adj[0][0].add_(0)  # in-place op bumps the version counter -> cache miss
    assert P.plan_for(adj, 5000) is not c1
