"""The edge plan's integer contract: python oracle == C oracle (bit-exact) + structural properties."""
import numpy as np
import pytest
import torch

from helpers import random_adjacency
from oracle import coracle
from oracle import ptgnn_oracle as O

CASES = [(10, [0]), (1, [5]), (50, [30, 0, 12]), (257, [1000, 3, 0, 511]), (4000, [9000, 7000, 1])]


@pytest.mark.parametrize("n,counts", CASES)
def test_plan_python_vs_c(n, counts):
    gen = torch.Generator().manual_seed(n)
    adj = random_adjacency(gen, n, counts)
    p = O.edge_plan(adj, n)
    c = coracle.edge_plan([(s.numpy(), t.numpy()) for s, t in adj], n)
    for k in ("row_ptr", "perm", "pos", "src_sorted", "etype_sorted"):
        assert np.array_equal(p[k], c[k]), k


def test_plan_properties():
    gen = torch.Generator().manual_seed(3)
    n, counts = 300, [700, 0, 450]
    adj = random_adjacency(gen, n, counts)
    p = O.edge_plan(adj, n)
    E = sum(counts)
    tgt = np.concatenate([a[1].numpy() for a in adj])
    assert p["row_ptr"][0] == 0 and p["row_ptr"][-1] == E
    assert np.array_equal(np.diff(p["row_ptr"]), np.bincount(tgt, minlength=n))
    assert np.array_equal(np.sort(p["perm"]), np.arange(E))           # permutation
    assert np.array_equal(p["perm"][p["pos"]], np.arange(E))            # inverse
    sorted_tgt = tgt[p["perm"]]
    assert np.all(np.diff(sorted_tgt) >= 0)                              # sorted by target
    same = np.diff(sorted_tgt) == 0
    assert np.all(np.diff(p["perm"])[same] > 0)                          # stable within a target
    assert np.array_equal(p["etype_sorted"], np.repeat(np.arange(3), counts)[p["perm"]])


def test_plan_rejects_out_of_range():
    adj = [(torch.tensor([0, 5]), torch.tensor([1, 2]))]
    with pytest.raises(AssertionError):
        O.edge_plan(adj, 3)
    with pytest.raises(ValueError):
        coracle.edge_plan([(np.array([0, 5]), np.array([1, 2]))], 3)
