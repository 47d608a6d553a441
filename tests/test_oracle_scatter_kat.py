"""Known-answer tests that pin the torch_scatter semantics the oracle restates.  torch-scatter's
source is not under /root/reference (third-party wheel, >=2.0.5), so these hand-computed cases are the anchor."""
import numpy as np
import pytest
import torch

from oracle import coracle
from oracle import ptgnn_oracle as O

SRC = torch.tensor([[1.0, -2.0], [3.0, -2.0], [0.5, 4.0], [3.0, 7.0], [-1.0, float("nan")]])
IDX = torch.tensor([2, 0, 2, 0, 3])
N = 5  # rows 1 and 4 receive nothing


def _both(reduce):
    py_out, py_arg = O.scatter_with_arg(SRC, IDX, N, reduce)
    c_out, c_arg = coracle.scatter(SRC.numpy(), IDX.numpy(), N, reduce)
    return (py_out, py_arg), (torch.from_numpy(c_out), None if c_arg is None else torch.from_numpy(c_arg))


def test_sum_kat():
    expect = torch.tensor([[6.0, 5.0], [0.0, 0.0], [1.5, 2.0], [-1.0, float("nan")], [0.0, 0.0]])
    for out, _ in _both("sum"):
        assert torch.equal(torch.nan_to_num(out, nan=-99.0), torch.nan_to_num(expect, nan=-99.0))


def test_mean_kat_empty_rows_are_zero_and_count_clamped():
    expect = torch.tensor([[3.0, 2.5], [0.0, 0.0], [0.75, 1.0], [-1.0, float("nan")], [0.0, 0.0]])
    for out, _ in _both("mean"):
        assert torch.equal(torch.nan_to_num(out, nan=-99.0), torch.nan_to_num(expect, nan=-99.0))


def test_max_kat_first_occurrence_wins_nan_never_wins_empty_is_zero():
    # row 0: col0 tie between edges 1 and 3 (3.0) -> arg 1; col1 max(-2, 7) = 7 -> arg 3
    # row 3: col1 only NaN -> never updated -> 0 with arg = E (=5)
    expect = torch.tensor([[3.0, 7.0], [0.0, 0.0], [1.0, 4.0], [-1.0, 0.0], [0.0, 0.0]])
    expect_arg = torch.tensor([[1, 3], [5, 5], [0, 2], [4, 5], [5, 5]])
    for out, arg in _both("max"):
        assert torch.equal(out, expect)
        assert torch.equal(arg, expect_arg)


def test_min_kat():
    expect = torch.tensor([[3.0, -2.0], [0.0, 0.0], [0.5, -2.0], [-1.0, 0.0], [0.0, 0.0]])
    expect_arg = torch.tensor([[1, 1], [5, 5], [2, 0], [4, 5], [5, 5]])
    for out, arg in _both("min"):
        assert torch.equal(out, expect)
        assert torch.equal(arg, expect_arg)


def test_values_equal_to_initial_never_win():
    lowest = torch.finfo(torch.float32).min
    src = torch.tensor([[lowest], [float("-inf")]])
    idx = torch.tensor([0, 1])
    out, arg = O.scatter_with_arg(src, idx, 2, "max")
    assert torch.equal(out, torch.zeros(2, 1)) and torch.equal(arg, torch.full((2, 1), 2))
    c_out, c_arg = coracle.scatter(src.numpy(), idx.numpy(), 2, "max")
    assert np.array_equal(c_out, np.zeros((2, 1), np.float32)) and np.array_equal(c_arg, np.full((2, 1), 2))


def test_empty_input():
    for r in O.REDUCE_OPS:
        out, _ = O.scatter_with_arg(torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64), 3, r)
        assert torch.equal(out, torch.zeros(3, 4))
        c_out, _ = coracle.scatter(np.zeros((0, 4), np.float32), np.zeros(0, np.int64), 3, r)
        assert np.array_equal(c_out, np.zeros((3, 4), np.float32))


@pytest.mark.parametrize("reduce", O.REDUCE_OPS)
def test_python_and_c_oracles_agree_on_random_input(reduce):
    gen = torch.Generator().manual_seed(7)
    src = torch.randn(500, 12, generator=gen)
    src[::37] = src[1::37][: src[::37].shape[0]]  # plant ties
    idx = torch.randint(0, 40, (500,), generator=gen)
    out, arg = O.scatter_with_arg(src, idx, 45, reduce)
    c_out, c_arg = coracle.scatter(src.numpy(), idx.numpy(), 45, reduce)
    if reduce in ("max", "min"):
        assert np.array_equal(out.numpy(), c_out) and np.array_equal(arg.numpy(), c_arg)
    else:  # same edge order, same fp32 adds
        np.testing.assert_allclose(out.numpy(), c_out, rtol=1e-6, atol=1e-6)
