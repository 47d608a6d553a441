"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ptgnn message-passing hot path.

A functional restatement, on stock PyTorch CPU ops, of the reference's algorithm for the path
named in BASELINE.json ``north_star``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (``cpu_baseline`` leg / ``--impl reference``) may import this module; the product
package ``ptgnn_b200`` never does, and fails loudly when its CUDA library is missing.

What each function follows (paths relative to /root/reference/ptgnn):

* ``scatter``                 -> ``neuralmodels/gnn/messagepassing/abstractmessagepassing.py:38-50`` and the
                                  third-party ``torch_scatter.scatter`` it calls (torch-scatter >= 2.0.5,
                                  ``setup.py:23``; CI pin 2.0.6).  torch-scatter's source is NOT under
                                  /root/reference; its published CPU semantics are restated here
                                   and pinned by known-answer tests
                                  (``tests/test_oracle_scatter_kat.py``).
* ``gated_layer_forward``     -> ``neuralmodels/gnn/messagepassing/gatedmessagepassing.py:37-69``
* ``mlp_layer_forward``       -> ``neuralmodels/gnn/messagepassing/mlpmessagepassing.py:68-117`` and
                                  ``neuralmodels/mlp.py:50-80``
* ``expand_adjacency``        -> ``neuralmodels/gnn/graphneuralnetwork.py:162-186``
* ``gnn_forward``             -> ``neuralmodels/gnn/graphneuralnetwork.py:121-134``
* ``edge_plan``               -> no reference counterpart: the canonical (stable, target-sorted) CSR edge
                                  plan that the CUDA path builds; defined here so that it can be checked
                                  bit-exactly.  It is the integer restatement of
                                  ``torch.cat([adj[1] ...])`` (gatedmessagepassing.py:46) + the
                                  index->row grouping performed inside ``torch_scatter.scatter``.

Pinning status: the reference ships no test, golden vector or fixture for this path
(``ptgnn/tests/simplemodel/test_model.py`` never imports a GNN class).  The oracle is therefore
pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container: the reference's own
``GatedMessagePassingLayer`` / ``MlpMessagePassingLayer`` / ``GraphNeuralNetwork`` classes are
imported unmodified from /root/reference (``oracle/refimport.py``) and their outputs on seeded inputs
are committed under ``tests/golden/`` by ``tests/golden/generate_golden.py``.  The one piece of
arithmetic that is not the reference's own code in those runs is ``torch_scatter`` (absent wheel,
restated in ``oracle/refstubs/torch_scatter``), which is what the KATs pin.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Adjacency = List[Tuple[torch.Tensor, torch.Tensor]]

REDUCE_OPS = ("sum", "mean", "max", "min")


# --------------------------------------------------------------------------------------------
# torch_scatter.scatter semantics
# --------------------------------------------------------------------------------------------
def scatter_with_arg(
    src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """``torch_scatter.scatter(src, index, dim=0, dim_size=dim_size, reduce=reduce)`` for a 2-D
    ``src [E, D]`` and 1-D int64 ``index [E]``.  Returns ``(out [dim_size, D], arg)`` where ``arg``
    (int64, ``E`` for untouched rows, first occurrence wins ties) is only produced for max/min."""
    assert src.dim() == 2 and index.dim() == 1 and index.shape[0] == src.shape[0]
    E, D = src.shape
    idx = index.reshape(E, 1).expand(E, D)
    if reduce in ("sum", "add"):
        out = torch.zeros(dim_size, D, dtype=src.dtype)
        return out.scatter_add_(0, idx, src), None
    if reduce == "mean":
        out = torch.zeros(dim_size, D, dtype=src.dtype).scatter_add_(0, idx, src)
        count = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, torch.ones(E, dtype=src.dtype))
        count.clamp_(min=1)
        return out / count.reshape(-1, 1), None
    if reduce in ("max", "min"):
        # torch_scatter: out starts at numeric_limits::lowest()/max(), is updated on a STRICT compare in
        # edge order (first occurrence wins ties; NaN, and a value equal to the initial one, never win);
        # rows that were never updated are set to 0 and keep arg == E.
        is_max = reduce == "max"
        init = torch.finfo(src.dtype).min if is_max else torch.finfo(src.dtype).max
        eligible = ~torch.isnan(src) & ((src > init) if is_max else (src < init))
        clean = torch.where(eligible, src, torch.full_like(src, init))
        out = torch.full((dim_size, D), init, dtype=src.dtype)
        out.scatter_reduce_(0, idx, clean, "amax" if is_max else "amin", include_self=True)
        winners = eligible & (clean == out.gather(0, idx))
        edge_ids = torch.arange(E, dtype=torch.int64).reshape(E, 1).expand(E, D)
        cand = torch.where(winners, edge_ids, torch.full_like(edge_ids, E))
        arg = torch.full((dim_size, D), E, dtype=torch.int64)
        arg.scatter_reduce_(0, idx, cand, "amin", include_self=True)
        out = torch.where(arg == E, torch.zeros_like(out), out)
        return out, arg
    raise ValueError(f"unknown reduce {reduce!r}")


def scatter(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str) -> torch.Tensor:
    return scatter_with_arg(src, index, dim_size, reduce)[0]


def aggregate_messages(messages: torch.Tensor, message_targets: torch.Tensor, num_nodes: int, aggregation_fn: str):
    """abstractmessagepassing.py:38-50 -- up-cast to fp32, scatter, cast back (AMP support)."""
    msg_dtype = messages.dtype
    return scatter(messages.to(torch.float32), message_targets, num_nodes, aggregation_fn).to(msg_dtype)


# --------------------------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------------------------
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """``nn.GRUCell`` forward (gate order r, z, n):  n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    return torch.gru_cell(x, h, w_ih, w_hh, b_ih, b_hh)


def gated_layer_forward(
    node_states: torch.Tensor,
    adjacency_lists: Adjacency,
    edge_features: Sequence[torch.Tensor],
    edge_weights: Sequence[torch.Tensor],
    gru_w_ih: torch.Tensor,
    gru_w_hh: torch.Tensor,
    gru_b_ih: torch.Tensor,
    gru_b_hh: torch.Tensor,
    aggregation_fn: str,
) -> torch.Tensor:
    """gatedmessagepassing.py:37-69 in eval mode (dropout = identity)."""
    assert len(adjacency_lists) == len(edge_weights)
    message_targets = torch.cat([adj[1] for adj in adjacency_lists])
    all_messages = []
    for (src, _tgt), feats, w in zip(adjacency_lists, edge_features, edge_weights):
        source_states = F.embedding(src, node_states)
        all_messages.append(F.linear(torch.cat([source_states, feats.to(source_states.dtype)], -1), w))
    agg = aggregate_messages(torch.cat(all_messages, 0), message_targets, node_states.shape[0], aggregation_fn)
    return gru_cell(agg, node_states, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh)


def mlp_apply(x: torch.Tensor, linear_weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """mlp.py:50-80 with use_biases=False, activation=ReLU, eval mode."""
    for i, w in enumerate(linear_weights):
        x = F.linear(x, w)
        if i + 1 < len(linear_weights):
            x = F.relu(x)
    return x


def mlp_layer_forward(
    node_states: torch.Tensor,
    adjacency_lists: Adjacency,
    edge_features: Sequence[torch.Tensor],
    edge_mlp_weights: Sequence[Sequence[torch.Tensor]],
    aggregation_fn: str,
    *,
    use_target_state_as_message_input: bool = True,
    message_activation: Optional[str] = "gelu",
    ln_weight: Optional[torch.Tensor] = None,
    ln_bias: Optional[torch.Tensor] = None,
    ln_eps: float = 1e-5,
    dense_weight: Optional[torch.Tensor] = None,
    dense_bias: Optional[torch.Tensor] = None,
    dense_activation: Optional[str] = "tanh",
) -> torch.Tensor:
    """mlpmessagepassing.py:68-117 in eval mode with a string aggregator."""
    assert len(adjacency_lists) == len(edge_mlp_weights)
    targets, messages = [], []
    for (src, tgt), feats, ws in zip(adjacency_lists, edge_features, edge_mlp_weights):
        targets.append(tgt)
        inp = F.embedding(src, node_states)
        if use_target_state_as_message_input:
            inp = torch.cat([inp, F.embedding(tgt, node_states)], -1)
        messages.append(mlp_apply(torch.cat([inp, feats.to(inp.dtype)], -1), ws))
    agg = aggregate_messages(torch.cat(messages, 0), torch.cat(targets, 0), node_states.shape[0], aggregation_fn)
    if message_activation == "gelu":
        agg = F.gelu(agg)
    elif message_activation == "relu":
        agg = F.relu(agg)
    elif message_activation is not None:
        raise ValueError(message_activation)
    if ln_weight is not None:
        agg = F.layer_norm(agg, (agg.shape[-1],), ln_weight, ln_bias, ln_eps)
    if dense_weight is not None:
        agg = F.linear(agg, dense_weight, dense_bias)
        if dense_activation == "tanh":
            agg = torch.tanh(agg)
        elif dense_activation == "relu":
            agg = F.relu(agg)
        elif dense_activation is not None:
            raise ValueError(dense_activation)
    return agg


def egc_layer_forward(
    node_states: torch.Tensor,
    adjacency_lists: Adjacency,
    bases_weights: Sequence[torch.Tensor],
    coeff_weight: torch.Tensor,
    coeff_bias: torch.Tensor,
    aggregation_fn: str,
    num_heads: int,
    num_bases: int,
) -> torch.Tensor:
    """egcmessagepassing.py:54-91 in eval mode (dropout = identity).  ``bases_weights[t]`` is ``[num_bases * out, H]``."""
    assert len(adjacency_lists) == len(bases_weights)
    out_dim = bases_weights[0].shape[0] // num_bases
    node_weights = F.linear(node_states, coeff_weight, coeff_bias).reshape(-1, num_heads, num_bases, 1)         # :64-66
    targets, messages = [], []
    for (src, tgt), w in zip(adjacency_lists, bases_weights):                                                   # :69-83
        targets.append(tgt)
        messages.append(F.linear(F.embedding(src, node_states), w).reshape(-1, num_heads, num_bases, out_dim // num_heads))
    msg = torch.cat(messages, 0)
    # torch_scatter.scatter on dim 0 of a 4-D tensor (:85-89) == the 2-D scatter on the flattened trailing dimensions
    agg = aggregate_messages(msg.reshape(msg.shape[0], -1), torch.cat(targets, 0), node_states.shape[0], aggregation_fn)
    agg = agg.reshape(-1, num_heads, num_bases, out_dim // num_heads)
    return (agg * node_weights).sum(dim=-2).reshape(-1, out_dim)                                                # :90


# --------------------------------------------------------------------------------------------
# Container bookkeeping
# --------------------------------------------------------------------------------------------
def expand_adjacency(
    adjacency_lists: Adjacency, num_nodes: int, introduce_backwards_edges: bool, add_self_edges: bool
) -> Adjacency:
    """graphneuralnetwork.py:172-186 -- returns a NEW list (the reference mutates the caller's)."""
    out = list(adjacency_lists)
    if introduce_backwards_edges:
        out += [(t, f) for f, t in adjacency_lists]
    if add_self_edges:
        ident = torch.arange(num_nodes, dtype=torch.int64)
        out.append((ident, ident))
    return out


def gnn_forward(node_states: torch.Tensor, adjacency_lists: Adjacency, layers: Sequence[dict]) -> List[torch.Tensor]:
    """graphneuralnetwork.py:121-131 -- ``layers`` is a list of dicts
    ``{"kind": "gated"|"mlp", ...kwargs of the matching *_layer_forward}``; returns all states."""
    feats = [torch.empty(a[0].shape[0], 0) for a in adjacency_lists]
    states = [node_states]
    for spec in layers:
        spec = dict(spec)
        kind = spec.pop("kind")
        fn = gated_layer_forward if kind == "gated" else mlp_layer_forward
        states.append(fn(states[-1], adjacency_lists, feats, **spec))
    return states


# --------------------------------------------------------------------------------------------
# Edge plan (integer bookkeeping; bit-exact contract with the CUDA plan builder)
# --------------------------------------------------------------------------------------------
def edge_plan(adjacency_lists: Adjacency, num_nodes: int) -> Dict[str, np.ndarray]:
    """Canonical target-sorted CSR of the concatenated per-type edge lists.

    * edge id ``e`` = position in ``cat(types)`` order (the order the reference feeds to scatter);
    * ``perm[j]``   = edge id at sorted position ``j`` -- STABLE sort by target, so edges of one target
                      keep the reference's summation order and ties resolve to the first occurrence;
    * ``pos[e]``    = inverse of ``perm``;
    * ``row_ptr``   = CSR offsets over targets (``row_ptr[v+1]-row_ptr[v]`` = in-degree of ``v``);
    * ``src_sorted[j]``, ``etype_sorted[j]`` = source node / edge type of the edge at position ``j``;
    * ``src32``, ``tgt32`` = the int64 lists down-converted, in edge-id order; ``type_off`` = type offsets.
    """
    T = len(adjacency_lists)
    counts = np.array([int(a[0].shape[0]) for a in adjacency_lists], dtype=np.int64)
    type_off = np.zeros(T + 1, dtype=np.int64)
    np.cumsum(counts, out=type_off[1:])
    E = int(type_off[-1])
    src = np.concatenate([a[0].numpy() for a in adjacency_lists]) if T else np.zeros(0, np.int64)
    tgt = np.concatenate([a[1].numpy() for a in adjacency_lists]) if T else np.zeros(0, np.int64)
    assert E < 2**31 and num_nodes < 2**31
    if E:
        assert src.min() >= 0 and src.max() < num_nodes and tgt.min() >= 0 and tgt.max() < num_nodes
    etype = np.repeat(np.arange(T, dtype=np.int64), counts)
    perm = np.argsort(tgt, kind="stable")
    pos = np.empty(E, dtype=np.int64)
    pos[perm] = np.arange(E)
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(tgt, minlength=num_nodes), out=row_ptr[1:])
    return {
        "row_ptr": row_ptr.astype(np.int32),
        "perm": perm.astype(np.int32),
        "pos": pos.astype(np.int32),
        "src_sorted": src[perm].astype(np.int32),
        "etype_sorted": etype[perm].astype(np.uint8),
        "src32": src.astype(np.int32),
        "tgt32": tgt.astype(np.int32),
        "type_off": type_off.astype(np.int32),
    }


def block_plan(adjacency_lists: Adjacency, num_nodes: int, block_targets: int) -> Dict[str, np.ndarray]:
    """Edge order of the fused aggregation kernel (``ptgnn_b200_block_plan_build``; no reference counterpart): the
    concatenated edges sorted, STABLY, by (target block, edge type, target) with blocks of ``block_targets`` consecutive
    nodes.  Inside one target the edges therefore keep the reference's scatter order (type-major, then list order).

    * ``group_off[b * T + t]`` = sorted position of the first edge of (block b, type t); last entry = E
    * ``src_f[j]``, ``tl_f[j]`` = source node / (target - block start) of the edge at sorted position ``j``
    """
    T = len(adjacency_lists)
    counts = np.array([int(a[0].shape[0]) for a in adjacency_lists], dtype=np.int64)
    src = np.concatenate([a[0].numpy() for a in adjacency_lists]) if T else np.zeros(0, np.int64)
    tgt = np.concatenate([a[1].numpy() for a in adjacency_lists]) if T else np.zeros(0, np.int64)
    etype = np.repeat(np.arange(T, dtype=np.int64), counts)
    B = int(block_targets)
    nblk = (num_nodes + B - 1) // B
    blk, tl = tgt // B, tgt % B
    key = (blk * T + etype) * B + tl
    perm = np.argsort(key, kind="stable")
    group_off = np.zeros(nblk * T + 1, dtype=np.int64)
    np.cumsum(np.bincount(blk * T + etype, minlength=nblk * T), out=group_off[1:])
    return {"group_off": group_off.astype(np.int32), "src_f": src[perm].astype(np.int32), "tl_f": tl[perm].astype(np.uint8),
            "perm": perm.astype(np.int32)}
