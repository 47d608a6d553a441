"""Backward of ``GatedMessagePassingLayer`` (first cut): makes the layer usable under ``torch.autograd``
(`/root/reference/ptgnn/baseneuralmodel/trainer.py:221-236` calls ``loss.backward()`` through the layers).

The forward is the unchanged native forward (fused aggregation + GRU kernels).  The backward re-computes what it needs and keeps the
edge-sized work on the native kernels:

* aggregate re-computation: ``edge_messages`` + ``segment_reduce`` kernels (with the arg-max edge ids for max / min);
* GRU: gate pre-activations and the two input-gradient products on the native dense kernel (``composed.linear``), the gate
  derivatives as pointwise torch ops;
* ``d h_src``: sum / mean -- the SAME gather -> Linear -> segmented-reduce kernels run on the transposed graph (edges reversed,
  weights ``W_t^T``, states = ``d agg``); max / min -- the routed message gradients times ``W_t`` on the dense kernel, then the
  native scatter-add by source;
* parameter gradients (``dW_t``, ``dW_ih``, ``dW_hh``): plain ``[out, rows] x [rows, in]`` GEMMs with a huge K (rows = edges or
  nodes) -- library GEMMs (``torch.matmul`` = cuBLAS), as are the bias column sums.

fp32 states only; training-mode dropout (a per-edge mask on the gathered rows, gatedmessagepassing.py:59) and edge features
still raise.  Parity: ``tests/test_gpu_backward.py`` against ``torch.autograd`` through the CPU oracle (1e-4).
"""
from typing import List, Tuple

import torch

from . import _native as N
from . import composed as C
from .edgeplan import plan_for
from .scatter import scatter_sum


def needs_grad(module: torch.nn.Module, node_states: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (node_states.requires_grad or any(p.requires_grad for p in module.parameters()))


class _GatedLayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, adjacency_lists, reduce_name, h, w_ih, w_hh, b_ih, b_hh, *weights):
        with torch.no_grad():
            out = layer(h.detach(), adjacency_lists)
        ctx.save_for_backward(h, w_ih, w_hh, b_ih, b_hh, *weights)
        ctx.adjacency_lists = adjacency_lists
        ctx.reduce_name = reduce_name
        return out

    @staticmethod
    def backward(ctx, grad_out):
        h, w_ih, w_hh, b_ih, b_hh, *weights = ctx.saved_tensors
        adj: List[Tuple[torch.Tensor, torch.Tensor]] = ctx.adjacency_lists
        reduce_name = ctx.reduce_name
        reduce = N.REDUCE[reduce_name]
        g = grad_out.contiguous().float()
        h = h.detach().contiguous()
        num_nodes, H = h.shape
        W = [w.detach().contiguous() for w in weights]
        w_ih, w_hh = w_ih.detach().contiguous(), w_hh.detach().contiguous()
        plan = plan_for(adj, num_nodes)
        E = plan.num_edges

        # ---- 1. re-compute the aggregate (and, for max / min, which edge won each (target, feature))
        msg = C.edge_messages(plan, h, None, W, False)                       # [E, D], cat(types) order
        arg = None
        if reduce_name in ("max", "min"):
            agg, arg = C.segment_reduce(msg, plan, reduce, return_arg=True)
        else:
            agg = C.segment_reduce(msg, plan, reduce)
        del msg

        # ---- 2. GRUCell backward (gate order r, z, n; torch.nn.GRUCell)
        gi = C.linear(agg, w_ih, b_ih.detach())
        gh = C.linear(h, w_hh, b_hh.detach())
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        d_n_pre = g * (1.0 - z) * (1.0 - n * n)
        d_z_pre = g * (h - n) * z * (1.0 - z)
        d_r_pre = d_n_pre * h_n * r * (1.0 - r)
        d_gi = torch.cat([d_r_pre, d_z_pre, d_n_pre], dim=1)                 # [N, 3H]
        d_gh = torch.cat([d_r_pre, d_z_pre, d_n_pre * r], dim=1)
        d_h = g * z + C.linear(d_gh, w_hh.t().contiguous())                  # direct path + through W_hh
        d_agg = C.linear(d_gi, w_ih.t().contiguous())                        # [N, D]
        d_w_ih = d_gi.t() @ agg                                              # K = num_nodes: library GEMMs
        d_w_hh = d_gh.t() @ h
        d_b_ih, d_b_hh = d_gi.sum(dim=0), d_gh.sum(dim=0)

        # ---- 3. aggregation + per-type Linear backward
        d_W = []
        if reduce_name in ("sum", "mean"):
            if reduce_name == "mean":
                cnt = (plan.row_ptr[1:] - plan.row_ptr[:-1]).clamp(min=1).to(torch.float32)
                d_agg = d_agg / cnt[:, None]
            for (src, tgt), w in zip(adj, W):
                d_W.append(d_agg.index_select(0, tgt).t() @ h.index_select(0, src) if src.numel() else torch.zeros_like(w))
            if E > 0:
                # d h_src[u] = sum over edges (u -> v, type t) of W_t^T d_agg[v]: the forward's aggregation on the transposed graph
                rev = [(tgt, src) for src, tgt in adj]
                rplan = plan_for(rev, num_nodes)
                back = C.edge_messages(rplan, d_agg.contiguous(), None, [w.t().contiguous() for w in W], False)   # [E, H]
                d_h = d_h + C.segment_reduce(back, rplan, N.REDUCE["sum"])
        else:
            D = d_agg.shape[1]
            d_msg = torch.zeros(E + 1, D, dtype=torch.float32, device=h.device)      # row E takes the empty targets' sentinel
            d_msg.scatter_(0, arg, d_agg)                                            # each (edge, feature) has one target: no collisions
            lo = 0
            for (src, tgt), w in zip(adj, W):
                e_t = src.numel()
                part = d_msg[lo:lo + e_t]
                lo += e_t
                if e_t == 0:
                    d_W.append(torch.zeros_like(w))
                    continue
                d_W.append(part.t() @ h.index_select(0, src))
                d_h = d_h + scatter_sum(C.linear(part.contiguous(), w.t().contiguous()), src, dim=0, dim_size=num_nodes)
        return (None, None, None, d_h, d_w_ih, d_w_hh, d_b_ih, d_b_hh, *d_W)


def gated_forward_with_grad(layer, node_states, adjacency_lists, reduce_name, w_ih, w_hh, b_ih, b_hh, weights):
    return _GatedLayerFunction.apply(layer, adjacency_lists, reduce_name, node_states, w_ih, w_hh, b_ih, b_hh, *weights)
