"""GraphConvolution / GCN / GCN3 with the reference's module API (gcn/layers.py:8-41, gcn/models.py:8-46):
same constructors, parameter names, [in,out] weight layout, U(±1/√out) init and state-dict keys
(gc1.weight, gc1.bias, gc2.weight, gc2.bias), so checkpoints written by either load in the other.

forward(input, adj) = mm(input, W) → spmm(adj, ·) → + bias runs on the sm_100a kernels
(lt_gemm_f32, lt_spmm_csr_f32).  Autograd is supported so that the reference's training loop
(gcn_trainer.py:144-206) keeps working through these layers: backward is the same two kernels
with Âᵀ (built once per adjacency and cached).
"""
from __future__ import annotations

import math
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import adj as _adj
from . import ops

_CSR_CACHE = {}   # id(adjacency object) → (weakref to it, [Â, Âᵀ])


def _csr_pair(adj):
    """[Â, Âᵀ] as DeviceCSR for the `adj` argument of forward; cached per adjacency OBJECT (identity,
    not value: tensors do not hash/compare by content) and dropped when that object dies."""
    key = adj.indptr if isinstance(adj, ops.DeviceCSR) else adj
    hit = _CSR_CACHE.get(id(key))
    if hit is not None and hit[0]() is key:
        return hit[1]
    pair = [_adj.device_csr_of(adj), None]
    k = id(key)
    _CSR_CACHE[k] = (weakref.ref(key, lambda _r, k=k: _CSR_CACHE.pop(k, None)), pair)
    return pair


def _transpose(csr: ops.DeviceCSR) -> ops.DeviceCSR:
    n = csr.n
    rows = torch.repeat_interleave(torch.arange(n, device=csr.indptr.device), csr.indptr[1:] - csr.indptr[:-1])
    cols = csr.indices.long()
    order = torch.argsort(cols * n + rows)
    return ops.coo_to_csr(cols[order].contiguous(), rows[order].contiguous(), csr.values[order].contiguous(), n,
                          check=False)


class _GraphConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight, bias, pair):
        csr = pair[0]
        x = input.detach().to(torch.float32)
        support = ops.gemm(x, weight.detach())                     # layers.py:31
        out = ops.spmm(csr, support, bias=None if bias is None else bias.detach())   # layers.py:32-34
        ctx.pair = pair
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight.detach())
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        pair = ctx.pair
        if pair[1] is None:
            pair[1] = _transpose(pair[0])
        g = grad_out.contiguous().to(torch.float32)
        g_support = ops.spmm(pair[1], g)                           # Âᵀ·dY
        grad_in = grad_w = grad_b = None
        if ctx.needs_input_grad[0]:
            grad_in = ops.gemm(g_support, weight, trans_b=True)    # dX = dS·Wᵀ
        if ctx.needs_input_grad[1]:
            grad_w = ops.gemm(x.t().contiguous(), g_support)       # dW = Xᵀ·dS
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_b = g.sum(dim=0)
        return grad_in, grad_w, grad_b, None


class GraphConvolution(nn.Module):
    """Simple GCN layer (gcn/layers.py:8-41)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, input, adj):
        if not input.is_cuda:
            raise RuntimeError("linkteller_b200.GraphConvolution runs on a CUDA device only (no CPU fallback)")
        return _GraphConvFn.apply(input, self.weight, self.bias, _csr_pair(adj))

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class GCN(nn.Module):
    """gc1 → ReLU → dropout → gc2, raw logits (gcn/models.py:8-25)."""

    def __init__(self, nfeat, nhid, nclass, dropout):
        super().__init__()
        self.gc1 = GraphConvolution(nfeat, nhid)
        self.gc2 = GraphConvolution(nhid, nclass)
        self.dropout = dropout

    def forward(self, x, adj):
        if not self.training and not torch.is_grad_enabled():
            # inference: one fused C call (lt_gcn_forward), ReLU in the SpMM epilogue
            return ops.gcn_forward(_csr_pair(adj)[0], x.to(torch.float32), self.gc1.weight.detach(),
                                   self.gc1.bias.detach(), self.gc2.weight.detach(), self.gc2.bias.detach())
        x = F.relu(self.gc1(x, adj))
        x = F.dropout(x, self.dropout, training=self.training)
        return self.gc2(x, adj)


class GCN3(nn.Module):
    """gc1 → ReLU → dropout → gc2 → ReLU → dropout → gc3, raw logits (gcn/models.py:28-46; --n-layer 3)."""

    def __init__(self, nfeat, nhid1, nhid2, nclass, dropout):
        super().__init__()
        self.gc1 = GraphConvolution(nfeat, nhid1)
        self.gc2 = GraphConvolution(nhid1, nhid2)
        self.gc3 = GraphConvolution(nhid2, nclass)
        self.dropout = dropout

    def forward(self, x, adj):
        if not self.training and not torch.is_grad_enabled():
            return ops.gcn3_forward(_csr_pair(adj)[0], x.to(torch.float32), self.gc1.weight.detach(),
                                    self.gc1.bias.detach(), self.gc2.weight.detach(), self.gc2.bias.detach(),
                                    self.gc3.weight.detach(), self.gc3.bias.detach())
        x = F.relu(self.gc1(x, adj))
        x = F.dropout(x, self.dropout, training=self.training)
        x = F.relu(self.gc2(x, adj))
        x = F.dropout(x, self.dropout, training=self.training)
        return self.gc3(x, adj)
