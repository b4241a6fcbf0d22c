"""Learned item-item graph of LATTICE / MICRO on the sparse form.

Reference (Model/LATTICE.py:45-74,142-164): every epoch the model rebuilds a DENSE [N, N] cosine matrix per
modality, keeps the top-k of each row (``scatter_`` into zeros), mixes the modalities with softmax weights,
normalises with two dense ``diagflat`` GEMMs, blends with the fixed original graph and multiplies the dense
result with the item table.  Gradients reach the feature transforms through the kept similarity VALUES.

Here the same operator is O(N*k*d): the neighbour INDICES come from the kNN kernel (csrc/rank_topk.cu /
csrc/tc_candidates.cu, no gradient, like ``torch.topk`` indices), the VALUES are recomputed on that pattern
by the SDDMM kernel (differentiable in the features), the normalisation is a row/column scaling and the
product is the SpMM kernel, differentiable in the edge values (SDDMM again) and in the table.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .knn import knn_topk
from .propagate import CSRGraph, _spmm_raw, csr_from_coo


class KnnPattern:
    """Row-regular sparsity pattern: row i has the k columns ``ind[i]`` (CSR without a sort)."""

    def __init__(self, ind: torch.Tensor, n_cols: Optional[int] = None):
        _lib.require_cuda(ind, "ind")
        self.n, self.k = int(ind.shape[0]), int(ind.shape[1])
        self.n_cols = int(n_cols) if n_cols is not None else self.n
        self.ind = ind.to(torch.int64).contiguous()
        self.rowptr = torch.arange(self.n + 1, device=ind.device, dtype=torch.int64) * self.k
        self.col = self.ind.flatten().to(torch.int32).contiguous()
        self.rows = torch.arange(self.n, device=ind.device, dtype=torch.int64).repeat_interleave(self.k)

    def graph(self, vals: torch.Tensor) -> CSRGraph:
        return CSRGraph(self.rowptr, self.col, vals.detach().flatten().to(torch.float32).contiguous(), self.n,
                        self.n_cols).plan_hubs()

    def transposed_graph(self, vals: torch.Tensor) -> CSRGraph:
        return csr_from_coo(self.col.to(torch.int64), self.rows, vals.detach().flatten().to(torch.float32), self.n_cols, self.n)


def _sddmm_raw(p: KnnPattern, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty(p.n * p.k, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.chaorec_sddmm_csr_f32(_lib.ptr(p.rowptr), _lib.ptr(p.col), _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), p.n,
                                       int(a.shape[1]), _lib.stream_ptr(a.device))
    _lib.check(rc, "chaorec_sddmm_csr_f32")
    return out.view(p.n, p.k)


def _spmm_vals(g: CSRGraph, x: torch.Tensor) -> torch.Tensor:
    y = torch.empty(g.M, x.shape[1], dtype=torch.float32, device=x.device)
    _spmm_raw(g, x.contiguous(), y, None, 0.0, 0)
    return y


class _SddmmFn(torch.autograd.Function):
    """vals[i, j] = <a[i], b[ind[i, j]]>; backward = two SpMMs on the pattern weighted by the incoming gradient."""

    @staticmethod
    def forward(ctx, a, b, p: KnnPattern):
        ctx.p = p
        ctx.save_for_backward(a, b)
        return _sddmm_raw(p, a.detach(), b.detach())

    @staticmethod
    def backward(ctx, gv):
        a, b = ctx.saved_tensors
        p = ctx.p
        ga = _spmm_vals(p.graph(gv), b) if ctx.needs_input_grad[0] else None
        gb = _spmm_vals(p.transposed_graph(gv), a) if ctx.needs_input_grad[1] else None
        return ga, gb, None


class _SpmmValFn(torch.autograd.Function):
    """y = A(vals) @ x on a KnnPattern, differentiable in ``vals`` (SDDMM) and ``x`` (transposed SpMM)."""

    @staticmethod
    def forward(ctx, vals, x, p: KnnPattern):
        ctx.p = p
        ctx.save_for_backward(vals, x)
        return _spmm_vals(p.graph(vals), x.detach())

    @staticmethod
    def backward(ctx, gy):
        vals, x = ctx.saved_tensors
        p = ctx.p
        gy = gy.contiguous()
        gvals = _sddmm_raw(p, gy, x.detach()) if ctx.needs_input_grad[0] else None
        gx = _spmm_vals(p.transposed_graph(vals), gy) if ctx.needs_input_grad[1] else None
        return gvals, gx, None


def pattern_matmul(p: KnnPattern, vals: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``scatter_(-1, ind, vals) @ x`` without the dense matrix; differentiable in ``vals`` and ``x``."""
    _lib.require_cuda(x, "x")
    return _SpmmValFn.apply(vals, x, p)


def knn_graph(feats: torch.Tensor, k: int) -> Tuple[KnnPattern, torch.Tensor]:
    """``build_knn_neighbourhood(build_sim(feats), k)`` (Model/LATTICE.py:45-52,67-74) as (pattern, values):
    indices from the kNN kernel on the detached features, values = cosine similarities on that pattern,
    differentiable in ``feats``."""
    _lib.require_cuda(feats, "feats")
    _, ind = knn_topk(feats.detach(), k)
    p = KnnPattern(ind)
    xn = feats.div(torch.norm(feats, p=2, dim=-1, keepdim=True))
    return p, _SddmmFn.apply(xn, xn, p)


def normalized_values(p: KnnPattern, vals: torch.Tensor) -> torch.Tensor:
    """``compute_normalized_laplacian`` (Model/LATTICE.py:56-64) on the pattern: d^-1/2[i] * v * d^-1/2[col]."""
    d_inv_sqrt = torch.pow(vals.sum(-1), -0.5)
    d_inv_sqrt = torch.where(torch.isinf(d_inv_sqrt), torch.zeros_like(d_inv_sqrt), d_inv_sqrt)
    return d_inv_sqrt.unsqueeze(1) * vals * d_inv_sqrt[p.ind]


class ItemGraph:
    """``item_adj`` of Model/LATTICE.py:142-160 as an operator:

        item_adj = (1 - lambda) * D^-1/2 (sum_m w_m A_m) D^-1/2 + lambda * sum_m w_m O_m

    A_m = current kNN graph of modality m (values differentiable), O_m = the fixed normalised original graph,
    w = softmax(modal_weight), D = row sums of the mixed learned graph.  ``mm(h)`` replaces
    ``torch.mm(self.item_adj, h)`` (Model/LATTICE.py:163-164); ``detach()`` mirrors ``self.item_adj.detach()``.
    """

    def __init__(self, learned: Sequence[Tuple[KnnPattern, torch.Tensor]], original: Sequence[Tuple[KnnPattern, torch.Tensor]],
                 weight: torch.Tensor, lambda_coeff: float):
        self.learned, self.original, self.weight, self.lambda_coeff = list(learned), list(original), weight, float(lambda_coeff)
        rowsum = None
        for m, (p, v) in enumerate(self.learned):
            rs = weight[m] * v.sum(-1)
            rowsum = rs if rowsum is None else rowsum + rs
        d = torch.pow(rowsum, -0.5)
        self.d_inv_sqrt = torch.where(torch.isinf(d), torch.zeros_like(d), d)

    def mm(self, h: torch.Tensor) -> torch.Tensor:
        hs = self.d_inv_sqrt.unsqueeze(1) * h
        out = None
        for m, (p, v) in enumerate(self.learned):
            t = self.weight[m] * pattern_matmul(p, v, hs)
            out = t if out is None else out + t
        out = (1.0 - self.lambda_coeff) * (self.d_inv_sqrt.unsqueeze(1) * out)
        for m, (p, v) in enumerate(self.original):
            out = out + self.lambda_coeff * self.weight[m] * pattern_matmul(p, v, h)
        return out

    def detach(self) -> "ItemGraph":
        return ItemGraph([(p, v.detach()) for p, v in self.learned], [(p, v.detach()) for p, v in self.original],
                         self.weight.detach(), self.lambda_coeff)


def original_graph(feats: torch.Tensor, k: int) -> Tuple[KnnPattern, torch.Tensor]:
    """The fixed per-modality graph of ``LATTICE.__init__`` (Model/LATTICE.py:117-123): kNN of the raw features,
    normalised, no gradient."""
    with torch.no_grad():
        p, v = knn_graph(feats.detach(), k)
        return p, normalized_values(p, v)
