"""Modality item-item kNN graph: drop-ins for the reference's three kNN conventions.

Reference call sites replaced (paths relative to the ChaoRec checkout):
  A  ``get_knn_adj_mat(mm_embeddings) -> (indices, sparse adj)``    Model/FREEDOM.py:111-140
  B  ``torch.topk(build_sim(x), k) -> (knn_val, knn_ind)`` + dense scatter / normalisation
                                                                    Model/LATTICE.py:46-75, Model/MGCN.py:16-69
  C  ``get_knn_adj_mat(x, k) -> knn_ind`` (self excluded)            gen_hypergraph_u_i.py:41-48

The N x N similarity matrix is never materialised: csrc/rank_topk.cu streams it through
registers and keeps k candidates per row.  Everything after the top-k (O(N*k) index and value
bookkeeping) is restated here with the reference's own torch operators.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def knn_topk(feats: torch.Tensor, k: int, exclude_self: bool = False,
             row_range: Optional[Tuple[int, int]] = None, stats: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Cosine top-k of every row (or of rows [lo, hi)) against all rows.

    Returns (val float32 [n, k] descending, idx int64 [n, k]); ties -> ascending column.
    """
    _lib.require_cuda(feats, "feats")
    lib = _lib.load()
    dev = feats.device
    feats = feats.detach().to(torch.float32).contiguous()
    N, D = feats.shape
    lo, hi = (0, N) if row_range is None else (int(row_range[0]), int(row_range[1]))
    out_val = torch.empty(hi - lo, k, dtype=torch.float32, device=dev)
    out_idx = torch.empty(hi - lo, k, dtype=torch.int64, device=dev)
    ws = _lib.workspace(lib.chaorec_knn_workspace_bytes(N, D, k), dev)
    with torch.cuda.device(dev):
        rc = lib.chaorec_knn_cosine_topk(_lib.ptr(feats), N, D, int(k), 1 if exclude_self else 0, lo, hi,
                                         _lib.ptr(out_val), _lib.ptr(out_idx), _lib.ptr(ws), ws.numel(),
                                         _lib.stream_ptr(dev))
    _lib.check(rc, "chaorec_knn_cosine_topk")
    if stats is not None:
        tensor = bool(lib.chaorec_uses_tensor_engine(hi - lo, N, D, int(k)))
        stats["engine"] = "tensor" if tensor else "exact"
        off = ((N * D * 4 + 255) // 256) * 256 + ((N * 4 + 255) // 256) * 256 + 64   # normalised rows, norms, counters
        stats["fallback_rows"] = int(ws[off:off + 4].view(torch.int32).item()) if tensor else 0
    return out_val, out_idx


def l2_normalize(feats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rows / ||row||_2, norms) with the canonical fp64-accumulated norm (Model/FREEDOM.py:114)."""
    _lib.require_cuda(feats, "feats")
    lib = _lib.load()
    feats = feats.detach().to(torch.float32).contiguous()
    out = torch.empty_like(feats)
    norms = torch.empty(feats.shape[0], dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        rc = lib.chaorec_l2_normalize_f32(_lib.ptr(feats), feats.shape[0], feats.shape[1], _lib.ptr(out),
                                          _lib.ptr(norms), _lib.stream_ptr(feats.device))
    _lib.check(rc, "chaorec_l2_normalize_f32")
    return out, norms


def knn_laplacian_ones(indices: torch.Tensor, adj_size) -> torch.Tensor:
    """Model/FREEDOM.py:130-140 in O(N*k): all-ones adjacency, (1e-7 + rowsum)^-1/2 on both ends."""
    n = adj_size[0]
    row_sum = 1e-7 + torch.bincount(indices[0], minlength=n)
    r_inv_sqrt = torch.pow(row_sum, -0.5)
    values = r_inv_sqrt[indices[0]] * r_inv_sqrt[indices[1]]
    return torch.sparse_coo_tensor(indices, values, adj_size)


def get_knn_adj_mat(mm_embeddings: torch.Tensor, knn_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Convention A, drop-in for ``FREEDOM.get_knn_adj_mat`` (Model/FREEDOM.py:111-128)."""
    _, knn_ind = knn_topk(mm_embeddings, knn_k, exclude_self=False)
    n = knn_ind.shape[0]
    indices0 = torch.arange(n, device=knn_ind.device).unsqueeze(1).expand(-1, knn_k)
    indices = torch.stack((torch.flatten(indices0), torch.flatten(knn_ind)), 0)
    return indices, knn_laplacian_ones(indices, torch.Size([n, n]))


def build_knn_neighbourhood(context: torch.Tensor, topk: int, dense: bool = True):
    """Convention B: ``build_knn_neighbourhood(build_sim(context), topk)`` (Model/LATTICE.py:46-52,68-75).

    ``dense=True`` returns the reference's dense N x N scatter of the kept similarities;
    ``dense=False`` returns (knn_val, knn_ind) for the O(N*k) consumers.
    """
    knn_val, knn_ind = knn_topk(context, topk, exclude_self=False)
    if not dense:
        return knn_val, knn_ind
    n = knn_ind.shape[0]
    return torch.zeros(n, n, dtype=knn_val.dtype, device=knn_val.device).scatter_(-1, knn_ind, knn_val)


def knn_sym_laplacian(knn_val: torch.Tensor, knn_ind: torch.Tensor) -> torch.Tensor:
    """``compute_normalized_laplacian`` of Model/LATTICE.py:56-64 / the 'sym' branch of
    Model/MGCN.py:21-30 on the sparse form: d^-1/2[r] * v * d^-1/2[c] with d = row sums, inf -> 0.
    Bit-identical to the reference's diagflat GEMMs at O(N*k) cost."""
    n, k = knn_ind.shape
    rowsum = knn_val.sum(-1)
    d_inv_sqrt = torch.pow(rowsum, -0.5)
    d_inv_sqrt[torch.isinf(d_inv_sqrt)] = 0.
    rows = torch.arange(n, device=knn_ind.device).unsqueeze(1).expand(-1, k)
    vals = d_inv_sqrt[rows] * knn_val * d_inv_sqrt[knn_ind]
    return torch.sparse_coo_tensor(torch.stack((rows.flatten(), knn_ind.flatten())), vals.flatten(), (n, n))


def knn_self_excluded(mm_embeddings: torch.Tensor, item_topk: int) -> torch.Tensor:
    """Convention C, drop-in for ``get_knn_adj_mat(mm_embeddings, item_topk)`` of
    gen_hypergraph_u_i.py:41-48 and Model/MHRec.py:671-678 -> knn_ind int64 [N, k]."""
    return knn_topk(mm_embeddings, item_topk, exclude_self=True)[1]
