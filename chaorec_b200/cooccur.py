"""User-user co-occurrence graph.

Reference: ``dualgnn-gen-u-u-matrix.py`` — ``gen_user_matrix`` (lines 20-50) intersects the item sets of every
pair of users in a Python double loop into a dense [U, U] float matrix; lines 85-103 keep the
min(200, nnz) largest entries of every row as ``user_graph_dict[u] = [indices, values]``.  DualGNN, DRAGON and
MHRec load that file.  One kernel (csrc/cooccur.cu) produces the same rows from the two CSR views of the
training interactions.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from .ranking import mask_csr_from_edges


def user_cooccurrence_topk(train_users: torch.Tensor, train_items_local: torch.Tensor, num_user: int, num_item: int,
                           k: int = 200, row_range: Optional[Tuple[int, int]] = None):
    """(idx int32 [U, k] (-1 = unused), cnt int32 [U, k], n int32 [U]) for the training pairs
    ``(train_users[e], train_items_local[e])`` on the GPU.  Row u lists the users sharing the most items with u,
    best first, ties towards the smaller user id; ``row_range`` restricts the rows computed (sharding)."""
    _lib.require_cuda(train_users, "train_users")
    lib = _lib.load()
    dev = train_users.device
    by_user = mask_csr_from_edges(train_users, train_items_local, num_user, num_item)            # dedup + sort
    users = torch.arange(num_user, device=dev, dtype=torch.int64).repeat_interleave(by_user.rowptr[1:] - by_user.rowptr[:-1])
    by_item = mask_csr_from_edges(by_user.col.to(torch.int64), users, num_item, num_user)         # transpose
    if int((by_user.rowptr[1:] - by_user.rowptr[:-1]).max()) >= 65536:
        raise RuntimeError("user_cooccurrence_topk: a user with 65536+ items overflows the 16-bit counters")
    lo, hi = (0, num_user) if row_range is None else (int(row_range[0]), int(row_range[1]))
    idx = torch.full((num_user, k), -1, dtype=torch.int32, device=dev)
    cnt = torch.zeros((num_user, k), dtype=torch.int32, device=dev)
    n = torch.zeros(num_user, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.chaorec_cooccurrence_topk(_lib.ptr(by_user.rowptr), _lib.ptr(by_user.col), _lib.ptr(by_item.rowptr),
                                           _lib.ptr(by_item.col), num_user, num_item, int(k), lo, hi, _lib.ptr(idx),
                                           _lib.ptr(cnt), _lib.ptr(n), _lib.stream_ptr(dev))
    _lib.check(rc, "chaorec_cooccurrence_topk")
    return idx, cnt, n


def user_graph_dict(idx: torch.Tensor, cnt: torch.Tensor, n: torch.Tensor) -> Dict[int, List[list]]:
    """The ``user_graph_dict`` the reference saves (dualgnn-gen-u-u-matrix.py:85-103):
    ``{u: [neighbour ids, co-occurrence counts as floats]}``."""
    idx_h, cnt_h, n_h = idx.cpu().tolist(), cnt.cpu().tolist(), n.cpu().tolist()
    return {u: [idx_h[u][:n_h[u]], [float(c) for c in cnt_h[u][:n_h[u]]]] for u in range(len(n_h))}
