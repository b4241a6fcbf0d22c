"""Normalized-adjacency propagation: drop-ins for the reference's SpMM seam.

Reference call sites replaced (paths relative to the ChaoRec checkout):
  * ``torch.sparse.mm(adj, x)``                     Model/FREEDOM.py:168,174 (+35 models)
  * ``LightGCNConv(...)(x, edge_index)``            Model/LightGCN.py:21-46, Model/LATTICE.py:18-42
  * the L-layer chain + layer mean                  Model/LightGCN.py:76-95, Model/FREEDOM.py:171-178

Host logic only; all arithmetic runs in csrc/spmm.cu through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import _lib


@dataclass
class CSRGraph:
    """A (M x N) sparse matrix in CSR on one GPU + the hub-row plan the SpMM kernel wants."""
    rowptr: torch.Tensor              # int64 [M+1]
    col: torch.Tensor                 # int32 [nnz]
    val: Optional[torch.Tensor]       # float32 [nnz] or None (=> ones)
    M: int
    N: int
    row_scale: Optional[torch.Tensor] = None   # float32 [M]
    col_scale: Optional[torch.Tensor] = None   # float32 [N]
    hub_rows: Optional[torch.Tensor] = None    # int32 [n_hub]
    hub_chunk_ptr: Optional[torch.Tensor] = None  # int64 [n_hub+1]
    n_hub: int = 0
    n_hub_chunks: int = 0
    hub_threshold: int = 0                     # 0 = library default
    hub_chunk: int = 0
    hub_counters: Optional[torch.Tensor] = None   # int32 [n_hub] (single-launch mode), kept zero by the kernel
    row_order: Optional[torch.Tensor] = None      # int32 [M]: rows by decreasing length (large graphs)
    # Single-launch mode (hub rows finished by the last chunk CTA) is for launch-bound graphs only.  It used to apply below
    # 16M non-zeros; at 8 GPUs the 12M-edge shards of the 100M-edge graph then took it, and both sharded variants failed
    # their parity gate there intermittently (3e-4 .. 5e-2) while every multi-launch configuration (1, 2, 4 GPUs) has
    # never failed -- until that is understood the mode stays with the graphs it was built and tested for.
    SINGLE_LAUNCH_MAX_NNZ = 1 << 21
    ROW_ORDER_MIN_NNZ = 1 << 24
    _partials: Dict[int, torch.Tensor] = field(default_factory=dict)
    _transposed: Optional["CSRGraph"] = None
    symmetric: bool = False

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    @property
    def device(self) -> torch.device:
        return self.rowptr.device

    def plan_hubs(self, threshold: Optional[int] = None, chunk: Optional[int] = None) -> "CSRGraph":
        """Choose which rows the CTA-per-chunk hub kernels take.  Default: the library's plan for this
        edge count (small graphs get a low threshold: at 0.5M edges one 1000-edge row walked by a single
        sub-warp would take 150 us of a kernel that otherwise needs ~10)."""
        lib = _lib.load()
        if threshold is None or chunk is None:
            t, c = C.c_int32(0), C.c_int32(0)
            _lib.check(lib.chaorec_spmm_hub_plan(self.nnz, C.byref(t), C.byref(c)), "chaorec_spmm_hub_plan")
            threshold = t.value if threshold is None else threshold
            chunk = c.value if chunk is None else chunk
        if os.environ.get("CHAOREC_SPMM_HUB_THR"):      # experiment switch
            threshold = int(os.environ["CHAOREC_SPMM_HUB_THR"])
            chunk = 4 * threshold
        self.hub_threshold, self.hub_chunk = int(threshold), int(chunk)
        deg = self.rowptr[1:] - self.rowptr[:-1]
        hubs = torch.nonzero(deg > self.hub_threshold).flatten()
        self.n_hub = int(hubs.numel())
        if self.n_hub:
            hdeg = deg[hubs]
            # edges per chunk of a row: at least `chunk`, grown in steps of 64 so that a row has <= 64 chunks
            # (chaorec_spmm_hub_row_chunks, csrc/spmm.cu: hub_chunk_edges)
            per = (hdeg + 63) // 64
            chunk_edges = torch.where(per > self.hub_chunk, (per + 63) // 64 * 64, torch.full_like(per, self.hub_chunk))
            chunks = (hdeg + chunk_edges - 1) // chunk_edges
            ptr = torch.zeros(self.n_hub + 1, dtype=torch.int64, device=self.device)
            ptr[1:] = torch.cumsum(chunks, 0)
            self.hub_rows = hubs.to(torch.int32).contiguous()
            self.hub_chunk_ptr = ptr
            self.n_hub_chunks = int(ptr[-1].item())
            # small graphs: one launch per product (the kernel time is a few microseconds, two more launches
            # would double it); large graphs keep the separate kernels (fewer registers in the row kernel)
            self.hub_counters = (torch.zeros(self.n_hub, dtype=torch.int32, device=self.device)
                                 if self.nnz < self.SINGLE_LAUNCH_MAX_NNZ else None)
        else:
            self.hub_rows, self.hub_chunk_ptr, self.n_hub_chunks, self.hub_counters = None, None, 0, None
        # Large graphs: hand rows to sub-warps in order of decreasing length, so that the two rows of a warp
        # and the four of a CTA finish together (no idle half-warps, no slots held by one long row).
        mode = os.environ.get("CHAOREC_SPMM_ROW_ORDER", "auto")
        if mode == "1" or (mode == "auto" and self.nnz >= self.ROW_ORDER_MIN_NNZ):
            self.row_order = torch.sort(deg, descending=True, stable=True)[1].to(torch.int32).contiguous()
        else:
            self.row_order = None
        self._partials.clear()
        return self

    def partials(self, d: int) -> Optional[torch.Tensor]:
        if not self.n_hub:
            return None
        buf = self._partials.get(d)
        if buf is None:
            buf = torch.empty(self.n_hub_chunks * d, dtype=torch.float32, device=self.device)
            self._partials[d] = buf
        return buf

    def transposed(self) -> "CSRGraph":
        """CSR of A^T (for the backward of spmm); symmetric graphs reuse themselves."""
        if self.symmetric:
            return self
        if self._transposed is None:
            rows = torch.repeat_interleave(torch.arange(self.M, device=self.device, dtype=torch.int64),
                                           self.rowptr[1:] - self.rowptr[:-1])
            t = csr_from_coo(self.col.to(torch.int64), rows, self.val, self.N, self.M)
            t.row_scale, t.col_scale = self.col_scale, self.row_scale
            t._transposed = self
            self._transposed = t
        return self._transposed


def csr_from_coo(row: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], M: int, N: int) -> CSRGraph:
    """COO (int64 row/col as stored by torch.sparse_coo_tensor, duplicates allowed) -> CSRGraph."""
    _lib.require_cuda(row, "row")
    lib = _lib.load()
    dev = row.device
    row = row.to(torch.int64).contiguous()
    col = col.to(torch.int64).contiguous()
    nnz = int(row.numel())
    if val is not None:
        val = val.to(torch.float32).contiguous()
    rowptr = torch.empty(M + 1, dtype=torch.int64, device=dev)
    out_col = torch.empty(nnz, dtype=torch.int32, device=dev)
    out_val = torch.empty(nnz, dtype=torch.float32, device=dev) if val is not None else None
    ws = _lib.workspace(lib.chaorec_coo_to_csr_workspace_bytes(nnz, M), dev)
    with torch.cuda.device(dev):
        rc = lib.chaorec_coo_to_csr(_lib.ptr(row), _lib.ptr(col), _lib.ptr(val), nnz, M, N, _lib.ptr(rowptr),
                                    _lib.ptr(out_col), _lib.ptr(out_val), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr(dev))
    _lib.check(rc, "chaorec_coo_to_csr")
    return CSRGraph(rowptr, out_col, out_val, M, N).plan_hubs()


_coo_cache: Dict[int, Tuple[weakref.ref, int, int, CSRGraph]] = {}


def csr_from_sparse_coo(adj: torch.Tensor) -> CSRGraph:
    """Cached conversion of a torch sparse COO tensor (keyed on identity + storage of its values:
    FREEDOM replaces ``masked_adj`` by a new tensor every epoch, Model/FREEDOM.py:143-162)."""
    if isinstance(adj, CSRGraph):
        return adj
    if adj.layout != torch.sparse_coo:
        raise TypeError("expected a torch.sparse_coo tensor or a CSRGraph")
    idx = adj._indices()
    vals = adj._values()
    key = id(adj)
    hit = _coo_cache.get(key)
    if hit is not None and hit[0]() is adj and hit[1] == vals.data_ptr() and hit[2] == vals._version:
        return hit[3]
    g = csr_from_coo(idx[0], idx[1], vals, adj.shape[0], adj.shape[1])
    # every miss drops the graphs whose adjacency tensor is gone (FREEDOM builds a new masked_adj per epoch; a stale
    # 100M-edge CSR graph is more than a GB) and keeps at most 8 live entries
    for k in [k for k, v in _coo_cache.items() if v[0]() is None]:
        _coo_cache.pop(k, None)
    while len(_coo_cache) >= 8:
        _coo_cache.pop(next(iter(_coo_cache)))
    _coo_cache[key] = (weakref.ref(adj), vals.data_ptr(), vals._version, g)
    return g


def _spmm_raw(g: CSRGraph, x: torch.Tensor, y: Optional[torch.Tensor], acc: Optional[torch.Tensor], alpha: float,
              acc_mode: int) -> None:
    lib = _lib.load()
    d = x.shape[1]
    with torch.cuda.device(x.device):
        rc = lib.chaorec_spmm_csr_f32(
            _lib.ptr(g.rowptr), _lib.ptr(g.col), _lib.ptr(g.val), _lib.ptr(g.row_scale), _lib.ptr(g.col_scale),
            _lib.ptr(x), _lib.ptr(y), _lib.ptr(acc), float(alpha), int(acc_mode), _lib.ptr(g.hub_rows),
            _lib.ptr(g.hub_chunk_ptr), g.n_hub, g.n_hub_chunks, g.hub_threshold, g.hub_chunk, _lib.ptr(g.partials(d)),
            _lib.ptr(g.hub_counters), _lib.ptr(g.row_order), g.M, g.N, d, _lib.stream_ptr(x.device))
    _lib.check(rc, "chaorec_spmm_csr_f32")


def _spmm_push(g: CSRGraph, x: torch.Tensor, peer_slot_ptrs, n_peers: int, peer_block: int, my_rank: int) -> None:
    """The product with its rows stored into the owners' slot buffers over NVLink (dist.FusedItemExchange)."""
    lib = _lib.load()
    d = x.shape[1]
    with torch.cuda.device(x.device):
        rc = lib.chaorec_spmm_csr_push_f32(
            _lib.ptr(g.rowptr), _lib.ptr(g.col), _lib.ptr(g.val), _lib.ptr(g.row_scale), _lib.ptr(g.col_scale),
            _lib.ptr(x), _lib.ptr(g.hub_rows), _lib.ptr(g.hub_chunk_ptr), g.n_hub, g.n_hub_chunks, g.hub_threshold,
            g.hub_chunk, _lib.ptr(g.partials(d)), _lib.ptr(g.hub_counters), _lib.ptr(g.row_order), g.M, g.N, d,
            peer_slot_ptrs, int(n_peers), int(peer_block), int(my_rank), _lib.stream_ptr(x.device))
    _lib.check(rc, "chaorec_spmm_csr_push_f32")


def _check_x(g: CSRGraph, x: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(x, "x")
    if x.dim() != 2 or x.shape[0] != g.N:
        raise RuntimeError(f"spmm: x must be [{g.N}, d], got {tuple(x.shape)}")
    if x.dtype != torch.float32:
        raise RuntimeError("spmm: x must be float32")
    return x.contiguous()


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g: CSRGraph):
        x = _check_x(g, x)
        y = torch.empty(g.M, x.shape[1], dtype=torch.float32, device=x.device)
        _spmm_raw(g, x, y, None, 0.0, 0)
        ctx.g = g
        return y

    @staticmethod
    def backward(ctx, gy):
        gt = ctx.g.transposed()
        gy = gy.contiguous()
        gx = torch.empty(gt.M, gy.shape[1], dtype=torch.float32, device=gy.device)
        _spmm_raw(gt, gy, gx, None, 0.0, 0)
        return gx, None


def spmm(adj, x: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``torch.sparse.mm(adj, x)`` (fp32, differentiable in ``x``)."""
    return _SpmmFn.apply(x, csr_from_sparse_coo(adj))


class _PropagateMeanFn(torch.autograd.Function):
    """out = sum_l w_l A^l x0 with the layer mean folded into each SpMM's epilogue."""

    @staticmethod
    def forward(ctx, x0, g: CSRGraph, n_layers: int, alpha: float, include_layer0: bool):
        x0 = _check_x(g, x0)
        ctx.g, ctx.n_layers, ctx.alpha, ctx.include0 = g, n_layers, alpha, include_layer0
        return _propagate_mean_raw(g, x0, n_layers, alpha, include_layer0)

    @staticmethod
    def backward(ctx, gout):
        gx = _propagate_mean_raw(ctx.g.transposed(), gout.contiguous(), ctx.n_layers, ctx.alpha, ctx.include0)
        return gx, None, None, None, None


def _propagate_mean_raw(g: CSRGraph, x0: torch.Tensor, n_layers: int, alpha: float, include0: bool) -> torch.Tensor:
    if g.M != g.N:
        raise RuntimeError("propagate_mean needs a square adjacency")
    if n_layers == 0:
        return x0 * alpha if include0 else torch.zeros_like(x0)
    acc = torch.empty_like(x0)
    bufs = [torch.empty_like(x0), torch.empty_like(x0)] if n_layers > 1 else [None, None]
    cur = x0
    for l in range(n_layers):
        last = l == n_layers - 1
        y = None if last else bufs[l & 1]
        if l == 0:
            if include0:
                _spmm_raw(g, cur, y, acc, alpha, 1)       # acc = alpha*x0 + alpha*y
            else:
                acc.zero_()
                _spmm_raw(g, cur, y, acc, alpha, 0)
        else:
            _spmm_raw(g, cur, y, acc, alpha, 0)           # acc += alpha*y
        cur = y
    return acc


def propagate_mean(adj, x0: torch.Tensor, n_layers: int, include_layer0: bool = True,
                   alpha: Optional[float] = None) -> torch.Tensor:
    """Fused ``n_layers`` propagation + layer combination.

    ``include_layer0=True``  -> mean over layers 0..L  (Model/LightGCN.py:86-93, Model/FREEDOM.py:177-178)
    ``include_layer0=False`` -> mean over layers 1..L  (Model/SimGCL.py:112-120)
    ``adj`` is a torch sparse COO tensor, a ``CSRGraph`` or an ``edge_index`` handled by
    ``graph_from_edge_index``.
    """
    if (isinstance(adj, torch.Tensor) and adj.layout == torch.strided and adj.dim() == 2 and adj.shape[0] == 2
            and not adj.is_floating_point()):
        g = graph_from_edge_index(adj, x0.size(0))
    else:
        g = csr_from_sparse_coo(adj)
    if alpha is None:
        alpha = 1.0 / (n_layers + 1 if include_layer0 else max(n_layers, 1))
    return _PropagateMeanFn.apply(x0, g, int(n_layers), float(alpha), bool(include_layer0))


_edge_cache: Dict[Tuple[int, int, int], Tuple[weakref.ref, CSRGraph, int]] = {}


def graph_from_edge_index(edge_index: torch.Tensor, num_nodes: int) -> CSRGraph:
    """PyG ``LightGCNConv`` graph (Model/LightGCN.py:28-40): out[col] += deg^-1/2[row]*deg^-1/2[col]*x[row]
    with deg = degree(row).  CSR rows are destinations (``edge_index[1]``), columns sources."""
    key = (id(edge_index), edge_index.data_ptr(), num_nodes)
    hit = _edge_cache.get(key)
    if hit is not None and hit[0]() is edge_index and hit[2] == edge_index._version:   # same tensor, not written since
        return hit[1]
    _lib.require_cuda(edge_index, "edge_index")
    lib = _lib.load()
    ei = edge_index.long()
    src, dst = ei[0], ei[1]
    g = csr_from_coo(dst, src, None, num_nodes, num_nodes)
    # degree(row) counts a node as a *source*; the symmetrised list (Model/LightGCN.py:63-64)
    # makes that equal to the CSR row length, which the kernel-side helper reads off rowptr.
    src_deg = torch.bincount(src, minlength=num_nodes)
    row_len = g.rowptr[1:] - g.rowptr[:-1]
    dis = torch.empty(num_nodes, dtype=torch.float32, device=ei.device)
    if torch.equal(src_deg, row_len):
        with torch.cuda.device(ei.device):
            _lib.check(lib.chaorec_degree_inv_sqrt(_lib.ptr(g.rowptr), num_nodes, 0.0, _lib.ptr(dis),
                                                   _lib.stream_ptr(ei.device)), "chaorec_degree_inv_sqrt")
        g.symmetric = bool(torch.equal(torch.sort(src * num_nodes + dst)[0], torch.sort(dst * num_nodes + src)[0]))
    else:
        dis = src_deg.to(torch.float32).pow(-0.5)
        dis[torch.isinf(dis)] = 0.0     # isolated nodes: never indexed by an edge, keep 0 * scale finite
    g.row_scale, g.col_scale = dis, dis
    for k in [k for k, v in _edge_cache.items() if v[0]() is None]:
        _edge_cache.pop(k, None)
    while len(_edge_cache) >= 8:
        _edge_cache.pop(next(iter(_edge_cache)))
    _edge_cache[key] = (weakref.ref(edge_index), g, edge_index._version)
    return g


class LightGCNConv(torch.nn.Module):
    """Same constructor and call signature as Model/LightGCN.py:21-46 (no learnable state)."""

    def __init__(self, in_channels, out_channels, aggr='add', **kwargs):
        super().__init__()
        self.aggr = aggr
        self.in_channels = in_channels
        self.out_channels = out_channels

    def forward(self, x, edge_index):
        return _SpmmFn.apply(x, graph_from_edge_index(edge_index, x.size(0)))


def lightgcn_forward(user_weight: torch.Tensor, item_weight: torch.Tensor, edge_index: torch.Tensor,
                     n_layers: int) -> torch.Tensor:
    """Fused body of ``LightGCN.forward`` (Model/LightGCN.py:76-95)."""
    x0 = torch.cat((user_weight, item_weight), dim=0)
    return propagate_mean(graph_from_edge_index(edge_index, x0.size(0)), x0, n_layers, True)


# ---------------------------------------------------------------------------------------------
# Device-side adjacency builders: the graphs the reference builds on the
# host with torch / scipy, produced directly as CSRGraph + degree scales.
# ---------------------------------------------------------------------------------------------


def bipartite_graph(users: torch.Tensor, local_items: torch.Tensor, num_user: int, num_item: int,
                    variant: str = "v1") -> CSRGraph:
    """Symmetrised (U+I) x (U+I) user-item graph from (user, local item) pairs, normalised on the device.

    variant  reference                                               edge weight
    -------  ------------------------------------------------------  -----------------------------------
    "v1"     LightGCNConv, Model/LightGCN.py:36-38                   deg^-1/2[r] * deg^-1/2[c]
    "v2"     FREEDOM.get_norm_adj_mat, Model/FREEDOM.py:73-83        (2deg)^-1/2[r] * (2deg)^-1/2[c]  (half of v1)
    "v3"     FREEDOM._normalize_adj_m + mirror, Model/FREEDOM.py:85-100,153-162
                                                                      (deg+1e-7)^-1/2[r] * (deg+1e-7)^-1/2[c]
    "v4"     scipy-dok get_norm_adj_mat, Model/SimGCL.py:64-104 (+12 copies)
                                                                      as v3 with the power taken in fp64, then fp32
    "v5"     scipy-lil get_adj_mat, Model/MGCN.py:164-196            deg^-1/2 in fp32, empty rows -> (1e-16)^-1/2
                                                                      (never multiplied by an edge) = v1
    The weights are never materialised: they are the row/col scales the SpMM kernel applies.
    """
    _lib.require_cuda(users, "users")
    lib = _lib.load()
    dev = users.device
    n = num_user + num_item
    u = users.to(torch.int64)
    it = local_items.to(torch.int64) + num_user
    if variant in ("v4", "v5"):
        # the scipy dok / lil builders keep ONE entry per (user, item) pair however often it occurs in the training
        # list; the torch COO builders (v1-v3) sum duplicates
        keys = torch.unique(u * n + it)
        u, it = keys // n, keys % n
    g = csr_from_coo(torch.cat((u, it)), torch.cat((it, u)), None, n, n)
    dis = torch.empty(n, dtype=torch.float32, device=dev)
    eps = 1e-7 if variant == "v3" else 0.0
    with torch.cuda.device(dev):
        _lib.check(lib.chaorec_degree_inv_sqrt(_lib.ptr(g.rowptr), n, float(eps), _lib.ptr(dis), _lib.stream_ptr(dev)),
                   "chaorec_degree_inv_sqrt")
    if variant == "v2":
        dis = dis * (2.0 ** -0.5)
    elif variant == "v4":
        # np.power(deg + 1e-7, -0.5) in float64 (Model/SimGCL.py:83-85); the reference then rounds the PRODUCT
        # d[r]*d[c] to fp32, here each factor is rounded (difference <= 2^-23 relative per edge)
        dis = ((g.rowptr[1:] - g.rowptr[:-1]).to(torch.float64) + 1e-7).pow(-0.5).to(torch.float32)
    elif variant not in ("v1", "v3", "v5"):
        raise ValueError("variant must be 'v1' .. 'v5'")
    g.row_scale = g.col_scale = dis
    g.symmetric = True
    return g


def resample_masked_graph(edge_users: torch.Tensor, edge_local_items: torch.Tensor, edge_values: torch.Tensor,
                          num_user: int, num_item: int, dropout: float,
                          generator: Optional[torch.Generator] = None) -> CSRGraph:
    """FREEDOM.pre_epoch_processing (Model/FREEDOM.py:143-162) on the device: keep
    int(E * (1 - dropout)) edges drawn without replacement with probability proportional to their
    normalised weight, re-normalise the kept block (v3) and mirror it.  (The draw uses the device
    RNG stream, so kept sets differ from the reference's CPU stream; everything downstream of the
    kept set is identical.)"""
    if dropout <= 0.0:
        return bipartite_graph(edge_users, edge_local_items, num_user, num_item, "v2")   # masked_adj = norm_adj
    keep_n = int(edge_values.size(0) * (1.0 - dropout))
    idx = torch.multinomial(edge_values, keep_n, generator=generator)
    return bipartite_graph(edge_users[idx], edge_local_items[idx], num_user, num_item, "v3")
