"""Full-ranking evaluation: drop-ins for ``gene_ranklist`` and ``evaluate``.

Reference call sites replaced (paths relative to the ChaoRec checkout):
  * ``model.gene_ranklist(topk=50)``            Model/LightGCN.py:137-162 (identical in 16 models;
                                                 batched form Model/MMGCN.py:204-244)
  * ``evaluate(model, data, ranklist, topk)``   train_and_evaluate.py:516-520 -> utils.gene_metrics
                                                 (utils.py:112-139) -> metrics.py:13-57

Host logic only: data-format conversion (python dict / object arrays -> CSR) and the calls into
csrc/rank_topk.cu and csrc/hits_metrics.cu through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

METRIC_NAMES = ('precision', 'recall', 'ndcg', 'hit_rate', 'map')


@dataclass
class MaskCSR:
    """Per-user train items as LOCAL item ids, sorted ascending and unique inside each row."""
    rowptr: torch.Tensor  # int64 [U+1]
    col: torch.Tensor     # int32 [nnz]


def build_mask_csr(user_item_dict: Dict[int, Iterable[int]], num_user: int, num_item: int,
                   device: torch.device) -> MaskCSR:
    """``user_item_dict`` {user: [global item ids]} (train items only, dataload.py:30) -> MaskCSR.
    Items outside [num_user, num_user+num_item) are dropped (they cannot index the score matrix)."""
    # no per-user numpy calls: one pass over the dict for the lengths, one flattening pass for the items
    n = len(user_item_dict)
    keys_u = np.fromiter(user_item_dict.keys(), dtype=np.int64, count=n)
    lens = np.fromiter((len(v) for v in user_item_dict.values()), dtype=np.int64, count=n)
    flat = np.fromiter((x for v in user_item_dict.values() for x in v), dtype=np.int64, count=int(lens.sum()))
    uu = np.repeat(keys_u, lens)
    ii = flat - num_user
    ok = (uu >= 0) & (uu < num_user) & (ii >= 0) & (ii < num_item)
    keys = np.unique(uu[ok] * num_item + ii[ok])
    uu, ii = keys // num_item, keys % num_item
    rowptr = np.zeros(num_user + 1, dtype=np.int64)
    np.cumsum(np.bincount(uu, minlength=num_user), out=rowptr[1:])
    return MaskCSR(torch.from_numpy(rowptr).to(device), torch.from_numpy(ii.astype(np.int32)).to(device))


def mask_csr_from_edges(users: torch.Tensor, local_items: torch.Tensor, num_user: int, num_item: int) -> MaskCSR:
    """Same as ``build_mask_csr`` from device tensors of (user, local item) pairs."""
    keys = torch.unique(users.to(torch.int64) * num_item + local_items.to(torch.int64))
    uu = keys // num_item
    rowptr = torch.zeros(num_user + 1, dtype=torch.int64, device=keys.device)
    rowptr[1:] = torch.cumsum(torch.bincount(uu, minlength=num_user), 0)
    return MaskCSR(rowptr, (keys % num_item).to(torch.int32).contiguous())


def rank_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, mask: Optional[MaskCSR], topk: int,
              mask_value: float = 1e-6, id_offset: int = 0, return_scores: bool = False, stats: Optional[dict] = None):
    """Fused score -> mask -> top-K.  Returns int64 [U, topk] (and float32 scores).

    ``stats`` (optional dict) receives ``engine`` ('tensor' | 'exact') and, for the tensor engine,
    ``fallback_rows`` = rows re-run by the exact engine (reading it synchronises the stream)."""
    _lib.require_cuda(user_emb, "user_emb")
    _lib.require_cuda(item_emb, "item_emb")
    lib = _lib.load()
    dev = user_emb.device
    user_emb = user_emb.detach().to(torch.float32).contiguous()
    item_emb = item_emb.detach().to(torch.float32).contiguous()
    U, d = user_emb.shape
    I = item_emb.shape[0]
    if item_emb.shape[1] != d:
        raise RuntimeError("user and item embeddings must have the same width")
    out_idx = torch.empty(U, topk, dtype=torch.int64, device=dev)
    out_score = torch.empty(U, topk, dtype=torch.float32, device=dev) if return_scores else None
    ws = _lib.workspace(lib.chaorec_score_mask_topk_workspace_bytes(U, I, d, topk), dev)
    with torch.cuda.device(dev):
        rc = lib.chaorec_score_mask_topk_f32(
            _lib.ptr(user_emb), _lib.ptr(item_emb), _lib.ptr(mask.rowptr) if mask else None,
            _lib.ptr(mask.col) if mask else None, float(np.float32(mask_value)), U, I, d, int(topk),
            int(id_offset), _lib.ptr(out_idx), _lib.ptr(out_score), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "chaorec_score_mask_topk_f32")
    if stats is not None:
        tensor = bool(lib.chaorec_uses_tensor_engine(U, I, d, int(topk)))
        stats["engine"] = "tensor" if tensor else "exact"
        stats["fallback_rows"] = int(ws[64:68].view(torch.int32).item()) if tensor else 0
        # which rows (d <= 64 engine only): parity gates re-derive every one of them on the host
        fb_ptr = lib.chaorec_score_mask_topk_fallback_rows(_lib.ptr(ws), U, I, d, int(topk)) if tensor else None
        if fb_ptr:
            off = (int(fb_ptr) - ws.data_ptr()) // 4
            stats["fallback_row_ids"] = ws.view(torch.int32)[off:off + stats["fallback_rows"]].to(torch.int64).sort()[0].cpu()
        # d <= 64 engine: first time-out record of the MMA pipeline's bounded waits (all zero = healthy; a non-zero
        # flag means every row was recomputed by the exact engine -- correct but slow, see csrc/tc_eval.cu)
        stats["pipeline_timeout"] = [int(v) for v in ws[144:176].view(torch.int32).tolist()] if tensor and d <= 64 else [0] * 8
        if stats["pipeline_timeout"][0]:
            stats["pipeline_trace"] = [hex(int(v) & 0xFFFFFFFF) for v in ws[176:256].view(torch.int32).tolist()]
    return (out_idx, out_score) if return_scores else out_idx


def rank_topk_sum(user_parts: Sequence[torch.Tensor], item_parts: Sequence[torch.Tensor], mask: Optional[MaskCSR], topk: int,
                  mask_value: float = 1e-6, id_offset: int = 0, **kw):
    """Full ranking under a SUM of inner products, ``sum_p user_parts[p] @ item_parts[p].T`` -- the score of the
    two-tower models (Model/SelfCF.py:213-238: ``u_online . i_target + u_target . i_online``) and of the three-term
    form (Model/MCLN.py:336-337).  A sum of inner products is the inner product of the concatenated embeddings, so
    this is one ``rank_topk`` call on ``cat(user_parts, 1)`` / ``cat(item_parts, 1)`` (canonical arithmetic: one
    fmaf chain over the concatenated width, parts in the given order)."""
    if len(user_parts) != len(item_parts) or not user_parts:
        raise RuntimeError("user_parts and item_parts must be non-empty and pair up")
    u = torch.cat([p.detach().to(torch.float32) for p in user_parts], dim=1)
    it = torch.cat([p.detach().to(torch.float32) for p in item_parts], dim=1)
    return rank_topk(u, it, mask, topk, mask_value, id_offset, **kw)


# Conversions of the reference's python containers are cached per OBJECT.  Every entry holds a strong reference
# to the object it was built from: an id() alone can be handed to a new object of the same length once the old
# one is collected (which made two consecutive val/test arrays share one conversion).  A few entries only.
_mask_cache: "OrderedDict[int, Tuple[object, int, MaskCSR]]" = OrderedDict()
_CACHE_ENTRIES = 8


def _cache_get(cache, obj, device):
    hit = cache.get(id(obj))
    if hit is not None and hit[0] is obj and hit[1] == len(obj) and hit[2].rowptr.device == device:
        cache.move_to_end(id(obj))
        return hit[2]
    return None


def _cache_put(cache, obj, value):
    cache[id(obj)] = (obj, len(obj), value)
    cache.move_to_end(id(obj))
    while len(cache) > _CACHE_ENTRIES:
        cache.popitem(last=False)


def clear_caches() -> None:
    """Drop the cached CSR conversions (call it after mutating ``user_item_dict`` / ``val_data`` in place)."""
    _mask_cache.clear()
    _pos_cache.clear()


def _model_mask(model, device) -> MaskCSR:
    d = model.user_item_dict
    m = _cache_get(_mask_cache, d, device)
    if m is None:
        m = build_mask_csr(d, model.num_user, model.num_item, device)
        _cache_put(_mask_cache, d, m)
    return m


def gene_ranklist(model, topk: int = 50, mask_value: float = 1e-6, device: Optional[torch.device] = None,
                  to_cpu: bool = True) -> torch.Tensor:
    """Drop-in for ``model.gene_ranklist(topk)`` (Model/LightGCN.py:137-162).

    Uses ``model.result`` (the cached embeddings of the last forward), ``model.num_user``,
    ``model.num_item`` and ``model.user_item_dict`` exactly like the reference method, and returns
    a LongTensor [num_user, topk] of GLOBAL item ids (on the CPU like the reference unless
    ``to_cpu=False``).  Bind it as a method with ``types.MethodType(gene_ranklist, model)``.
    """
    result = model.result
    if device is None:
        device = result.device if result.is_cuda else torch.device('cuda', torch.cuda.current_device())
    result = result.detach().to(device)
    U, I = model.num_user, model.num_item
    out = rank_topk(result[:U], result[U:U + I], _model_mask(model, device), topk, mask_value, id_offset=U)
    return out.cpu() if to_cpu else out


@dataclass
class PositivesCSR:
    """``val_data`` / ``test_data`` rows ([user, pos...] lists) in device CSR form."""
    row_user: torch.Tensor      # int64 [n_rows]  (device)
    rowptr: torch.Tensor        # int64 [n_rows+1] (device)
    col: torch.Tensor           # int64 [nnz] global ids, sorted inside each row (device)
    n_pos: np.ndarray           # int32 [n_rows] len(data[1:]) incl. duplicates (host)


_pos_cache: "OrderedDict[int, Tuple[object, int, PositivesCSR]]" = OrderedDict()


def build_positives_csr(data: Sequence[Sequence[int]], device: torch.device) -> PositivesCSR:
    cached = _cache_get(_pos_cache, data, device)
    if cached is not None:
        return cached
    n = len(data)
    lens = np.fromiter((len(r) - 1 for r in data), dtype=np.int64, count=n)
    flat = np.fromiter((v for r in data for v in r), dtype=np.int64, count=int(lens.sum()) + n)
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens + 1, out=starts[1:])
    row_user = flat[starts[:-1]]
    keep = np.ones(flat.shape[0], dtype=bool)
    keep[starts[:-1]] = False
    pos = flat[keep]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    row_of = np.repeat(np.arange(n, dtype=np.int64), lens)
    order = np.lexsort((pos, row_of))
    out = PositivesCSR(torch.from_numpy(row_user).to(device), torch.from_numpy(rowptr).to(device),
                       torch.from_numpy(pos[order]).to(device), lens.astype(np.int32))
    _cache_put(_pos_cache, data, out)
    return out


def hit_masks(ranklist: torch.Tensor, pos: PositivesCSR) -> torch.Tensor:
    """uint64-as-int64 [n_rows]: bit r <=> ranked item r of the row's user is a positive."""
    lib = _lib.load()
    dev = pos.rowptr.device
    ranklist = ranklist.to(device=dev, dtype=torch.int64).contiguous()
    n_rows = int(pos.row_user.numel())
    K = ranklist.shape[1]
    out = torch.empty(n_rows, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.chaorec_topk_hits(_lib.ptr(ranklist), _lib.ptr(pos.row_user), _lib.ptr(pos.rowptr), _lib.ptr(pos.col),
                                   n_rows, K, _lib.ptr(out), _lib.stream_ptr(dev))
    _lib.check(rc, "chaorec_topk_hits")
    return out


def metrics_from_hits(hits_host: np.ndarray, n_pos: np.ndarray, k_list: Sequence[int]) -> Dict[int, Dict[str, float]]:
    """Host reduction (sequential fp64, utils.py:116-137) of hit masks into the metrics dict."""
    lib = _lib.load()
    ks = np.asarray([int(k) for k in k_list], dtype=np.int32)
    max_k = int(ks.max()) if ks.size else 0
    inv_log = np.asarray([1.0 / np.log(i + 2) for i in range(max_k)], dtype=np.float64)  # metrics.py:31,34
    hits_host = np.ascontiguousarray(hits_host).view(np.uint64)
    n_pos = np.ascontiguousarray(n_pos, dtype=np.int32)
    out = np.zeros(len(ks) * 5, dtype=np.float64)
    rc = lib.chaorec_metrics_from_hits(hits_host.ctypes.data_as(C.c_void_p), n_pos.ctypes.data_as(C.c_void_p),
                                       int(hits_host.shape[0]), ks.ctypes.data_as(C.c_void_p), int(ks.size),
                                       inv_log.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    _lib.check(rc, "chaorec_metrics_from_hits")
    return {k: {name: float(out[t * 5 + j]) for j, name in enumerate(METRIC_NAMES)} for t, k in enumerate(k_list)}


def evaluate(model, data, ranklist: torch.Tensor, topk: Sequence[int], device: Optional[torch.device] = None):
    """Drop-in for ``evaluate(model, data, ranklist, topk)`` (train_and_evaluate.py:516-520).

    ``data``: val/test rows ``[user, pos...]``; ``ranklist``: LongTensor [num_user, >=max(topk)] of
    global ids (CPU or CUDA).  Returns ``{k: {'precision','recall','ndcg','hit_rate','map'}}``.
    """
    if model is not None and hasattr(model, 'eval'):
        model.eval()
    if device is None:
        device = ranklist.device if ranklist.is_cuda else torch.device('cuda', torch.cuda.current_device())
    if ranklist.shape[1] > 64:
        ranklist = ranklist[:, :64]
    pos = build_positives_csr(data, device)
    hits = hit_masks(ranklist, pos)
    return metrics_from_hits(hits.cpu().numpy(), pos.n_pos, list(topk))


def evaluate_full(user_emb: torch.Tensor, item_emb: torch.Tensor, mask: Optional[MaskCSR], splits: Sequence,
                  k_list: Sequence[int], topk: int = 50, mask_value: float = 1e-6, id_offset: Optional[int] = None):
    """Fused evaluation: one ranking pass, then metrics for every split (val, test, ...)."""
    if id_offset is None:
        id_offset = user_emb.shape[0]
    ranks = rank_topk(user_emb, item_emb, mask, topk, mask_value, id_offset)
    return ranks, [evaluate(None, s, ranks, k_list, device=ranks.device) for s in splits]


@dataclass
class HostEvalBatch:
    """One evaluation job in HOST (pinned) memory: what a reference-side caller holds after
    ``model.result.cpu()`` (Model/LightGCN.py:140-141) plus the CSR forms of its python inputs."""
    user_emb: torch.Tensor      # float32 [U_local, d]
    item_emb: torch.Tensor      # float32 [I, d]
    mask_rowptr: torch.Tensor   # int64 [U_local+1]
    mask_col: torch.Tensor      # int32 [nnz]
    pos_rowptr: torch.Tensor    # int64 [U_local+1]  (row r <-> local user r)
    pos_col: torch.Tensor       # int64 [nnz] global ids, sorted inside rows
    n_pos: np.ndarray           # int32 [U_local]
    id_offset: int              # num_user

    def pin(self) -> "HostEvalBatch":
        for name in ("user_emb", "item_emb", "mask_rowptr", "mask_col", "pos_rowptr", "pos_col"):
            setattr(self, name, getattr(self, name).contiguous().pin_memory())
        return self

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.user_emb, self.item_emb, self.mask_rowptr,
                                                          self.mask_col, self.pos_rowptr, self.pos_col))


_host_path_cache: Dict[Tuple[int, int], Tuple[torch.cuda.Stream, torch.Tensor]] = {}


def evaluate_host_batch(batch, topk: int, device: torch.device, mask_value: float = 1e-6, copy: bool = True):
    """Host buffers in, hit masks out: H2D copies, fused ranking, hit masks, D2H -- the part of an
    end-to-end evaluation that runs per rank.  ``batch`` is one ``HostEvalBatch`` or a list of them
    (consecutive user chunks sharing one item table): with a list, the copies of chunk i+1 run on a
    second stream while chunk i is ranked, and the hit masks stream back as they are produced.
    Returns the hit masks as an int64 host tensor [U_local] (synchronised).  The pinned staging buffer is cached per
    (device, user count): with ``copy=False`` the returned tensor IS that buffer and is only valid until the next
    call with the same user count (val followed by test!); the default hands out a copy."""
    chunks = batch if isinstance(batch, (list, tuple)) else [batch]
    lib = _lib.load()
    cur = torch.cuda.current_stream(device)
    total = sum(c.user_emb.shape[0] for c in chunks)
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev_index, total)
    cached = _host_path_cache.get(key)
    if cached is None:     # pinned allocations and stream creation are slow: once per shape
        cached = (torch.cuda.Stream(device=device), torch.empty(total, dtype=torch.int64, pin_memory=True))
        _host_path_cache.clear()
        _host_path_cache[key] = cached
    copy_stream, hits_host = cached
    copy_stream.wait_stream(cur)
    keep = []
    staged = []
    with torch.cuda.stream(copy_stream):
        ie = chunks[0].item_emb.to(device, non_blocking=True)
        for c in chunks:
            dev_t = (c.user_emb.to(device, non_blocking=True), c.mask_rowptr.to(device, non_blocking=True),
                     c.mask_col.to(device, non_blocking=True), c.pos_rowptr.to(device, non_blocking=True),
                     c.pos_col.to(device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            staged.append((dev_t, ev))
    off = 0
    for c, (dev_t, ev) in zip(chunks, staged):
        cur.wait_event(ev)
        ue, m_rp, m_col, p_rp, p_col = dev_t
        ranks = rank_topk(ue, ie, MaskCSR(m_rp, m_col), topk, mask_value, c.id_offset)
        n = ue.shape[0]
        hits = torch.empty(n, dtype=torch.int64, device=device)
        with torch.cuda.device(device):
            rc = lib.chaorec_topk_hits(_lib.ptr(ranks), None, _lib.ptr(p_rp), _lib.ptr(p_col), n, int(topk),
                                       _lib.ptr(hits), _lib.stream_ptr(device))
        _lib.check(rc, "chaorec_topk_hits")
        hits_host[off:off + n].copy_(hits, non_blocking=True)
        keep.append((ranks, hits))
        off += n
    cur.synchronize()
    copy_stream.synchronize()
    return hits_host.clone() if copy else hits_host
