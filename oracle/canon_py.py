"""ctypes wrapper of oracle/libcanon.so (the C restatement of the canonical arithmetic).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libcanon.so"

_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32


def build(force: bool = False) -> Path:
    src = HERE / "canon.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "libcanon.so"], check=True, capture_output=True)
    return LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Canon:
    def __init__(self):
        build()
        lib = C.CDLL(str(LIB))
        lib.canon_score_mask_topk_f32.restype = C.c_int
        lib.canon_score_mask_topk_f32.argtypes = [_vp, _vp, _vp, _vp, C.c_float, _i64, _i64, _i32, _i32, _i64, _vp, _vp]
        lib.canon_knn_cosine_topk.restype = C.c_int
        lib.canon_knn_cosine_topk.argtypes = [_vp, _i64, _i32, _i32, C.c_int, _i64, _i64, _vp, _vp]
        lib.canon_l2_normalize_f32.restype = C.c_int
        lib.canon_l2_normalize_f32.argtypes = [_vp, _i64, _i32, _vp]
        lib.canon_pair_scores_f32.restype = C.c_int
        lib.canon_pair_scores_f32.argtypes = [_vp, _vp, _i32, _vp, _vp, _i64, _vp]
        lib.canon_spmm_csr_f32.restype = C.c_int
        lib.canon_spmm_csr_f32.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32]
        lib.canon_set_threads.argtypes = [C.c_int]
        lib.canon_get_threads.restype = C.c_int
        self.lib = lib

    def set_threads(self, n: int) -> None:
        self.lib.canon_set_threads(int(n))

    def threads(self) -> int:
        return int(self.lib.canon_get_threads())

    def score_mask_topk(self, user_emb, item_emb, mask_rowptr, mask_col, mask_value, K, id_offset=0):
        u = np.ascontiguousarray(user_emb, dtype=np.float32)
        it = np.ascontiguousarray(item_emb, dtype=np.float32)
        U, d = u.shape
        I = it.shape[0]
        rp = None if mask_rowptr is None else np.ascontiguousarray(mask_rowptr, dtype=np.int64)
        mc = None if mask_col is None else np.ascontiguousarray(mask_col, dtype=np.int32)
        if mc is not None and mc.size == 0:
            mc = np.zeros(1, dtype=np.int32)
        idx = np.empty((U, K), dtype=np.int64)
        sc = np.empty((U, K), dtype=np.float32)
        rc = self.lib.canon_score_mask_topk_f32(_p(u), _p(it), _p(rp), _p(mc), float(np.float32(mask_value)), U, I, d, K,
                                                int(id_offset), _p(idx), _p(sc))
        if rc != 0:
            raise RuntimeError(f"canon_score_mask_topk_f32 -> {rc}")
        return idx, sc

    def knn_cosine_topk(self, feats, k, exclude_self=0, row_lo=0, row_hi=None):
        f = np.ascontiguousarray(feats, dtype=np.float32)
        N, D = f.shape
        row_hi = N if row_hi is None else row_hi
        val = np.empty((row_hi - row_lo, k), dtype=np.float32)
        idx = np.empty((row_hi - row_lo, k), dtype=np.int64)
        rc = self.lib.canon_knn_cosine_topk(_p(f), N, D, k, int(exclude_self), row_lo, row_hi, _p(val), _p(idx))
        if rc != 0:
            raise RuntimeError(f"canon_knn_cosine_topk -> {rc}")
        return val, idx

    def l2_normalize(self, feats):
        f = np.ascontiguousarray(feats, dtype=np.float32)
        out = np.empty_like(f)
        self.lib.canon_l2_normalize_f32(_p(f), f.shape[0], f.shape[1], _p(out))
        return out

    def pair_scores(self, a, b, rows, cols):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        out = np.empty(rows.shape[0], dtype=np.float32)
        self.lib.canon_pair_scores_f32(_p(a), _p(b), a.shape[1], _p(rows), _p(cols), rows.shape[0], _p(out))
        return out

    def spmm_csr(self, rowptr, col, val, row_scale, col_scale, x):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        val = None if val is None else np.ascontiguousarray(val, dtype=np.float32)
        rs = None if row_scale is None else np.ascontiguousarray(row_scale, dtype=np.float32)
        cs = None if col_scale is None else np.ascontiguousarray(col_scale, dtype=np.float32)
        M = rowptr.shape[0] - 1
        y = np.empty((M, x.shape[1]), dtype=np.float32)
        self.lib.canon_spmm_csr_f32(_p(rowptr), _p(col), _p(val), _p(rs), _p(cs), _p(x), _p(y), M, x.shape[1])
        return y
