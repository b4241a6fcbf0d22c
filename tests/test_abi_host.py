"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
host metric reduction reproduces the reference's numbers, and the host-side data-format logic
(dict / object arrays -> CSR) is right.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import chaorec_b200 as cb
from chaorec_b200 import _lib
from conftest import ROOT, dict_from_train, load_golden, metrics_array, rows_to_object_array, unflatten_rows


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "chaorec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chaorec_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_strings():
    lib = _lib.load()
    assert lib.chaorec_abi_version() == 3
    assert lib.chaorec_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.chaorec_error_string(code)) > 4
    assert lib.chaorec_spmm_hub_chunk() >= lib.chaorec_spmm_hub_threshold() > 0
    import ctypes
    t, c = ctypes.c_int32(0), ctypes.c_int32(0)
    for nnz, want in ((0, 64), (562_350, 64), (4_000_000, 128), (20_000_000, 256), (200_000_000, 256)):
        assert lib.chaorec_spmm_hub_plan(nnz, ctypes.byref(t), ctypes.byref(c)) == 0
        assert (t.value, c.value) == (want, 4 * want)
    assert lib.chaorec_spmm_hub_row_chunks(30_000, 256) == 59       # ceil(30000/64)=469 -> 512-edge chunks
    assert lib.chaorec_spmm_hub_row_chunks(4097, 4096) == 2 and lib.chaorec_spmm_hub_row_chunks(100, 256) == 1


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA"):
        cb.rank_topk(torch.zeros(4, 64), torch.zeros(8, 64), None, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        cb.knn_topk(torch.zeros(8, 64), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        cb.csr_from_coo(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), None, 4, 4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "chaorec_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def _hits_numpy(ranklist, rows, K):
    out = np.zeros(len(rows), dtype=np.uint64)
    for r, row in enumerate(rows):
        pos = set(row[1:])
        bits = 0
        for j, item in enumerate(ranklist[row[0]][:K].tolist()):
            if item in pos:
                bits |= 1 << j
        out[r] = bits
    return out


def test_host_metrics_reproduce_reference_exactly():
    g = load_golden("lightgcn_small.npz")
    ks = [int(k) for k in g["ks"]]
    for rows_key, m_key in (("val_rows", "metrics_val"), ("test_rows", "metrics_test")):
        rows = rows_to_object_array(g[rows_key])
        hits = _hits_numpy(g["ranklist"], rows, 50)
        n_pos = np.array([len(r) - 1 for r in rows], dtype=np.int32)
        m = cb.metrics_from_hits(hits, n_pos, ks)
        assert metrics_array(m, ks).tolist() == g[m_key].tolist()
    a = load_golden("metrics_adversarial.npz")
    ks = [int(k) for k in a["ks"]]
    rows = unflatten_rows(a["rows_flat"])
    hits = _hits_numpy(a["ranklist"], rows, 50)
    m = cb.metrics_from_hits(hits, np.array([len(r) - 1 for r in rows], dtype=np.int32), ks)
    assert metrics_array(m, ks).tolist() == a["metrics"].tolist()
    assert isinstance(m[ks[0]]["ndcg"], float)


def test_host_metrics_empty_positive_rows():
    hits = np.zeros(3, dtype=np.uint64)
    m = cb.metrics_from_hits(hits, np.array([0, 2, 0], dtype=np.int32), [5])
    assert m[5] == {'precision': 0.0, 'recall': 0.0, 'ndcg': 0.0, 'hit_rate': 0.0, 'map': 0.0}


def test_mask_csr_from_user_item_dict():
    g = load_golden("lightgcn_small.npz")
    U, I = int(g["num_user"]), int(g["num_item"])
    uid = dict_from_train(g["train"])
    uid[3] = list(uid[3]) + [uid[3][0]]            # duplicate
    uid[5] = list(uid[5]) + [U + I + 7, U - 1]     # out of range on both sides
    m = cb.build_mask_csr(uid, U, I, torch.device("cpu"))
    rp, col = m.rowptr.numpy(), m.col.numpy()
    assert rp[0] == 0 and rp[-1] == col.size and col.dtype == np.int32
    for u in range(U):
        want = sorted(set(i - U for i in uid.get(u, []) if U <= i < U + I))
        assert col[rp[u]:rp[u + 1]].tolist() == want
    m2 = cb.mask_csr_from_edges(torch.from_numpy(g["train"][:, 0].astype(np.int64)),
                                torch.from_numpy(g["train"][:, 1].astype(np.int64) - U), U, I)
    m0 = cb.build_mask_csr(dict_from_train(g["train"]), U, I, torch.device("cpu"))
    assert torch.equal(m2.rowptr, m0.rowptr) and torch.equal(m2.col, m0.col)


def test_positives_csr_from_val_rows():
    a = load_golden("metrics_adversarial.npz")
    rows = unflatten_rows(a["rows_flat"])
    p = cb.build_positives_csr(rows, torch.device("cpu"))
    assert p.row_user.tolist() == [r[0] for r in rows]
    assert p.n_pos.tolist() == [len(r) - 1 for r in rows]
    rp = p.rowptr.numpy()
    for i, r in enumerate(rows):
        assert p.col[rp[i]:rp[i + 1]].tolist() == sorted(r[1:])


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under chaorec_b200/ may import, load or execute it."""
    import pathlib
    import re
    pkg = pathlib.Path(__file__).resolve().parents[1] / "chaorec_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        code = "\n".join(line for line in text.splitlines() if not line.lstrip().startswith(("#", "//", "*")))
        code = re.sub(r'"""[\s\S]*?"""', "", code)             # docstrings may cite oracle/ files
        assert not re.search(r"^\s*(from|import)\s+oracle\b", code, re.M), path
        assert "canon_py" not in code and "libcanon" not in code and "ref_ops" not in code, path


def test_cuda_entry_points_refuse_cpu_tensors():
    """No CPU fallback: the public calls raise on CPU tensors instead of computing somewhere else."""
    import pytest
    import torch
    import chaorec_b200 as cb
    x = torch.zeros(8, 64)
    ei = torch.zeros(2, 4, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        cb.rank_topk(x, x, None, 2)
    with pytest.raises(RuntimeError):
        cb.knn_topk(x, 2)
    with pytest.raises(RuntimeError):
        cb.lightgcn_forward(x, x, ei, 1)
    with pytest.raises(RuntimeError):
        cb.sample_negatives(torch.zeros(4, dtype=torch.int64), None, 8, 0)
    with pytest.raises(RuntimeError):
        cb.user_cooccurrence_topk(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64), 4, 4)


def test_conversion_caches_are_keyed_by_object_identity_not_by_a_recyclable_id():
    """A cached CSR conversion must never be served for a different container that happens to reuse the id()
    and length of a collected one (two consecutive val/test arrays did exactly that)."""
    import types
    from collections import OrderedDict
    import numpy as np
    import torch
    from chaorec_b200 import ranking
    cache = OrderedDict()
    dev = torch.device("cpu")
    fake = lambda tag: types.SimpleNamespace(rowptr=torch.zeros(1), tag=tag)      # noqa: E731
    first = np.empty(5, dtype=object)
    ranking._cache_put(cache, first, fake("first"))
    assert ranking._cache_get(cache, first, dev).tag == "first"
    ident = id(first)
    del first                                           # the entry keeps the array alive: its id cannot be recycled
    assert cache[ident][0] is not None
    for _ in range(64):                                 # same shape, same length: must always miss
        other = np.empty(5, dtype=object)
        assert id(other) != ident and ranking._cache_get(cache, other, dev) is None
    for i in range(20):                                 # bounded
        ranking._cache_put(cache, np.empty(3, dtype=object), fake(i))
    assert len(cache) <= ranking._CACHE_ENTRIES


def test_host_side_argument_checks_of_the_round_2_entry_points():
    """No GPU needed: argument validation of the fused-exchange entry points (C side) and of the python shims."""
    lib = _lib.load()
    import ctypes as C
    one = (C.c_void_p * 1)(0)
    # push: needs peers, a slot inside [0, n_peers), enough blocks for the rows
    assert lib.chaorec_spmm_csr_push_f32(None, None, None, None, None, None, None, None, 0, 0, 0, 0, None, None, None, 8, 8, 64,
                                         one, 0, 4, 0, None) != 0
    assert lib.chaorec_reduce_slots_broadcast_f32(None, 2, 4, 4, 64, 0, one, None, 0.25, 0, None, 0, 0, None) != 0
    assert b"bad argument" in lib.chaorec_error_string(-1) or lib.chaorec_error_string(-1) is not None
    # python shims
    with pytest.raises(RuntimeError, match="pair up"):
        cb.rank_topk_sum([torch.zeros(2, 4)], [], None, 1)
    from chaorec_b200 import dist as cdist
    with pytest.raises(RuntimeError, match="process group"):
        cdist.FusedItemExchange(10, 64, torch.device("cpu"))
    assert cdist.shard_range(10, 8, 7) == (9, 10)
