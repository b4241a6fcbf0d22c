"""World-size-2 checks of the multi-GPU host logic over gloo on CPU tensors: user sharding + hit-mask
gather gives bit-identical metrics, the cyclic row deal + per-layer all-gather reproduces the
single-process propagation.  (The kernels themselves are GPU-only; torch CPU ops stand in for the
local SpMM here so that only the partitioning / collective logic is under test.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import chaorec_b200 as cb
from chaorec_b200 import dist as cdist
from conftest import load_golden, metrics_array, rows_to_object_array
from oracle import ref_ops


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _hits_numpy(ranklist, rows, K):
    out = np.zeros(len(rows), dtype=np.int64)
    for r, row in enumerate(rows):
        pos = set(row[1:])
        bits = 0
        for j, item in enumerate(ranklist[row[0]][:K].tolist()):
            if item in pos:
                bits |= 1 << j
        out[r] = bits
    return out


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- evaluation: shard rows, gather hit masks, reduce on rank 0 -----------------------
        g = load_golden("lightgcn_small.npz")
        rows = rows_to_object_array(g["test_rows"])
        ks = [int(k) for k in g["ks"]]
        lo, hi = cdist.shard_range(len(rows), world, rank)
        local_hits = torch.from_numpy(_hits_numpy(g["ranklist"], rows[lo:hi], 50))
        all_hits = cdist.gather_to_rank0(local_hits)
        if rank == 0:
            n_pos = np.array([len(r) - 1 for r in rows], dtype=np.int32)
            m = cb.metrics_from_hits(all_hits.numpy(), n_pos, ks)
            results["metrics_ok"] = metrics_array(m, ks).tolist() == g["metrics_test"].tolist()
        else:
            assert all_hits is None
        cat = cdist.allgather_concat(torch.arange(rank * 3, rank * 3 + 2 + rank))
        assert cat.tolist() == [0, 1, 3, 4, 5]

        # --- propagation: cyclic deal, relabel, local product, all-gather per layer ------------
        n, d, L = 301, 8, 3
        gen = torch.Generator().manual_seed(3)
        r = torch.randint(0, n, (4000,), generator=gen)
        c = torch.randint(0, n, (4000,), generator=gen)
        v = torch.rand(4000, generator=gen)
        x0 = torch.randn(n, d, generator=gen)
        full = torch.sparse_coo_tensor(torch.stack((r, c)), v, (n, n))
        want, e = x0 / (L + 1), x0
        for _ in range(L):
            e = torch.sparse.mm(full, e)
            want = want + e / (L + 1)
        part = cdist.CyclicPartition(n, world)
        assert torch.equal(part.from_permuted(part.to_permuted(torch.arange(n))), torch.arange(n))
        lr, pc, lv = part.local_edges(r, c, v, rank)
        local = torch.sparse_coo_tensor(torch.stack((lr, pc)), lv, (part.block, part.n_padded))
        cur = part.permute_rows(x0)
        B = part.block
        acc = cur[rank * B:(rank + 1) * B] / (L + 1)
        for l in range(L):
            y = torch.sparse.mm(local, cur)
            acc = acc + y / (L + 1)
            if l < L - 1:
                nxt = torch.empty_like(cur)
                cdist.allgather_blocks(y, nxt)
                cur = nxt
        out_perm = torch.empty(part.n_padded, d)
        cdist.allgather_blocks(acc.contiguous(), out_perm)
        got = part.unpermute_rows(out_perm)
        if rank == 0:
            results["prop_err"] = float((got - want).abs().max() / want.abs().max())

        # --- bipartite propagation: users stay local, item partials are all-reduced ----------------
        U, I, d, L = 203, 57, 8, 3
        tu, ti = ref_ops.synth_interactions(U, I, 1500, 5)
        ei = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy())
        gen = torch.Generator().manual_seed(6)
        xu0, xi0 = torch.randn(U, d, generator=gen), torch.randn(I, d, generator=gen)
        want = ref_ops.lightgcn_forward(xu0, xi0, ei, L)
        u_lo, u_hi = cdist.shard_range(U, world, rank)
        mine = (tu >= u_lo) & (tu < u_hi)
        lu, li = tu[mine] - u_lo, ti[mine]
        item_deg = torch.bincount(ti, minlength=I).float()
        user_deg = torch.bincount(lu, minlength=u_hi - u_lo).float()
        w = user_deg.pow(-0.5)[lu] * item_deg.pow(-0.5)[li]
        shard = cdist.BipartiteShard(U, I, u_lo, u_hi,
                                     torch.sparse_coo_tensor(torch.stack((lu, li)), w, (u_hi - u_lo, I)),
                                     torch.sparse_coo_tensor(torch.stack((li, lu)), w, (I, u_hi - u_lo)))
        au, ai = cdist.propagate_mean_bipartite(shard, xu0[u_lo:u_hi].clone(), xi0.clone(), L, spmm=torch.sparse.mm)
        all_u = cdist.allgather_concat(au)
        if rank == 0:
            got = torch.cat((all_u, ai))
            results["bip_err"] = float((got - want).abs().max() / want.abs().max())

        # --- kNN build: rows sharded, [N/G, k] outputs all-gathered --------------------------------
        feats = torch.randn(97, 12, generator=torch.Generator().manual_seed(9))

        def cpu_topk(x, k, exclude_self=False, row_range=None):
            xn = x / x.norm(dim=1, keepdim=True)
            lo_, hi_ = row_range
            sim = xn[lo_:hi_] @ xn.t()
            v_, i_ = torch.sort(sim, dim=1, descending=True, stable=True)
            return v_[:, :k].contiguous(), i_[:, :k].contiguous()

        kv, ki = cdist.knn_topk_sharded(feats, 5, topk_fn=cpu_topk)
        wv, wi = cpu_topk(feats, 5, row_range=(0, 97))
        if rank == 0:
            results["knn_ok"] = bool(torch.equal(ki, wi) and torch.allclose(kv, wv, atol=1e-6))   # (per-row-block GEMMs may round differently)
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharding_and_collectives():
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert results["metrics_ok"] is True
    assert results["prop_err"] < 1e-5
    assert results["bip_err"] < 1e-5
    assert results["knn_ok"] is True


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            bounds = [cdist.shard_range(n, w, r) for r in range(w)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1
