"""GPU parity of the propagation path (csrc/spmm.cu, csrc/csr_build.cu) through the C ABI.

Tolerance (north_star): embeddings within 1e-5 relative error in fp32, measured as
max|delta| / max|reference|.
"""
import numpy as np
import pytest
import torch

import chaorec_b200 as cb
from conftest import load_golden
from oracle import ref_ops

pytestmark = pytest.mark.gpu
REL = 1e-5


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    ref = ref.to(torch.float32)
    return float((got.detach().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def dev():
    return torch.device("cuda", 0)


def test_lightgcn_conv_and_forward_match_reference_fixture():
    g = load_golden("lightgcn_small.npz")
    x0 = torch.from_numpy(np.concatenate([g["user_w"], g["item_w"]])).to(dev())
    ei = torch.from_numpy(g["edge_index"]).to(dev())
    conv = cb.LightGCNConv(64, 64, aggr='add')
    assert rel_err(conv(x0, ei), torch.from_numpy(g["one_layer"])) <= REL
    # the reference forward loop with the drop-in conv (Model/LightGCN.py:76-95)
    embs, x = [x0], x0
    for _ in range(int(g["n_layers"])):
        x = conv(x, ei)
        embs.append(x)
    final = torch.zeros_like(x0)
    for e in embs:
        final += (1.0 / len(embs)) * e
    assert rel_err(final, torch.from_numpy(g["result"])) <= REL
    fused = cb.lightgcn_forward(x0[:int(g["num_user"])], x0[int(g["num_user"]):], ei, int(g["n_layers"]))
    assert rel_err(fused, torch.from_numpy(g["result"])) <= REL


def test_freedom_sparse_mm_chain_matches_reference_fixture():
    g = load_golden("freedom_small.npz")
    U, I = int(g["num_user"]), int(g["num_item"])
    n = U + I
    uw, iw = torch.from_numpy(g["user_w"]).to(dev()), torch.from_numpy(g["item_w"]).to(dev())
    mm = torch.sparse_coo_tensor(torch.from_numpy(g["mm_adj_idx"]), torch.from_numpy(g["mm_adj_val"]), (I, I)).to(dev())
    for prefix, key in (("masked_adj", "result_masked"), ("norm_adj", "result_norm")):
        adj = torch.sparse_coo_tensor(torch.from_numpy(g[prefix + "_idx"]), torch.from_numpy(g[prefix + "_val"]),
                                      (n, n)).to(dev())      # uncoalesced, like the reference's
        # Model/FREEDOM.py:164-183 with torch.sparse.mm -> cb.spmm
        h = iw
        for _ in range(int(g["mm_layers"])):
            h = cb.spmm(mm, h)
        ego = torch.cat((uw, iw), 0)
        layers = [ego]
        for _ in range(int(g["n_layers"])):
            ego = cb.spmm(adj, ego)
            layers.append(ego)
        allm = torch.stack(layers, 1).mean(1)
        res = torch.cat((allm[:U], allm[U:] + h), 0)
        assert rel_err(res, torch.from_numpy(g[key])) <= REL
        fused = cb.propagate_mean(adj, torch.cat((uw, iw), 0), int(g["n_layers"]))
        res2 = torch.cat((fused[:U], fused[U:] + h), 0)
        assert rel_err(res2, torch.from_numpy(g[key])) <= REL


def test_real_graph_slice_matches_reference_fixture():
    g = load_golden("baby_slice.npz")
    U, I = int(g["num_user"]), int(g["num_item"])
    uw = ref_ops.xavier_embeddings(U, 64, int(g["seed_user"])).to(dev())
    iw = ref_ops.xavier_embeddings(I, 64, int(g["seed_item"])).to(dev())
    ei = ref_ops.symmetrise_edges(g["train"]).to(dev())
    out = cb.lightgcn_forward(uw, iw, ei, int(g["n_layers"]))
    rows = torch.from_numpy(g["check_rows"])
    err = (out.cpu()[rows] - torch.from_numpy(g["result_rows"])).abs().max() / float(g["result_absmax"])
    assert float(err) <= REL
    # isolated nodes: zero rows, no inf/nan leaking out of deg^-1/2
    assert torch.isfinite(out).all()


def _random_coo(M, N, nnz, seed, dup=True):
    g = torch.Generator().manual_seed(seed)
    r = torch.randint(0, M, (nnz,), generator=g)
    c = torch.randint(0, N, (nnz,), generator=g)
    v = torch.randn(nnz, generator=g)
    if dup:
        r = torch.cat((r, r[: nnz // 10]))
        c = torch.cat((c, c[: nnz // 10]))
        v = torch.cat((v, torch.randn(nnz // 10, generator=g)))
    return r, c, v


@pytest.mark.parametrize("M,N,d", [(700, 700, 64), (300, 900, 64), (900, 300, 32), (257, 513, 128), (200, 300, 100),
                                   (150, 150, 384), (64, 64, 512), (1000, 1000, 4)])
def test_spmm_matches_torch_sparse_mm(M, N, d):
    r, c, v = _random_coo(M, N, 6000, 100 + d)
    x = torch.randn(N, d, generator=torch.Generator().manual_seed(7))
    adj = torch.sparse_coo_tensor(torch.stack((r, c)), v, (M, N))
    ref = torch.sparse.mm(adj, x)
    got = cb.spmm(adj.to(dev()), x.to(dev()))
    assert got.shape == (M, d)
    assert rel_err(got, ref) <= REL


def test_spmm_hub_rows_empty_rows_and_scales():
    M = N = 3000
    g = torch.Generator().manual_seed(3)
    rows = [torch.randint(0, M, (20000,), generator=g), torch.full((30000,), 5), torch.full((1500,), 2999),
            torch.full((4097,), 17)]
    r = torch.cat(rows)
    c = torch.randint(0, N, (r.numel(),), generator=g)
    r[r == 100] = 101                                   # row 100 is empty
    v = torch.rand(r.numel(), generator=g)
    rs, cs = torch.rand(M, generator=g) + 0.5, torch.rand(N, generator=g) + 0.5
    x = torch.randn(N, 64, generator=g)
    graph = cb.csr_from_coo(r.to(dev()), c.to(dev()), v.to(dev()), M, N)
    graph.row_scale, graph.col_scale = rs.to(dev()), cs.to(dev())
    adj = torch.sparse_coo_tensor(torch.stack((r, c)), v.double() * rs.double()[r] * cs.double()[c], (M, N))
    ref = torch.sparse.mm(adj, x.double())
    # the library's plan for a 55k-edge graph: threshold 64, 256-edge chunks, at most 64 chunks per row
    assert (graph.hub_threshold, graph.hub_chunk) == (64, 256)
    assert graph.n_hub == 3 and graph.n_hub_chunks == 59 + 6 + 17      # 30k-edge row: 512-edge chunks
    got = cb.spmm(graph, x.to(dev()))
    assert rel_err(got, ref) <= REL
    assert float(got[100].abs().max()) == 0.0
    # single-launch mode (small graphs): the counters reset themselves, and the result is bit-identical to
    # the three-launch mode because the partial rows are summed in chunk order either way
    assert graph.hub_counters is not None
    again = cb.spmm(graph, x.to(dev()))
    assert torch.equal(again, got) and int(graph.hub_counters.abs().sum()) == 0
    graph.hub_counters = None
    assert torch.equal(cb.spmm(graph, x.to(dev())), got)
    # rows handed out by decreasing length (what large graphs do): same per-row arithmetic, same bits
    deg = graph.rowptr[1:] - graph.rowptr[:-1]
    graph.row_order = torch.sort(deg, descending=True, stable=True)[1].to(torch.int32).contiguous()
    assert torch.equal(cb.spmm(graph, x.to(dev())), got)
    graph.row_order = None
    # the large-graph plan on the same matrix
    graph.plan_hubs(threshold=1024, chunk=4096)
    assert graph.n_hub == 3 and graph.n_hub_chunks == 8 + 1 + 2
    assert rel_err(cb.spmm(graph, x.to(dev())), ref) <= REL
    # a tiny chunk size exercises the 64-chunks-per-row cap on every hub row
    graph.plan_hubs(threshold=16, chunk=16)
    assert rel_err(cb.spmm(graph, x.to(dev())), ref) <= REL
    # without a hub plan the same rows go through the sub-warp path
    graph.hub_rows, graph.n_hub, graph.n_hub_chunks = None, 0, 0
    assert rel_err(cb.spmm(graph, x.to(dev())), ref) <= REL


def test_csr_build_is_sorted_and_keeps_duplicates():
    r, c, v = _random_coo(500, 400, 4000, 11)
    graph = cb.csr_from_coo(r.to(dev()), c.to(dev()), v.to(dev()), 500, 400)
    rp, col, val = graph.rowptr.cpu().numpy(), graph.col.cpu().numpy(), graph.val.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == r.numel() and np.all(np.diff(rp) == np.bincount(r.numpy(), minlength=500))
    order = np.lexsort((np.arange(r.numel()), c.numpy(), r.numpy()))     # stable by (row, col)
    assert np.array_equal(col, c.numpy()[order].astype(np.int32))
    assert np.array_equal(val, v.numpy()[order])
    empty = cb.csr_from_coo(r[:0].to(dev()), c[:0].to(dev()), None, 7, 7)
    assert empty.rowptr.tolist() == [0] * 8


def test_propagate_mean_variants_and_autograd():
    U, I, d, L = 400, 300, 64, 3
    tu, ti = ref_ops.synth_interactions(U, I, 3000, 5)
    train = torch.stack((tu, ti + U), 1)
    ei = ref_ops.symmetrise_edges(train.numpy())
    x0 = ref_ops.xavier_embeddings(U + I, d, 9)
    # reference-side: autograd through the restated PyG conv on the CPU
    xr = x0.clone().requires_grad_(True)
    out_ref = ref_ops.lightgcn_forward(xr[:U], xr[U:], ei, L)
    w = torch.randn(U + I, d, generator=torch.Generator().manual_seed(1))
    (out_ref * w).sum().backward()
    xg = x0.clone().to(dev()).requires_grad_(True)
    out = cb.lightgcn_forward(xg[:U], xg[U:], ei.to(dev()), L)
    (out * w.to(dev())).sum().backward()
    assert rel_err(out, out_ref.detach()) <= REL
    assert rel_err(xg.grad, xr.grad) <= REL
    # mean without layer 0 (Model/SimGCL.py:112-120) on an explicit COO adjacency, non-symmetric values
    n = U + I
    vals = torch.rand(ei.shape[1], generator=torch.Generator().manual_seed(2))
    adj = torch.sparse_coo_tensor(ei, vals, (n, n))
    xr = x0.clone().requires_grad_(True)
    e, acc = xr, []
    for _ in range(2):
        e = torch.sparse.mm(adj, e)
        acc.append(e)
    ref = torch.stack(acc, 1).mean(1)
    (ref * w).sum().backward()
    xg = x0.clone().to(dev()).requires_grad_(True)
    got = cb.propagate_mean(adj.to(dev()), xg, 2, include_layer0=False)
    (got * w.to(dev())).sum().backward()
    assert rel_err(got, ref.detach()) <= REL
    assert rel_err(xg.grad, xr.grad) <= REL


def test_sports_shaped_propagation_matches_oracle():
    """BASELINE.json configs[1] shape: 35,598 users x 18,357 items, ~296k interactions, d=64, L=3."""
    U, I, d, L = 35598, 18357, 64, 3
    tu, ti = ref_ops.synth_interactions(U, I, 296337, 20262)
    ei = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy())
    uw, iw = ref_ops.xavier_embeddings(U, d, 1), ref_ops.xavier_embeddings(I, d, 2)
    ref = ref_ops.lightgcn_forward(uw, iw, ei, L)
    got = cb.lightgcn_forward(uw.to(dev()), iw.to(dev()), ei.to(dev()), L)
    assert rel_err(got, ref) <= REL
    # FREEDOM's halved V2 matrix through the sparse.mm seam
    adj = ref_ops.freedom_norm_adj(ei, U + I)
    ref2 = ref_ops.sparse_mm_chain_mean(adj, torch.cat((uw, iw)), 2)
    got2 = cb.propagate_mean(adj.to(dev()), torch.cat((uw, iw)).to(dev()), 2)
    assert rel_err(got2, ref2) <= REL


def test_spmm_rejects_bad_shapes():
    r, c, v = _random_coo(50, 60, 200, 1, dup=False)
    adj = torch.sparse_coo_tensor(torch.stack((r, c)), v, (50, 60)).to(dev())
    with pytest.raises(RuntimeError):
        cb.spmm(adj, torch.zeros(61, 64, device=dev()))
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.spmm(adj, torch.zeros(60, 66, device=dev()))
    with pytest.raises(RuntimeError, match="CUDA"):
        cb.spmm(adj, torch.zeros(60, 64))


def test_device_side_adjacency_builders_match_reference_fixture():
    """V1/V2/V3 graphs built on the device reproduce the reference's results."""
    g = load_golden("freedom_small.npz")
    U, I = int(g["num_user"]), int(g["num_item"])
    uw, iw = torch.from_numpy(g["user_w"]).to(dev()), torch.from_numpy(g["item_w"]).to(dev())
    x0 = torch.cat((uw, iw), 0)
    mm = torch.sparse_coo_tensor(torch.from_numpy(g["mm_adj_idx"]), torch.from_numpy(g["mm_adj_val"]), (I, I)).to(dev())
    h = cb.spmm(mm, iw)
    train = torch.from_numpy(g["train"].astype(np.int64)).to(dev())
    # V2: FREEDOM.norm_adj (halved values)
    v2 = cb.bipartite_graph(train[:, 0], train[:, 1] - U, U, I, "v2")
    res = cb.propagate_mean(v2, x0, int(g["n_layers"]))
    assert rel_err(torch.cat((res[:U], res[U:] + h)), torch.from_numpy(g["result_norm"])) <= REL
    # V3: the per-epoch masked graph, from the reference's own kept edges
    midx = torch.from_numpy(g["masked_adj_idx"])
    keep = midx[:, : midx.shape[1] // 2].to(dev())
    v3 = cb.bipartite_graph(keep[0], keep[1] - U, U, I, "v3")
    res = cb.propagate_mean(v3, x0, int(g["n_layers"]))
    assert rel_err(torch.cat((res[:U], res[U:] + h)), torch.from_numpy(g["result_masked"])) <= REL
    # V1 == the PyG conv graph
    l = load_golden("lightgcn_small.npz")
    U1, I1 = int(l["num_user"]), int(l["num_item"])
    tr = torch.from_numpy(l["train"].astype(np.int64)).to(dev())
    v1 = cb.bipartite_graph(tr[:, 0], tr[:, 1] - U1, U1, I1, "v1")
    x1 = torch.from_numpy(np.concatenate([l["user_w"], l["item_w"]])).to(dev())
    assert rel_err(cb.propagate_mean(v1, x1, int(l["n_layers"])), torch.from_numpy(l["result"])) <= REL
    # V4 / V5 against the adjacency the UNMODIFIED reference methods built (tests/golden/adjacency_v4_v5.npz)
    a = load_golden("adjacency_v4_v5.npz")
    Ua, Ia = int(a["num_user"]), int(a["num_item"])
    xa = ref_ops.xavier_embeddings(Ua + Ia, 64, 5)
    for variant in ("v4", "v5"):
        ref_adj = torch.sparse_coo_tensor(torch.from_numpy(a[variant + "_index"]), torch.from_numpy(a[variant + "_value"]), (Ua + Ia, Ua + Ia))
        gv = cb.bipartite_graph(torch.from_numpy(a["train_u"]).to(dev()), torch.from_numpy(a["train_i"]).to(dev()), Ua, Ia, variant)
        assert rel_err(cb.spmm(gv, xa.to(dev())), torch.sparse.mm(ref_adj, xa)) <= REL
        # duplicated training pairs collapse to one entry in the scipy dok / lil builders
        gd = cb.bipartite_graph(torch.from_numpy(np.concatenate((a["train_u"], a["train_u"][:50]))).to(dev()),
                                torch.from_numpy(np.concatenate((a["train_i"], a["train_i"][:50]))).to(dev()), Ua, Ia, variant)
        assert torch.equal(cb.spmm(gd, xa.to(dev())), cb.spmm(gv, xa.to(dev())))
    # ... and against the oracle restatements on the LightGCN fixture graph
    tu4, ti4 = tr[:, 0].cpu(), (tr[:, 1] - U1).cpu()
    for variant, build in (("v4", ref_ops.scipy_norm_adj_v4), ("v5", ref_ops.scipy_norm_adj_v5)):
        gv = cb.bipartite_graph(tr[:, 0], tr[:, 1] - U1, U1, I1, variant)
        want = torch.sparse.mm(build(tu4.numpy(), ti4.numpy(), U1, I1), x1.cpu())
        assert rel_err(cb.spmm(gv, x1), want) <= REL
    # resampling keeps int(E*(1-p)) distinct edges and renormalises them (structure-level check)
    vals = torch.ones(train.shape[0], device=dev())
    rg = cb.resample_masked_graph(train[:, 0], train[:, 1] - U, vals, U, I, 0.3, torch.Generator(device=dev()).manual_seed(1))
    assert rg.nnz == 2 * int(train.shape[0] * 0.7)
    deg = (rg.rowptr[1:] - rg.rowptr[:-1]).float()
    want = (deg + 1e-7).pow(-0.5)          # like the reference, also on empty rows (never indexed)
    assert torch.allclose(rg.row_scale, want, rtol=1e-6, atol=0)


def test_bpr_training_step_matches_reference_fixture():
    """BASELINE configs[1] "BPR step": loss and parameter gradients of the reference LightGCN.loss
    (Model/LightGCN.py:97-135 + autograd) through the fused forward/backward."""
    g = load_golden("lightgcn_small.npz")
    s = load_golden("lightgcn_train_step.npz")
    uw = torch.from_numpy(g["user_w"]).to(dev()).requires_grad_(True)
    iw = torch.from_numpy(g["item_w"]).to(dev()).requires_grad_(True)
    ei = torch.from_numpy(g["edge_index"]).to(dev())
    loss, emb = cb.lightgcn_loss(uw, iw, ei, int(s["n_layers"]), torch.from_numpy(s["users"]), torch.from_numpy(s["pos"]),
                                 torch.from_numpy(s["neg"]), float(s["reg_weight"]))
    loss.backward()
    assert abs(float(loss) - float(s["loss"])) <= 1e-6 * abs(float(s["loss"]))
    assert rel_err(uw.grad, torch.from_numpy(s["grad_user"])) <= REL
    assert rel_err(iw.grad, torch.from_numpy(s["grad_item"])) <= REL
    assert rel_err(emb, torch.from_numpy(g["result"])) <= REL
    # one Adam step on top of it stays in step with the reference optimiser (main.py:397)
    ref_u = torch.from_numpy(g["user_w"]).clone().requires_grad_(True)
    ref_u.grad = torch.from_numpy(s["grad_user"]).clone()
    torch.optim.Adam([ref_u], lr=1e-3).step()
    torch.optim.Adam([uw], lr=1e-3).step()
    assert rel_err(uw, ref_u.detach()) <= REL


def test_cuda_graph_training_step_matches_eager():
    U, I, d, L, B = 500, 400, 64, 2, 128
    tu, ti = ref_ops.synth_interactions(U, I, 4000, 21)
    ei = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy()).to(dev())
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randint(0, U, (B,), generator=g).to(dev()), (torch.randint(0, I, (B,), generator=g) + U).to(dev()),
                (torch.randint(0, I, (B,), generator=g) + U).to(dev())) for _ in range(4)]

    def params():
        return (ref_ops.xavier_embeddings(U, d, 1).to(dev()).requires_grad_(True),
                ref_ops.xavier_embeddings(I, d, 2).to(dev()).requires_grad_(True))

    uw, iw = params()
    opt = torch.optim.Adam([uw, iw], lr=1e-2, capturable=True)
    losses = []
    for u, p, n in batches:
        opt.zero_grad(set_to_none=False)
        loss, _ = cb.lightgcn_loss(uw, iw, ei, L, u, p, n, 1e-4)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    uw2, iw2 = params()
    opt2 = torch.optim.Adam([uw2, iw2], lr=1e-2, capturable=True)
    step = cb.GraphedLightGCNStep(uw2, iw2, ei, L, 1e-4, B, opt2)
    for (u, p, n), want in zip(batches, losses):
        gl = step(u, p, n)
        assert abs(float(gl) - want) <= 1e-5 * abs(want)
    torch.cuda.synchronize()
    assert rel_err(uw2, uw.detach().cpu()) <= 1e-5 and rel_err(iw2, iw.detach().cpu()) <= 1e-5


def test_negative_sampler_bit_exact_vs_oracle():
    U, I = 300, 97
    tu, ti = ref_ops.synth_interactions(U, I, 6000, 31)
    ti[tu == 5] = ti[tu == 5] % 3                       # a user with a tiny train set
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    sets = {u: set() for u in range(U)}
    for u, i in zip(tu.tolist(), ti.tolist()):
        sets[u].add(i)
    users = torch.randint(0, U, (5000,), generator=torch.Generator().manual_seed(4))
    got = cb.sample_negatives(users.to(dev()), mask, I, seed=12345, first_draw=777, item_offset=U)
    want = ref_ops.sample_negatives(users.tolist(), sets, I, 12345, first_draw=777, item_offset=U)
    assert got.cpu().tolist() == want
    assert all((g - U) not in sets[u] for g, u in zip(got.cpu().tolist(), users.tolist()))
    # batch boundaries do not matter
    part = cb.sample_negatives(users[1000:2000].to(dev()), mask, I, seed=12345, first_draw=1777, item_offset=U)
    assert torch.equal(part, got[1000:2000])


def test_training_batches_cover_every_edge_once_with_valid_negatives():
    U, I = 200, 150
    tu, ti = ref_ops.synth_interactions(U, I, 3000, 33)
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    seen = []
    for users, pos, neg in cb.training_batches(tu.to(dev()), (ti + U).to(dev()), mask, U, I, 256, epoch=3, seed=9):
        assert users.shape == pos.shape == neg.shape and users.numel() <= 256
        assert int(neg.min()) >= U and int(neg.max()) < U + I
        seen.append(torch.stack((users, pos, neg), 1).cpu())
    trip = torch.cat(seen)
    assert sorted((trip[:, 0] * (U + I) + trip[:, 1]).tolist()) == sorted((tu * (U + I) + ti + U).tolist())
    train = set((tu * I + ti).tolist())
    assert all((int(u) * I + int(n) - U) not in train for u, n in zip(trip[:, 0], trip[:, 2]))
    again = torch.cat([torch.stack(b, 1).cpu() for b in cb.training_batches(tu.to(dev()), (ti + U).to(dev()), mask, U, I, 256, epoch=3, seed=9)])
    assert torch.equal(again, trip)
