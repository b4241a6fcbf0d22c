"""GPU parity of the modality kNN graph (csrc/rank_topk.cu kNN entry) through the C ABI.

Bar: neighbour indices (and their order) bit-exact against the canonical oracle and against the
reference's own outputs under tests/golden; kept similarities bit-exact against the oracle."""
import numpy as np
import pytest
import torch

import chaorec_b200 as cb
from conftest import load_golden
from oracle import ref_ops

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("engine")]


def dev():
    return torch.device("cuda", 0)


def test_reference_fixture_conventions_a_and_c():
    g = load_golden("freedom_small.npz")
    for mod in ("v", "t"):
        feats = torch.from_numpy(g[f"{mod}_feat"]).to(dev())
        indices, adj = cb.get_knn_adj_mat(feats, 10)                    # Model/FREEDOM.py:63,66
        assert indices.dtype == torch.int64
        assert torch.equal(indices.cpu(), torch.from_numpy(g[f"knn_a_{mod}_indices"]))
        assert adj.layout == torch.sparse_coo and tuple(adj.shape) == (feats.shape[0], feats.shape[0])
        # O(N*k) torch post-processing on the GPU: pow(-0.5) may differ from the CPU's by an ulp
        assert torch.allclose(adj._values().cpu(), torch.from_numpy(g[f"knn_a_{mod}_values"]), rtol=1e-6, atol=0)
    knn_c = cb.knn_self_excluded(torch.from_numpy(g["v_feat"]).to(dev()), 10)   # gen_hypergraph_u_i.py:41
    assert torch.equal(knn_c.cpu(), torch.from_numpy(g["knn_c_v"]))
    knn_c5 = cb.knn_self_excluded(torch.from_numpy(g["t_feat"]).to(dev()), 5)
    assert torch.equal(knn_c5.cpu(), torch.from_numpy(g["knn_c_t"]))


def test_l2_normalize_bit_exact(canon):
    x = torch.relu(torch.randn(777, 384, generator=torch.Generator().manual_seed(1)))
    x[5] *= 1e-3
    cn, norms = cb.l2_normalize(x.to(dev()))
    assert np.array_equal(cn.cpu().numpy().view(np.uint32), canon.l2_normalize(x.numpy()).view(np.uint32))
    want = np.sqrt((x.double() ** 2).sum(1).numpy()).astype(np.float32)
    assert np.max(np.abs(norms.cpu().numpy() - want) / want) < 2e-7


@pytest.mark.parametrize("N,D,k,excl", [(2500, 384, 10, 0), (1300, 512, 10, 1), (900, 4096, 10, 0), (700, 64, 20, 0),
                                        (257, 100, 5, 1)])
def test_seeded_knn_bit_exact_vs_canon(canon, N, D, k, excl):
    g = torch.Generator().manual_seed(N + D)
    x = torch.randn(N, D, generator=g)
    if D >= 512:
        x = torch.relu(x)                                              # CNN-feature-like, non-negative
    val, idx = cb.knn_topk(x.to(dev()), k, exclude_self=bool(excl))
    cval, cidx = canon.knn_cosine_topk(x.numpy(), k, excl)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    assert np.array_equal(val.cpu().numpy().view(np.uint32), cval.view(np.uint32))
    if not excl:
        assert bool((idx[:, 0].cpu() == torch.arange(N)).all())        # self is rank 0


def test_duplicate_rows_tie_break_and_row_sharding(canon):
    N, D, k = 1200, 384, 10
    x = torch.randn(N, D, generator=torch.Generator().manual_seed(4))
    x[300:330] = x[17]                                                 # exact duplicates -> ties at sim == 1
    val, idx = cb.knn_topk(x.to(dev()), k)
    cval, cidx = canon.knn_cosine_topk(x.numpy(), k, 0)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    assert idx[300].cpu().tolist() == [17] + list(range(300, 309))     # lowest index first among equals
    # item-row sharding (multi-GPU split): rows [lo, hi) of the same matrix
    parts = [cb.knn_topk(x.to(dev()), k, row_range=(lo, hi)) for lo, hi in ((0, 400), (400, 401), (401, 1200))]
    assert torch.equal(torch.cat([p[1] for p in parts]), idx)
    assert torch.equal(torch.cat([p[0] for p in parts]), val)


def test_convention_b_dense_and_sparse_laplacian():
    feats = torch.randn(300, 64, generator=torch.Generator().manual_seed(6))
    val_ref, ind_ref, dense_ref = ref_ops.knn_conv_b(feats, 10)
    dense = cb.build_knn_neighbourhood(feats.to(dev()), 10)            # Model/LATTICE.py:46-52
    assert torch.equal((dense != 0).cpu(), dense_ref != 0)
    assert float((dense.cpu() - dense_ref).abs().max()) <= 1e-6
    val, ind = cb.build_knn_neighbourhood(feats.to(dev()), 10, dense=False)
    lap = cb.knn_sym_laplacian(val, ind).to_dense().cpu()
    lap_ref = ref_ops.dense_sym_laplacian(dense_ref)
    assert float((lap - lap_ref).abs().max()) <= 1e-5 * float(lap_ref.abs().max())


def test_clothing_shaped_text_modality_sampled(canon):
    """BASELINE.json configs[2]: 23,033 items, t_feat 384-d, k=10 -- full build on the GPU, a seeded
    sample of rows re-derived by the oracle must match bit for bit."""
    N, D, k = 23033, 384, 10
    g = torch.Generator().manual_seed(20263)
    x = torch.randn(N, D, generator=g)
    x[5000:5230] = x[123]                                              # 1 % exact-duplicate block
    val, idx = cb.knn_topk(x.to(dev()), k)
    assert bool((val[:, 1:] <= val[:, :-1]).all())
    for lo in (0, 5000, 23000):
        hi = min(lo + 33, N)
        cval, cidx = canon.knn_cosine_topk(x.numpy(), k, 0, lo, hi)
        assert np.array_equal(idx[lo:hi].cpu().numpy(), cidx)
        assert np.array_equal(val[lo:hi].cpu().numpy().view(np.uint32), cval.view(np.uint32))


def test_knn_bad_arguments():
    x = torch.randn(40, 64, device=dev())
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.knn_topk(x, 65)
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.knn_topk(x[:, :62].contiguous(), 5)
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.knn_topk(x, 40, exclude_self=True)


def test_lattice_item_graph_sparse_operator_matches_reference():
    """Model/LATTICE.py forward(build_item_graph=True) run unmodified (tests/golden/lattice_small.npz) against
    the O(N*k) path: kNN kernel for the indices, SDDMM kernel for the (differentiable) similarity values,
    SpMM kernel for item_adj @ h, fused propagation for the user-item graph.  Values within 1e-5, gradients
    of every parameter within 1e-4 relative (fp32 sums in a different order)."""
    from chaorec_b200 import item_graph as ig
    g = load_golden("lattice_small.npz")
    U, I, k = int(g["U"]), int(g["I"]), int(g["k"])
    names = ("user", "item", "image", "text", "image_w", "image_b", "text_w", "text_b", "modal")
    p = {n: torch.from_numpy(g["p_" + n]).to(dev()).requires_grad_(True) for n in names}
    image_feats = p["image"] @ p["image_w"].t() + p["image_b"]
    text_feats = p["text"] @ p["text_w"].t() + p["text_b"]
    weight = torch.softmax(p["modal"], dim=0)
    graph = ig.ItemGraph([ig.knn_graph(image_feats, k), ig.knn_graph(text_feats, k)],
                         [ig.original_graph(p["image"], k), ig.original_graph(p["text"], k)], weight, float(g["lambda_coeff"]))
    h = p["item"]
    for _ in range(int(g["mm_layers"])):
        h = graph.mm(h)
    ei = ref_ops.symmetrise_edges(g["train"]).to(dev())
    mean = cb.lightgcn_forward(p["user"], p["item"], ei, int(g["n_layers"]))
    u_g, i_g = torch.split(mean, [U, I], dim=0)
    result = torch.cat((u_g, i_g + torch.nn.functional.normalize(h, p=2, dim=1)), dim=0)
    assert float((result.detach().cpu() - torch.from_numpy(g["result"])).abs().max()) <= 1e-5
    (result * torch.from_numpy(g["R"]).to(dev())).sum().backward()
    for n, t in p.items():
        ref = torch.from_numpy(g["g_" + n])
        err = float((t.grad.cpu() - ref).abs().max())
        assert err <= 1e-4 * max(1.0, float(ref.abs().max())), (n, err)
    # the detached operator (Model/LATTICE.py:161) gives the same product and carries no graph
    hd = graph.detach().mm(p["item"].detach())
    assert not hd.requires_grad and float((hd - graph.mm(p["item"]).detach()).abs().max()) == 0.0


def test_user_cooccurrence_graph_matches_reference_and_oracle():
    g = load_golden("cooccurrence_small.npz")
    U, I, k = int(g["U"]), int(g["I"]), int(g["k"])
    train = torch.from_numpy(g["train"])
    idx, cnt, n = cb.user_cooccurrence_topk(train[:, 0].to(dev()), (train[:, 1] - U).to(dev()), U, I, k)
    want = ref_ops.user_graph_topk(g["matrix"].astype(np.int64), k)
    got = cb.user_graph_dict(idx, cnt, n)
    assert got == want                                                      # ids and counts, canonical order
    assert np.array_equal(n.cpu().numpy(), g["topk_n"])
    for u in range(U):                                                      # the reference's own values
        assert got[u][1] == g["topk_values"][u, :len(got[u][1])].tolist()
    # a bigger seeded graph against the oracle, k = 200 like the script, and row sharding
    U2, I2 = 3000, 800
    tu, ti = ref_ops.synth_interactions(U2, I2, 30000, 882)
    idx, cnt, n = cb.user_cooccurrence_topk(tu.to(dev()), ti.to(dev()), U2, I2, 200)
    C = ref_ops.user_cooccurrence_matrix(np.stack((tu.numpy(), ti.numpy()), 1), U2)
    assert cb.user_graph_dict(idx, cnt, n) == ref_ops.user_graph_topk(C, 200)
    idx_b, cnt_b, n_b = cb.user_cooccurrence_topk(tu.to(dev()), ti.to(dev()), U2, I2, 200, row_range=(1000, 2000))
    assert torch.equal(idx_b[1000:2000], idx[1000:2000]) and torch.equal(cnt_b[1000:2000], cnt[1000:2000])
    assert int(n_b[:1000].sum()) == 0 and int(n_b[2000:].sum()) == 0
