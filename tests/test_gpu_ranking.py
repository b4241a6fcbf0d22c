"""GPU parity of the full-ranking evaluation (csrc/rank_topk.cu, csrc/hits_metrics.cu).

Bar: top-K item ids bit-exact against the canonical oracle (oracle/canon.c: fp32 fmaf chain,
descending score, ties -> ascending index) and against the reference's own lists stored under
tests/golden; Recall/NDCG/... identical (==) to utils.gene_metrics."""
import types

import numpy as np
import pytest
import torch

import chaorec_b200 as cb
from conftest import dict_from_train, load_golden, metrics_array, rows_to_object_array, unflatten_rows
from oracle import ref_ops

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("engine")]


def dev():
    return torch.device("cuda", 0)


class _Model:
    """The four attributes gene_ranklist reads off a reference model (Model/LightGCN.py:137-152)."""

    def __init__(self, result, num_user, num_item, user_item_dict):
        self.result, self.num_user, self.num_item, self.user_item_dict = result, num_user, num_item, user_item_dict

    def eval(self):
        return self


def _csr_np(mask):
    return (None, None) if mask is None else (mask.rowptr.cpu().numpy(), mask.col.cpu().numpy())


def _check_against_canon(canon, u, it, mask, K, mask_value=1e-6, id_offset=0):
    idx, sc = cb.rank_topk(u.to(dev()), it.to(dev()), mask, K, mask_value, id_offset, return_scores=True)
    rp, col = _csr_np(mask)
    cidx, csc = canon.score_mask_topk(u.numpy(), it.numpy(), rp, col, mask_value, K, id_offset)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), csc.view(np.uint32))
    return idx


def test_reference_fixture_ranklists_bit_exact():
    for name in ("lightgcn_small.npz", "freedom_small.npz"):
        g = load_golden(name)
        U, I = int(g["num_user"]), int(g["num_item"])
        res = torch.from_numpy(g["result"] if "result" in g else g["result_norm"])
        model = _Model(res.to(dev()), U, I, dict_from_train(g["train"]))
        model.gene_ranklist = types.MethodType(cb.gene_ranklist, model)      # bound like the reference method
        rank = model.gene_ranklist()
        assert rank.dtype == torch.int64 and rank.device.type == "cpu" and tuple(rank.shape) == (U, 50)
        assert torch.equal(rank, torch.from_numpy(g["ranklist"]))
        assert torch.equal(cb.gene_ranklist(model, topk=20), torch.from_numpy(g["ranklist"])[:, :20])
        # CPU-resident result (the reference keeps it wherever training ran)
        model_cpu = _Model(res, U, I, model.user_item_dict)
        assert torch.equal(cb.gene_ranklist(model_cpu, topk=50), rank)


def test_reference_fixture_metrics_identical():
    g = load_golden("lightgcn_small.npz")
    ks = [int(k) for k in g["ks"]]
    rank = torch.from_numpy(g["ranklist"])
    for rows_key, m_key in (("val_rows", "metrics_val"), ("test_rows", "metrics_test")):
        rows = rows_to_object_array(g[rows_key])
        for rl in (rank, rank.to(dev())):
            m = cb.evaluate(None, rows, rl, ks)
            assert metrics_array(m, ks).tolist() == g[m_key].tolist()
    a = load_golden("metrics_adversarial.npz")
    ks = [int(k) for k in a["ks"]]
    m = cb.evaluate(None, unflatten_rows(a["rows_flat"]), torch.from_numpy(a["ranklist"]), ks)
    assert metrics_array(m, ks).tolist() == a["metrics"].tolist()


@pytest.mark.parametrize("U,I,d,K", [(3000, 5000, 64, 20), (1111, 777, 64, 50), (129, 130, 64, 64), (500, 4100, 32, 10),
                                     (300, 1500, 128, 20), (260, 900, 200, 5), (128, 128, 64, 1)])
def test_seeded_scores_bit_exact_vs_canon(canon, U, I, d, K):
    u = ref_ops.normal_embeddings(U, d, 100 + U)
    it = ref_ops.normal_embeddings(I, d, 200 + I)
    tu, ti = ref_ops.synth_interactions(U, I, 8 * U, 300 + U)
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    _check_against_canon(canon, u, it, mask, K, id_offset=U)
    _check_against_canon(canon, u, it, None, K)
    _check_against_canon(canon, u, it, mask, K, mask_value=1e-5)          # GRCN/MMGCN/LightGT value


def test_masked_items_outrank_negative_scores(canon):
    """masked = +1e-6, so train items beat every negatively scored item."""
    U, I, d, K = 200, 300, 64, 50
    u = ref_ops.normal_embeddings(U, d, 1).abs()
    it = -ref_ops.normal_embeddings(I, d, 2).abs()                       # every score negative
    tu, ti = ref_ops.synth_interactions(U, I, 10 * U, 3)
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    idx = _check_against_canon(canon, u, it, mask, K).cpu().numpy()
    rp, col = _csr_np(mask)
    for uu in range(U):
        m = col[rp[uu]:rp[uu + 1]]
        n = min(len(m), K)
        assert idx[uu, :n].tolist() == m[:n].tolist()                   # masked first, ascending index


def test_heavy_users_ties_and_degenerate_inputs(canon):
    U, I, d, K = 140, 400, 64, 50
    u = ref_ops.normal_embeddings(U, d, 5)
    it = ref_ops.normal_embeddings(I, d, 6)
    it[100:164] = it[7]                                                # 64 duplicate item rows -> exact ties
    users = torch.arange(U).repeat_interleave(I - 10)
    items = torch.arange(I - 10).repeat(U)                             # every user trained on all but 10 items
    mask = cb.mask_csr_from_edges(users.to(dev()), items.to(dev()), U, I)
    _check_against_canon(canon, u, it, mask, K)
    _check_against_canon(canon, u, it, None, K)
    z = torch.zeros(U, d)
    idx = _check_against_canon(canon, z, it, None, K)                  # all scores tie at 0
    assert idx.cpu()[0].tolist() == list(range(K))
    full = cb.mask_csr_from_edges(torch.arange(U).repeat_interleave(I).to(dev()), torch.arange(I).repeat(U).to(dev()), U, I)
    idx = _check_against_canon(canon, u, it, full, K)                  # everything masked
    assert idx.cpu()[3].tolist() == list(range(K))


def test_adversarial_ascending_scores_force_many_merges(canon):
    """Scores increasing with the item index: every tile beats the current threshold."""
    U, I, d, K = 130, 20000, 64, 20
    u = torch.zeros(U, d)
    u[:, 0] = 1.0
    it = torch.zeros(I, d)
    it[:, 0] = torch.arange(I, dtype=torch.float32) * 1e-3
    it[:, 1:] = ref_ops.normal_embeddings(I, d - 1, 8)
    idx = _check_against_canon(canon, u, it, None, K)
    assert idx.cpu()[0].tolist() == list(range(I - 1, I - 1 - K, -1))


def test_bad_arguments_raise():
    u = torch.zeros(8, 64, device=dev())
    it = torch.zeros(16, 64, device=dev())
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.rank_topk(u, it, None, 65)
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.rank_topk(u, it, None, 17)                                   # K > I
    with pytest.raises(RuntimeError, match="unsupported"):
        cb.rank_topk(torch.zeros(8, 66, device=dev()), torch.zeros(16, 66, device=dev()), None, 4)


def test_baby_and_sports_shaped_eval_end_to_end(canon):
    """BASELINE.json configs[0]/[1] shapes: propagate -> rank -> metrics vs the oracle chain."""
    for (U, I, E, seed) in ((19445, 7050, 160792, 20261), (35598, 18357, 296337, 20262)):
        tu, ti = ref_ops.synth_interactions(U, I, E, seed)
        heldout = ref_ops.synth_holdout(U, I, tu, ti, 2, seed + 2)
        u = ref_ops.normal_embeddings(U, 64, seed + 10)
        it = ref_ops.normal_embeddings(I, 64, seed + 11)
        ei = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy())
        res = cb.lightgcn_forward(u.to(dev()), it.to(dev()), ei.to(dev()), 3)
        mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
        rank = cb.rank_topk(res[:U], res[U:], mask, 50, 1e-6, id_offset=U)
        res_h = res.cpu().numpy()
        rp, col = _csr_np(mask)
        cidx, _ = canon.score_mask_topk(res_h[:U], res_h[U:], rp, col, 1e-6, 50, id_offset=U)
        assert np.array_equal(rank.cpu().numpy(), cidx)
        # metrics: reference python loop (restated) on a 3000-user slice vs the fused path
        rows = np.empty(3000, dtype=object)
        held = (heldout[:3000] + U).tolist()
        for r in range(3000):
            rows[r] = [r] + held[r]
        ks = [5, 10, 20]
        got = cb.evaluate(None, rows, rank, ks)
        want = ref_ops.gene_metrics(rows, torch.from_numpy(cidx), ks)
        assert metrics_array(got, ks).tolist() == metrics_array(want, ks).tolist()


def test_million_user_config_sampled_bit_exact(canon):
    """BASELINE.json configs[3]: 1M users x 200k items, d=64, top-20.  The whole job runs on the GPU;
    a seeded sample of users is re-derived by the oracle and must match bit for bit, and the
    size-independent properties (descending scores, unique in-range ids) are checked for all."""
    U, I, d, K = 1_000_000, 200_000, 64, 20
    u = ref_ops.normal_embeddings(U, d, 20264)
    it = ref_ops.normal_embeddings(I, d, 20265)
    g = torch.Generator().manual_seed(20266)
    cdf, perm = ref_ops.zipf_item_sampler(I, g)
    mu = torch.arange(U).repeat_interleave(16)
    mi = perm[torch.searchsorted(cdf, torch.rand(mu.numel(), generator=g, dtype=torch.float64)).clamp_(max=I - 1)]
    mask = cb.mask_csr_from_edges(mu.to(dev()), mi.to(dev()), U, I)
    st = {}
    idx, sc = cb.rank_topk(u.to(dev()), it.to(dev()), mask, K, 1e-6, id_offset=U, return_scores=True, stats=st)
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())
    assert int(idx.min()) >= U and int(idx.max()) < U + I
    srt = torch.sort(idx, dim=1)[0]
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # 2048 seeded users + every row the tensor engine handed to the exact engine
    sample = torch.randperm(U, generator=g)[:2048]
    if st.get("fallback_row_ids") is not None:
        assert st["fallback_rows"] == st["fallback_row_ids"].numel()
        sample = torch.unique(torch.cat((sample, st["fallback_row_ids"][:2048])))
    sample = sample.sort()[0]
    rp, col = _csr_np(mask)
    sub_rp = np.zeros(sample.numel() + 1, dtype=np.int64)
    cols = []
    for j, uu in enumerate(sample.tolist()):
        cols.append(col[rp[uu]:rp[uu + 1]])
        sub_rp[j + 1] = sub_rp[j] + len(cols[-1])
    cidx, csc = canon.score_mask_topk(u[sample].numpy(), it.numpy(), sub_rp, np.concatenate(cols), 1e-6, K, id_offset=U)
    assert np.array_equal(idx[sample.to(dev())].cpu().numpy(), cidx)
    assert np.array_equal(sc[sample.to(dev())].cpu().numpy().view(np.uint32), csc.view(np.uint32))


@pytest.mark.parametrize("K", [20, 50])
def test_sample_prepass_thresholds_bit_exact(canon, K):
    """Item tables of 65536+ rows start every row's threshold from a sample pre-pass (tc_candidates.cu):
    all rows must still match the oracle bit for bit, and only a few may need the exact engine."""
    U, I, d = 2500, 70_000, 64
    u = ref_ops.normal_embeddings(U, d, 91)
    it = ref_ops.normal_embeddings(I, d, 92)
    tu, ti = ref_ops.synth_interactions(U, I, 12 * U, 93)
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    st = {}
    idx, sc = cb.rank_topk(u.to(dev()), it.to(dev()), mask, K, 1e-6, id_offset=U, return_scores=True, stats=st)
    rp, col = _csr_np(mask)
    cidx, csc = canon.score_mask_topk(u.numpy(), it.numpy(), rp, col, 1e-6, K, id_offset=U)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), csc.view(np.uint32))
    if st["engine"] == "tensor":
        assert st["fallback_rows"] <= U // 20


def test_tensor_engine_is_used_and_rarely_falls_back(canon, engine):
    """The tcgen05 pre-filter must actually run for the big shapes, certify almost every row, and
    hand the rest to the exact engine (results are bit-exact either way)."""
    U, I, d, K = 5000, 6000, 64, 20
    u = ref_ops.normal_embeddings(U, d, 71)
    it = ref_ops.normal_embeddings(I, d, 72)
    tu, ti = ref_ops.synth_interactions(U, I, 8 * U, 73)
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    stats = {}
    idx = cb.rank_topk(u.to(dev()), it.to(dev()), mask, K, 1e-6, 0, stats=stats)
    assert stats["engine"] == ("tensor" if engine == "auto" else "exact")
    rp, col = _csr_np(mask)
    cidx, _ = canon.score_mask_topk(u.numpy(), it.numpy(), rp, col, 1e-6, K, 0)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    if engine == "auto":
        assert stats["fallback_rows"] <= U // 20
    # adversarial for the pre-filter: 300 near-duplicate items within bf16 resolution of each other
    it2 = it.clone()
    it2[1000:1300] = it[5] * (1.0 + 1e-4 * torch.arange(300).float().unsqueeze(1))
    stats = {}
    idx = cb.rank_topk(u.to(dev()), it2.to(dev()), None, K, 1e-6, 0, stats=stats)
    cidx, _ = canon.score_mask_topk(u.numpy(), it2.numpy(), None, None, 1e-6, K, 0)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    if engine == "auto":
        assert 0 < stats["fallback_rows"] <= 4096          # short list -> column-sliced exact pass
    # every item within bf16 resolution of every other: nothing can be certified, all 5000 rows take
    # the unsliced exact pass
    it3 = it[5].unsqueeze(0) * (1.0 + 1e-5 * torch.arange(I).float().unsqueeze(1))
    stats = {}
    idx = cb.rank_topk(u.to(dev()), it3.to(dev()), mask, K, 1e-6, 0, stats=stats)
    cidx, _ = canon.score_mask_topk(u.numpy(), it3.numpy(), rp, col, 1e-6, K, 0)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    if engine == "auto":
        assert stats["fallback_rows"] > 4096
    # a user with hundreds of train items is handed to the exact engine up front
    heavy = cb.mask_csr_from_edges(torch.cat((tu, torch.zeros(400, dtype=torch.int64))).to(dev()),
                                   torch.cat((ti, torch.arange(400))).to(dev()), U, I)
    idx = cb.rank_topk(u.to(dev()), it.to(dev()), heavy, K, 1e-6, 0)
    hrp, hcol = _csr_np(heavy)
    cidx, _ = canon.score_mask_topk(u.numpy(), it.numpy(), hrp, hcol, 1e-6, K, 0)
    assert np.array_equal(idx.cpu().numpy(), cidx)


def test_host_batch_path_chunked_equals_single():
    """The end-to-end host-buffer API (pinned host -> device copies overlapped with ranking)."""
    import bench
    from chaorec_b200.ranking import evaluate_host_batch, metrics_from_hits
    job = bench.make_eval_job(dict(bench.EVAL_CFG, U=6000, I=3000), 1.0)
    single = bench.shard_eval_job(job, 0, job["U"])
    chunks = bench.shard_eval_job(job, 0, job["U"], n_chunks=3)
    assert isinstance(chunks, list) and len(chunks) == 2
    h1 = evaluate_host_batch(single, job["K"], dev()).numpy().copy()
    h3 = evaluate_host_batch(chunks, job["K"], dev()).numpy().copy()
    assert np.array_equal(h1, h3)
    # the same job through rank_topk + evaluate
    mask = cb.MaskCSR(job["mask_rowptr"].to(dev()), job["mask_col"].to(dev()))
    rank = cb.rank_topk(job["user"].to(dev()), job["item"].to(dev()), mask, job["K"], 1e-6, job["U"])
    rows = np.empty(job["U"], dtype=object)
    for u, p in enumerate(job["pos"].tolist()):
        rows[u] = [u] + p
    want = cb.evaluate(None, rows, rank, [5, 10, 20])
    got = metrics_from_hits(h3, np.full(job["U"], job["pos"].shape[1], dtype=np.int32), [5, 10, 20])
    assert got == want


def test_certification_margin_near_ties(canon, engine):
    """Columns planted just inside / far outside the certification margin of the tensor engine.  A copy of a user's
    K-th best item scaled by (1 - delta) scores s_K (1 - delta).  Near-ties INSIDE the candidate list are resolved by
    the exact re-score, so a row is uncertifiable only when more near-tied columns exist than the list can hold beyond
    K (K' - K = 12): 16 copies at delta = j * 2^-14 (all within the fp16 pre-filter's error bound ~1e-3 |a||b| of the
    K-th score) leave at least four of them outside the list -> the row MUST be re-run by the exact engine.  The same
    16 copies at delta = 0.2 + ... sit far below the margin -> the row stays certified.  Bit-exact either way."""
    U, I, d, K = 1536, 8192, 64, 20
    u = ref_ops.normal_embeddings(U, d, 301)
    it = ref_ops.normal_embeddings(I, d, 302)
    base_idx, base_sc = canon.score_mask_topk(u.numpy(), it.numpy(), None, None, 1e-6, K, 0)
    it2 = it.clone()
    near, far = list(range(0, 8)), list(range(8, 16))
    top_items = set(base_idx[: 16].reshape(-1).tolist())
    free = [c for c in range(I - 1, 0, -1) if c not in top_items][: 16 * 16]     # columns we may overwrite
    for r in near + far:
        assert base_sc[r, K - 1] > 0
        for j in range(16):
            delta = (j + 1) * 2.0 ** -14 if r in near else 0.2 + 0.01 * j
            it2[free[r * 16 + j]] = it[int(base_idx[r, K - 1])] * (1.0 - delta)
    st = {}
    idx, sc = cb.rank_topk(u.to(dev()), it2.to(dev()), None, K, 1e-6, 0, return_scores=True, stats=st)
    cidx, csc = canon.score_mask_topk(u.numpy(), it2.numpy(), None, None, 1e-6, K, 0)
    assert np.array_equal(idx.cpu().numpy(), cidx)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), csc.view(np.uint32))
    if engine == "auto":
        assert st["engine"] == "tensor"
        fb = set(st["fallback_row_ids"].tolist())
        assert st["fallback_rows"] == len(fb)
        assert set(near) <= fb, sorted(set(near) - fb)               # inside the margin: never certified
        assert len(fb & set(far)) <= 1                                # far outside: certified (a natural near-tie aside)
        assert len(fb) <= U // 8


def test_sum_of_inner_products_two_tower_and_three_term(canon):
    """Score-combination variants of gene_ranklist: u.i' + u'.i (Model/SelfCF.py:213-238) and the three-term sum
    (Model/MCLN.py:336-337) are one ranking call on the concatenated embeddings.  Bit-exact against the oracle's chain
    over the concatenated width; the reference's own arithmetic (separate fp32 matmuls, then adds) agrees on every row
    up to near-ties (scores within 1e-6 relative)."""
    U, I, d, K = 700, 3000, 64, 20
    tu, ti = ref_ops.synth_interactions(U, I, 6 * U, 41)
    mask = cb.mask_csr_from_edges(tu.to(dev()), ti.to(dev()), U, I)
    rp, col = _csr_np(mask)
    for n_parts in (2, 3):
        ups = [ref_ops.normal_embeddings(U, d, 50 + p) for p in range(n_parts)]
        ips = [ref_ops.normal_embeddings(I, d, 60 + p) for p in range(n_parts)]
        idx, sc = cb.rank_topk_sum([x.to(dev()) for x in ups], [x.to(dev()) for x in ips], mask, K, 1e-6, id_offset=U, return_scores=True)
        cidx, csc = canon.score_mask_topk(torch.cat(ups, 1).numpy(), torch.cat(ips, 1).numpy(), rp, col, 1e-6, K, id_offset=U)
        assert np.array_equal(idx.cpu().numpy(), cidx)
        assert np.array_equal(sc.cpu().numpy().view(np.uint32), csc.view(np.uint32))
        # the reference's form
        s = sum(torch.matmul(a, b.t()) for a, b in zip(ups, ips))
        s[tu, ti] = 1e-6
        ref_sc, ref_idx = torch.sort(s, dim=1, descending=True, stable=True)
        assert torch.allclose(sc.cpu(), ref_sc[:, :K], rtol=1e-5, atol=1e-7)
        same = (idx.cpu() == ref_idx[:, :K] + U).all(dim=1).float().mean()
        assert float(same) >= 0.995
    with pytest.raises(RuntimeError):
        cb.rank_topk_sum([ups[0].to(dev())], [], mask, K)
