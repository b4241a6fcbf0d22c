"""Generates the parity fixtures under tests/golden from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own classes/functions -- Model/FREEDOM.py as is, Model/LightGCN.py
through the 30-line torch_geometric stand-in of oracle/pyg_stub, utils.gene_metrics,
train_and_evaluate.evaluate, gen_hypergraph_u_i.get_knn_adj_mat -- runs them on seeded CPU
inputs and stores inputs + outputs as small .npz files.  tests/test_oracle_golden.py pins
oracle/ref_ops.py and oracle/canon.c against these files; the -m gpu tests check the CUDA path
against the same files.  Nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CHAOREC_REFERENCE", "/root/reference")

_ARGS = sys.argv[1:]
sys.argv = ["make_golden"]          # arg_parser.parse_args() runs at import time
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_stub"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
os.chdir(REF)

from oracle import ref_ops  # noqa: E402
from oracle.canon_py import Canon  # noqa: E402

import utils as ref_utils  # noqa: E402  (reference)
from Model.LightGCN import LightGCN  # noqa: E402  (reference, via stub)
from Model.FREEDOM import FREEDOM  # noqa: E402  (reference, unchanged)
import train_and_evaluate as ref_te  # noqa: E402
import gen_hypergraph_u_i as ref_hyper  # noqa: E402

canon = Canon()


def dataset(num_user, num_item, n_edges, seed):
    tu, ti = ref_ops.synth_interactions(num_user, num_item, n_edges, seed)
    val = ref_ops.synth_holdout(num_user, num_item, tu, ti, 1, seed + 1)
    test = ref_ops.synth_holdout(num_user, num_item, tu, ti, 2, seed + 2)
    train, uid, val_rows = ref_ops.to_reference_formats(num_user, tu, ti, val)
    _, _, test_rows = ref_ops.to_reference_formats(num_user, tu, ti, test)
    return train, uid, val_rows, test_rows, val.numpy(), test.numpy()


def metrics_to_array(m, ks):
    return np.array([[m[k][n] for n in ('precision', 'recall', 'ndcg', 'hit_rate', 'map')] for k in ks], dtype=np.float64)


def check_rank_agrees(result, U, I, uid, ref_rank, mask_value, what):
    """The reference's torch.matmul+torch.topk list must equal the canonical list on the fixture
    (no sub-ulp near-ties inside the top-K of these seeds); otherwise pick another seed."""
    rp, cols = ref_ops_mask_csr(uid, U, I)
    idx, _ = canon.score_mask_topk(result[:U].numpy(), result[U:U + I].numpy(), rp, cols, mask_value, ref_rank.shape[1], 0)
    same = np.array_equal(idx, ref_rank.numpy())
    print(f"  {what}: reference top-{ref_rank.shape[1]} == canonical: {same}")
    if not same:
        bad = np.nonzero((idx != ref_rank.numpy()).any(1))[0]
        raise SystemExit(f"{what}: {bad.size} users differ (near-ties); choose another seed")


def ref_ops_mask_csr(uid, U, I):
    rows = [sorted(set(int(i) - U for i in uid.get(u, []))) for u in range(U)]
    rowptr = np.zeros(U + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    cols = np.array([c for r in rows for c in r], dtype=np.int32)
    return rowptr, cols


def stable_topk(x, k, dim=-1, **kw):
    """torch.topk with the stated tie-break (descending value, ties -> ascending index)."""
    v, i = torch.sort(x, dim=dim, descending=True, stable=True)
    return v[..., :k], i[..., :k]


def reference_ranklists(model, topk=50):
    """(raw, stable): the reference method as is (torch.topk leaves the order of exactly tied
    scores -- the masked items at 1e-6 -- unspecified), and the same
    reference method with torch.topk swapped for the stated tie-break."""
    raw = model.gene_ranklist(topk=topk)
    orig = torch.topk
    torch.topk = stable_topk
    try:
        stable = model.gene_ranklist(topk=topk)
    finally:
        torch.topk = orig
    # the two may differ only inside runs of exactly equal scores
    U, I = model.num_user, model.num_item
    S = ref_ops.masked_score_matrix(model.result[:U].detach(), model.result[U:U + I].detach(), model.user_item_dict, U)
    assert torch.equal(torch.gather(S, 1, raw - U), torch.gather(S, 1, stable - U))
    print(f"  raw torch.topk order differs from the stated tie-break for {(raw != stable).any(1).sum().item()} users (exact ties)")
    return raw, stable


def lightgcn_case():
    print("LightGCN (reference Model/LightGCN.py via pyg_stub)")
    U, I, d, L = 300, 200, 64, 3
    train, uid, val_rows, test_rows, _, _ = dataset(U, I, 1800, 20261)
    torch.manual_seed(1234)
    model = LightGCN(U, I, train, uid, d, 1e-4, L, 'add', 'cpu')
    with torch.no_grad():
        # trained-like magnitudes so that scores are far above the 1e-6 mask value
        model.user_embedding.weight.copy_(ref_ops.normal_embeddings(U, d, 11))
        model.item_embedding.weight.copy_(ref_ops.normal_embeddings(I, d, 12))
        result = model.forward().clone()
        one_layer = model.conv_layers[0](torch.cat((model.user_embedding.weight, model.item_embedding.weight)),
                                         model.edge_index).clone()
        rank_raw, rank = reference_ranklists(model, 50)
    check_rank_agrees(result, U, I, uid, rank - U, np.float32(1e-6), "LightGCN")
    ks = [5, 10, 20]
    m_val = ref_te.evaluate(model, val_rows, rank, ks)
    m_test = ref_utils.gene_metrics(test_rows, rank, ks)
    m_val_raw = ref_te.evaluate(model, val_rows, rank_raw, ks)
    print("  metrics identical for raw and stable order:", metrics_to_array(m_val, ks).tolist() == metrics_to_array(m_val_raw, ks).tolist())
    np.savez_compressed(
        os.path.join(HERE, "lightgcn_small.npz"), num_user=U, num_item=I, n_layers=L, train=train,
        user_w=model.user_embedding.weight.detach().numpy(), item_w=model.item_embedding.weight.detach().numpy(),
        edge_index=model.edge_index.numpy(), one_layer=one_layer.numpy(), result=result.numpy(),
        ranklist=rank.numpy(), ranklist_raw=rank_raw.numpy(), val_rows=np.array([np.array(r) for r in val_rows.tolist()]),
        test_rows=np.array([np.array(r) for r in test_rows.tolist()]), ks=np.array(ks),
        metrics_val=metrics_to_array(m_val, ks), metrics_test=metrics_to_array(m_test, ks))


def train_step_case():
    """One BPR training step of the reference LightGCN (loss + backward), same model as lightgcn_case."""
    print("LightGCN training step (reference Model/LightGCN.py:97-135 + autograd)")
    U, I, d, L = 300, 200, 64, 3
    train, uid, _, _, _, _ = dataset(U, I, 1800, 20261)
    torch.manual_seed(1234)
    model = LightGCN(U, I, train, uid, d, 1e-4, L, 'add', 'cpu')
    with torch.no_grad():
        model.user_embedding.weight.copy_(ref_ops.normal_embeddings(U, d, 11))
        model.item_embedding.weight.copy_(ref_ops.normal_embeddings(I, d, 12))
    g = torch.Generator().manual_seed(99)
    B = 256
    pick = torch.randint(0, train.shape[0], (B,), generator=g)
    users = torch.from_numpy(train[pick.numpy(), 0].astype(np.int64))
    pos = torch.from_numpy(train[pick.numpy(), 1].astype(np.int64))            # global item ids
    neg = torch.randint(0, I, (B,), generator=g) + U                          # global item ids
    users[:8] = users[0]                                                      # repeated rows inside the batch
    loss = model.loss(users, pos, neg)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "lightgcn_train_step.npz"), num_user=U, num_item=I, n_layers=L, reg_weight=1e-4,
                        users=users.numpy(), pos=pos.numpy(), neg=neg.numpy(), loss=np.float32(loss.item()),
                        grad_user=model.user_embedding.weight.grad.numpy(), grad_item=model.item_embedding.weight.grad.numpy())
    print("  loss", float(loss))


def freedom_case():
    print("FREEDOM (reference Model/FREEDOM.py, unchanged)")
    U, I, d = 260, 180, 64
    train, uid, val_rows, test_rows, _, _ = dataset(U, I, 1500, 20262)
    g = torch.Generator().manual_seed(77)
    v_feat = torch.relu(torch.randn(I, 128, generator=g))
    t_feat = torch.randn(I, 32, generator=g)
    torch.manual_seed(4321)
    model = FREEDOM(U, I, train, uid, v_feat, t_feat, d, 64, 1e-3, 0.3, 2, 1, 10, 0.1, 'cpu')
    with torch.no_grad():
        model.user_embedding.weight.copy_(ref_ops.normal_embeddings(U, d, 21))
        model.item_embedding.weight.copy_(ref_ops.normal_embeddings(I, d, 22))
        model.pre_epoch_processing()
        masked = model.masked_adj
        model.forward(masked)
        result_masked = model.result.clone()
        model.forward(model.norm_adj)
        result_norm = model.result.clone()
        rank_raw, rank = reference_ranklists(model, 50)
        ind_v, adj_v = model.get_knn_adj_mat(v_feat)
        ind_t, adj_t = model.get_knn_adj_mat(t_feat)
        knn_c_v = ref_hyper.get_knn_adj_mat(v_feat, 10)
        knn_c_t = ref_hyper.get_knn_adj_mat(t_feat, 5)
    check_rank_agrees(result_norm, U, I, uid, rank - U, np.float32(1e-6), "FREEDOM")
    # kNN: the reference's index sets must equal the canonical ones on the fixture
    for name, feats, ind, k, excl in (("v", v_feat, ind_v[1].view(I, 10), 10, 0), ("t", t_feat, ind_t[1].view(I, 10), 10, 0),
                                      ("c_v", v_feat, knn_c_v, 10, 1), ("c_t", t_feat, knn_c_t, 5, 1)):
        _, cidx = canon.knn_cosine_topk(feats.numpy(), k, excl)
        same = np.array_equal(cidx, ind.numpy())
        print(f"  kNN {name}: reference == canonical: {same}")
        if not same:
            raise SystemExit("kNN near-tie on the fixture; choose another seed")
    mm = model.mm_adj.coalesce()
    ks = [5, 10, 20]
    m_val = ref_utils.gene_metrics(val_rows, rank, ks)
    np.savez_compressed(
        os.path.join(HERE, "freedom_small.npz"), num_user=U, num_item=I, n_layers=2, mm_layers=1, train=train,
        user_w=model.user_embedding.weight.detach().numpy(), item_w=model.item_embedding.weight.detach().numpy(),
        v_feat=v_feat.numpy(), t_feat=t_feat.numpy(),
        norm_adj_idx=model.norm_adj._indices().numpy(), norm_adj_val=model.norm_adj._values().numpy(),
        masked_adj_idx=masked._indices().numpy(), masked_adj_val=masked._values().numpy(),
        mm_adj_idx=mm.indices().numpy(), mm_adj_val=mm.values().numpy(),
        result_masked=result_masked.numpy(), result_norm=result_norm.numpy(), ranklist=rank.numpy(),
        ranklist_raw=rank_raw.numpy(),
        knn_a_v_indices=ind_v.numpy(), knn_a_v_values=adj_v._values().numpy(),
        knn_a_t_indices=ind_t.numpy(), knn_a_t_values=adj_t._values().numpy(),
        knn_c_v=knn_c_v.numpy(), knn_c_t=knn_c_t.numpy(),
        val_rows=np.array([np.array(r) for r in val_rows.tolist()]), ks=np.array(ks),
        metrics_val=metrics_to_array(m_val, ks))


def metrics_case():
    """utils.gene_metrics on adversarial rows: duplicates in the positives, many positives,
    no hit, all hit, positives > k."""
    print("metrics (reference utils.gene_metrics / metrics.py)")
    g = torch.Generator().manual_seed(5)
    U, I, K = 64, 120, 50
    rank = torch.stack([torch.randperm(I, generator=g)[:K] for _ in range(U)]) + U
    rows = []
    for u in range(U):
        r = rank[u].tolist()
        if u % 6 == 0:
            pos = r[:3] + [r[0]]                      # duplicates inside the positives
        elif u % 6 == 1:
            pos = r[10:40]                            # more positives than k
        elif u % 6 == 2:
            pos = [U + I + 5]                         # never ranked
        elif u % 6 == 3:
            pos = r[:25]
        elif u % 6 == 4:
            pos = [r[4], r[19], r[49], U + I + 9]
        else:
            pos = [r[u % K]]
        rows.append([u] + pos)
    rows = rows[::-1]                                  # rows need not be in user order
    val = np.empty(len(rows), dtype=object)
    for i, r in enumerate(rows):
        val[i] = r
    ks = [1, 5, 10, 20, 50]
    m = ref_utils.gene_metrics(val, rank, ks)
    flat = np.array([v for r in rows for v in [len(r)] + r], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "metrics_adversarial.npz"), ranklist=rank.numpy(), rows_flat=flat,
                        ks=np.array(ks), metrics=metrics_to_array(m, ks))


def real_graph_case():
    """A real interaction graph shipped with the reference (Data/baby/train.npy, first 12 000
    edges) through the reference LightGCN forward: realistic degree skew incl. isolated nodes."""
    print("real graph slice (reference Data/baby/train.npy)")
    train = np.load(os.path.join(REF, "Data/baby/train.npy"), allow_pickle=True)
    U, I = 12351, 4794                                  # dataload.py:45-46
    sl = train[:12000]
    uid = {}
    for u, i in sl.tolist():
        uid.setdefault(u, []).append(i)
    torch.manual_seed(99)
    model = LightGCN(U, I, sl, uid, 64, 1e-4, 2, 'add', 'cpu')
    with torch.no_grad():
        # seeded tables the test can regenerate (ref_ops.xavier_embeddings) instead of storing 4 MB
        model.user_embedding.weight.copy_(ref_ops.xavier_embeddings(U, 64, 31))
        model.item_embedding.weight.copy_(ref_ops.xavier_embeddings(I, 64, 32))
        result = model.forward()
    deg = np.bincount(np.concatenate([sl[:, 0], sl[:, 1]]), minlength=U + I)
    keep = np.unique(np.concatenate([np.arange(0, 64), U + np.arange(0, 64), sl[:64, 0], sl[:64, 1],
                                     np.argsort(-deg)[:64]]))
    np.savez_compressed(os.path.join(HERE, "baby_slice.npz"), num_user=U, num_item=I, n_layers=2, train=sl.astype(np.int32),
                        seed_user=31, seed_item=32, check_rows=keep, result_rows=result.numpy()[keep],
                        result_absmax=np.abs(result.numpy()).max())


def cooccurrence_case():
    """dualgnn-gen-u-u-matrix.py run as is (gen_user_matrix + the top-200 loop body) on a small seeded graph."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dualgnn_gen", os.path.join(REF, "dualgnn-gen-u-u-matrix.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    U, I, E = 150, 60, 1200
    tu, ti = ref_ops.synth_interactions(U, I, E, 881)
    train = np.stack((tu.numpy(), ti.numpy() + U), 1)
    M = mod.gen_user_matrix(train, U)
    k = 20                                     # the script's 200, scaled to the fixture (rows here have up to ~140 entries)
    vals = np.zeros((U, k), dtype=np.float32)
    n = np.zeros(U, dtype=np.int64)
    for i in range(U):                         # dualgnn-gen-u-u-matrix.py:85-103 with 200 -> k
        nz = len(torch.nonzero(M[i]))
        kk = nz if nz <= k else k
        top = torch.topk(M[i], int(kk))
        vals[i, :kk] = top.values.numpy()
        n[i] = kk
    np.savez_compressed(os.path.join(HERE, "cooccurrence_small.npz"), U=U, I=I, k=k, train=train, matrix=M.numpy(), topk_values=vals, topk_n=n)
    print("cooccurrence_small.npz", M.shape, float(M.max()))


def lattice_case():
    """Model/LATTICE.py as is (through the PyG stand-in): forward(build_item_graph=True) and the gradients of
    a fixed linear functional of the result with respect to every parameter the item graph touches."""
    from Model.LATTICE import LATTICE
    U, I, E, dv, dt, fe, dE, k = 120, 200, 1500, 48, 40, 32, 64, 10
    torch.manual_seed(77)
    tu, ti = ref_ops.synth_interactions(U, I, E, 771)
    train = np.stack((tu.numpy(), ti.numpy() + U), 1)
    uid = {u: set() for u in range(U)}
    for u, i in train:
        uid[int(u)].add(int(i))
    g = torch.Generator().manual_seed(772)
    v_feat = torch.randn(I, dv, generator=g)
    t_feat = torch.randn(I, dt, generator=g)
    model = LATTICE(U, I, train, uid, v_feat.clone(), t_feat.clone(), dE, fe, 1e-4, 2, 1, k, 'add', 0.9, torch.device('cpu'))
    with torch.no_grad():
        model.modal_weight.copy_(torch.tensor([0.3, 0.8]))
    R = torch.randn(U + I, dE, generator=g)
    params = dict(user=model.user_embedding.weight, item=model.item_embedding.weight, image=model.image_embedding.weight,
                  text=model.text_embedding.weight, image_w=model.image_trs.weight, image_b=model.image_trs.bias,
                  text_w=model.text_trs.weight, text_b=model.text_trs.bias, modal=model.modal_weight)
    start = {n: p.detach().clone().numpy() for n, p in params.items()}
    result = model.forward(build_item_graph=True)
    (result * R).sum().backward()
    out = dict(U=U, I=I, k=k, lambda_coeff=0.9, mm_layers=1, n_layers=2, train=train, R=R.numpy(), result=result.detach().numpy(),
               item_adj=model.item_adj.detach().numpy(), image_original_adj=model.image_original_adj.numpy(),
               text_original_adj=model.text_original_adj.numpy())
    for n, p in params.items():
        out["p_" + n] = start[n]
        out["g_" + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "lattice_small.npz"), **out)
    print("lattice_small.npz", result.shape, float(result.abs().max()))
def scipy_adjacency_case():
    """Adjacency builders V4 / V5 run by the UNMODIFIED reference methods (Model/SimGCL.py:64-104 `get_norm_adj_mat`,
    Model/MGCN.py:164-196 `get_adj_mat`) on a seeded interaction matrix.  The methods only read num_user, num_item,
    interaction_matrix and device off `self`, so they are called unbound on a namespace (building the whole models needs
    feature files).  SimGCL's builder calls `dok_matrix._update`, which newer scipy releases removed: a five-line shim
    restores exactly what it did (a bulk dict update of the (row, col) -> value entries)."""
    import types
    import scipy.sparse as sp
    if not hasattr(sp.dok_matrix, "_update"):
        def _update(self, data):
            for (r, c), v in data.items():
                self[r, c] = v
        sp.dok_matrix._update = _update
    from Model.SimGCL import SimGCL
    from Model.MGCN import MGCN
    U, I = 300, 120
    tu, ti = ref_ops.synth_interactions(U, I, 2200, 17)
    tu, ti = tu.numpy(), ti.numpy()
    R = sp.coo_matrix((np.ones(tu.size, dtype=np.float32), (tu, ti)), shape=(U, I))
    ns = types.SimpleNamespace(num_user=U, num_item=I, interaction_matrix=R, device=torch.device("cpu"))
    v4 = SimGCL.get_norm_adj_mat(ns).coalesce()
    v5 = sp.coo_matrix(MGCN.get_adj_mat(ns)).astype(np.float32)
    # the oracle restatements must reproduce both bit for bit
    o4 = ref_ops.scipy_norm_adj_v4(tu, ti, U, I).coalesce()
    o5 = ref_ops.scipy_norm_adj_v5(tu, ti, U, I).coalesce()
    assert torch.equal(o4.indices(), v4.indices()) and torch.equal(o4.values(), v4.values()), "V4 restatement differs"
    v5t = torch.sparse_coo_tensor(torch.tensor(np.array([v5.row, v5.col]), dtype=torch.long), torch.tensor(v5.data), (U + I, U + I)).coalesce()
    assert torch.equal(o5.indices(), v5t.indices()) and torch.equal(o5.values(), v5t.values()), "V5 restatement differs"
    np.savez_compressed(os.path.join(HERE, "adjacency_v4_v5.npz"), num_user=U, num_item=I, train_u=tu, train_i=ti,
                        v4_index=v4.indices().numpy(), v4_value=v4.values().numpy(),
                        v5_index=v5t.indices().numpy(), v5_value=v5t.values().numpy())
    print("adjacency_v4_v5.npz: V4 nnz", v4.values().numel(), "V5 nnz", v5t.values().numel())


if __name__ == "__main__":
    which = _ARGS or ["lightgcn", "train", "freedom", "metrics", "real", "lattice", "cooccurrence", "scipy_adjacency"]
    if "scipy_adjacency" in which:
        scipy_adjacency_case()
    if "lattice" in which:
        lattice_case()
    if "cooccurrence" in which:
        cooccurrence_case()
    if "lightgcn" in which:
        lightgcn_case()
    if "train" in which:
        train_step_case()
    if "freedom" in which:
        freedom_case()
    if "metrics" in which:
        metrics_case()
    if "real" in which:
        real_graph_case()
    print("done")
