"""CPU restatement of ChaoRec's hot path (propagation, full-rank evaluation, kNN graph).

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` may import this module.  Nothing under ``chaorec_b200/`` imports it; the
product path fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against the *reference itself*: ``tests/golden/make_golden.py``
imports the unmodified ``Model/FREEDOM.py``, ``Model/LightGCN.py`` (through
``oracle/pyg_stub``), ``utils.gene_metrics`` and ``gen_hypergraph_u_i.get_knn_adj_mat`` from
``/root/reference`` in the authoring container, runs them on seeded inputs and stores the
outputs under ``tests/golden``; ``tests/test_oracle_golden.py`` checks every function below
against those files.

Every function cites the reference lines it follows (paths relative to /root/reference).
The arithmetic is kept in the same torch CPU operators the reference calls, so the
restatement is the reference's CPU path minus the nn.Module plumbing.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------
# (1) propagation
# ----------------------------------------------------------------------------------------


def symmetrise_edges(train_edges: np.ndarray | torch.Tensor) -> torch.Tensor:
    """[E,2] (user, global item id) -> [2,2E]: first E columns u->i, last E columns i->u.

    Model/LightGCN.py:63-64 and Model/FREEDOM.py:38-39.
    """
    e = torch.as_tensor(np.asarray(train_edges)).t().contiguous()
    return torch.cat((e, e[[1, 0]]), dim=1)


def pyg_lightgcn_conv(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """One LightGCNConv layer, Model/LightGCN.py:28-46 (same as Model/LATTICE.py:23-42).

    deg = degree(row); norm = deg^-1/2[row] * deg^-1/2[col]; out[col] += norm * x[row].
    """
    ei = edge_index.long()
    src, dst = ei[0], ei[1]
    deg = torch.zeros(x.size(0), dtype=x.dtype).scatter_add_(0, src, torch.ones(src.numel(), dtype=x.dtype))
    dis = deg.pow(-0.5)
    norm = dis[src] * dis[dst]
    msg = norm.view(-1, 1) * x[src]
    return torch.zeros_like(x).index_add_(0, dst, msg)


def lightgcn_forward(user_w: torch.Tensor, item_w: torch.Tensor, edge_index: torch.Tensor,
                     n_layers: int) -> torch.Tensor:
    """Model/LightGCN.py:76-95: L convs, then sum_l (1/(L+1)) * e_l accumulated from zeros."""
    x = torch.cat((user_w, item_w), dim=0)
    layers = [x]
    for _ in range(n_layers):
        x = pyg_lightgcn_conv(x, edge_index)
        layers.append(x)
    w = 1.0 / len(layers)
    out = torch.zeros_like(layers[0])
    for e in layers:
        out += w * e
    return out


def lightgcn_loss(user_w, item_w, edge_index, n_layers, users, pos_items_global, neg_items_global, reg_weight):
    """LightGCN.loss, Model/LightGCN.py:97-135: BPR term -mean(log(sigmoid(pos-neg)+1e-5)) plus
    reg * (mean(u^2) + mean(p^2) + mean(n^2)) on rows of the propagated table (differentiable)."""
    num_user = user_w.shape[0]
    pos = pos_items_global - num_user
    neg = neg_items_global - num_user
    emb = lightgcn_forward(user_w, item_w, edge_index, n_layers)
    u, p, n = emb[users], emb[num_user + pos], emb[num_user + neg]
    pos_s = torch.sum(u * p, dim=1)
    neg_s = torch.sum(u * n, dim=1)
    bpr = -torch.mean(torch.log(torch.sigmoid(pos_s - neg_s) + 1e-5))
    reg = reg_weight * (torch.mean(u ** 2) + torch.mean(p ** 2) + torch.mean(n ** 2))
    return bpr + reg


def freedom_norm_adj(edge_index: torch.Tensor, n_nodes: int) -> torch.Tensor:
    """V2 values, Model/FREEDOM.py:73-83: degree = bincount(cat(row, col)) (twice the true
    degree on the symmetrised list) -> values are half the textbook D^-1/2 A D^-1/2."""
    ei = edge_index.long()
    r, c = ei[0], ei[1]
    deg = torch.bincount(torch.cat([r, c]))
    dis = deg.pow(-0.5)
    return torch.sparse_coo_tensor(ei, dis[r] * dis[c], torch.Size([n_nodes, n_nodes]))


def normalize_adj_m(indices: torch.Tensor, adj_size: Sequence[int]) -> torch.Tensor:
    """V3 values, Model/FREEDOM.py:85-100: (1e-7+rowsum)^-1/2[r] * (1e-7+colsum)^-1/2[c]
    over the bipartite R block (int64 ones)."""
    adj = torch.sparse_coo_tensor(indices, torch.ones_like(indices[0]), torch.Size(adj_size))
    rs = 1e-7 + torch.sparse.sum(adj, -1).to_dense()
    cs = 1e-7 + torch.sparse.sum(adj.t(), -1).to_dense()
    return torch.pow(rs, -0.5)[indices[0]] * torch.pow(cs, -0.5)[indices[1]]


def freedom_masked_adj(keep_indices: torch.Tensor, num_user: int, num_item: int) -> torch.Tensor:
    """Model/FREEDOM.py:153-162 for an already-sampled set of kept (user, local item) edges."""
    ki = keep_indices.clone()
    vals = normalize_adj_m(ki, (num_user, num_item))
    ki[1] += num_user
    n = num_user + num_item
    return torch.sparse_coo_tensor(torch.cat((ki, torch.flip(ki, [0])), 1), torch.cat((vals, vals)),
                                   torch.Size([n, n]))


def sparse_mm_chain_mean(adj: torch.Tensor, x0: torch.Tensor, n_layers: int) -> torch.Tensor:
    """Model/FREEDOM.py:171-178: ego chain through torch.sparse.mm, stack -> mean(dim=1)."""
    ego = x0
    stack = [ego]
    for _ in range(n_layers):
        ego = torch.sparse.mm(adj, ego)
        stack.append(ego)
    return torch.stack(stack, dim=1).mean(dim=1, keepdim=False)


def freedom_forward(user_w, item_w, adj, mm_adj, n_layers: int, mm_layers: int) -> torch.Tensor:
    """Model/FREEDOM.py:164-183 -> the cached ``self.result`` [U+I, d]."""
    h = item_w
    for _ in range(mm_layers):
        h = torch.sparse.mm(mm_adj, h)
    allm = sparse_mm_chain_mean(adj, torch.cat((user_w, item_w), dim=0), n_layers)
    u_g, i_g = torch.split(allm, [user_w.size(0), item_w.size(0)], dim=0)
    return torch.cat((u_g, i_g + h), dim=0)


# ----------------------------------------------------------------------------------------
# (2) full-ranking evaluation
# ----------------------------------------------------------------------------------------


def masked_score_matrix(user_t: torch.Tensor, item_t: torch.Tensor, user_item_dict: Dict[int, Iterable[int]],
                        num_user: int, mask_value: float = 1e-6, row_lo: int = 0) -> torch.Tensor:
    """Model/LightGCN.py:147-152: S = U I^T (CPU fp32), then S[u][train items - U] = 1e-6."""
    s = torch.matmul(user_t, item_t.t())
    hi = row_lo + s.size(0)
    for u, items in user_item_dict.items():
        if row_lo <= u < hi:
            cols = torch.LongTensor(list(items)) - num_user
            s[u - row_lo][cols] = mask_value
    return s


def gene_ranklist(result: torch.Tensor, num_user: int, num_item: int, user_item_dict, topk: int = 50,
                  mask_value: float = 1e-6, stable: bool = False) -> torch.Tensor:
    """Model/LightGCN.py:137-162.  ``stable=False`` is the reference verbatim (torch.topk,
    tie order unspecified); ``stable=True`` applies the stated tie-break (descending score,
    ties -> ascending item index) through a stable descending sort."""
    u = result[:num_user].cpu()
    i = result[num_user:num_user + num_item].cpu()
    s = masked_score_matrix(u, i, user_item_dict, num_user, mask_value)
    if stable:
        idx = torch.sort(s, dim=1, descending=True, stable=True)[1][:, :topk]
    else:
        idx = torch.topk(s, topk)[1]
    return idx + num_user


def gene_ranklist_batched(result, num_user, num_item, user_item_dict, step: int = 200, topk: int = 50,
                          mask_value: float = 1e-5, stable: bool = False) -> torch.Tensor:
    """Model/MMGCN.py:204-244 user-batched form (the only form that can run at 1M users)."""
    u = result[:num_user].cpu()
    i = result[num_user:num_user + num_item].cpu()
    by_user = user_item_dict
    out = []
    lo = 0
    while lo < num_user:
        hi = min(lo + step, num_user)
        s = torch.matmul(u[lo:hi], i.t())
        for r in range(lo, hi):
            items = by_user.get(r)
            if items is not None:
                s[r - lo][torch.LongTensor(list(items)) - num_user] = mask_value
        if stable:
            out.append(torch.sort(s, dim=1, descending=True, stable=True)[1][:, :topk] + num_user)
        else:
            out.append(torch.topk(s, topk)[1] + num_user)
        lo = hi
    return torch.cat(out, dim=0)


def user_metrics_at_k(ranked: List[int], positives: List[int], k: int, inv_log) -> Tuple[float, float, float, int, float]:
    """(precision, recall, ndcg, hit_rate, map)@k for one user, metrics.py:13-57.

    ``inv_log[i] = 1.0 / np.log(i + 2)`` (natural log, metrics.py:31,34).
    """
    top = ranked[:k]
    pos_set = set(positives)
    inter = len(set(top) & pos_set)
    precision = inter / k                                            # metrics.py:16
    recall = 0 if len(positives) == 0 else inter / len(positives)    # metrics.py:20-23
    hit = int(bool(inter))                                           # metrics.py:44
    if not positives:                                                # metrics.py:28, :49
        return precision, recall, 0, hit, 0
    idcg = sum([inv_log[i] for i in range(min(len(positives), k))])  # metrics.py:31
    dcg = 0
    ap = 0
    n_hit = 0
    for pos, item in enumerate(top):
        if item in positives:
            dcg += inv_log[pos]                                      # metrics.py:32-34
            n_hit += 1
            ap += n_hit / (pos + 1)                                  # metrics.py:53-56
    return precision, recall, dcg / idcg, hit, ap / len(positives)


def gene_metrics(val_data, rank_list, k_list) -> Dict[int, Dict[str, float]]:
    """utils.py:112-139: sequential python-float sums in ``val_data`` order, / len(val_data)."""
    inv_log = [1.0 / np.log(i + 2) for i in range(max(k_list) if len(k_list) else 0)]
    names = ('precision', 'recall', 'ndcg', 'hit_rate', 'map')
    acc = {k: {n: 0 for n in names} for k in k_list}
    for row in val_data:
        user, positives = row[0], row[1:]
        ranked = rank_list[user].tolist()
        for k in k_list:
            vals = user_metrics_at_k(ranked, positives, int(k), inv_log)
            for n, v in zip(names, vals):
                acc[k][n] += v
    n_users = len(val_data)
    for k in k_list:
        for n in names:
            acc[k][n] /= n_users
    return acc


# ----------------------------------------------------------------------------------------
# (3) modality item-item kNN graph
# ----------------------------------------------------------------------------------------


def build_sim(context: torch.Tensor) -> torch.Tensor:
    """Model/LATTICE.py:68-75 / Model/FREEDOM.py:114-116: rows / ||row||_2, then X X^T."""
    cn = context.div(torch.norm(context, p=2, dim=-1, keepdim=True))
    return torch.mm(cn, cn.transpose(1, 0))


def knn_laplacian_ones(indices: torch.Tensor, adj_size) -> torch.Tensor:
    """Model/FREEDOM.py:130-140: ones -> (1e-7+rowsum)^-1/2[r] * (1e-7+rowsum)^-1/2[c]."""
    adj = torch.sparse_coo_tensor(indices, torch.ones_like(indices[0]), adj_size)
    rs = 1e-7 + torch.sparse.sum(adj, -1).to_dense()
    ris = torch.pow(rs, -0.5)
    return torch.sparse_coo_tensor(indices, ris[indices[0]] * ris[indices[1]], adj_size)


def knn_conv_a(feats: torch.Tensor, k: int, stable: bool = False):
    """Model/FREEDOM.py:111-128 (self included): -> (indices [2,N*k], sparse adj)."""
    sim = build_sim(feats)
    if stable:
        ind = torch.sort(sim, dim=-1, descending=True, stable=True)[1][:, :k]
    else:
        ind = torch.topk(sim, k, dim=-1)[1]
    rows = torch.arange(ind.shape[0]).unsqueeze(1).expand(-1, k)
    indices = torch.stack((rows.flatten(), ind.flatten()), 0)
    return indices, knn_laplacian_ones(indices, sim.size())


def knn_conv_b(feats: torch.Tensor, k: int):
    """Model/LATTICE.py:46-52: top-k values kept -> (knn_val, knn_ind, dense scatter)."""
    sim = build_sim(feats)
    val, ind = torch.topk(sim, k, dim=-1)
    return val, ind, torch.zeros_like(sim).scatter_(-1, ind, val)


def dense_sym_laplacian(adj: torch.Tensor) -> torch.Tensor:
    """Model/LATTICE.py:56-64 as written (diagflat + two dense GEMMs)."""
    rs = torch.sum(adj, -1)
    dis = torch.pow(rs, -0.5)
    dis[torch.isinf(dis)] = 0.
    dm = torch.diagflat(dis)
    return torch.mm(torch.mm(dm, adj), dm)


def knn_conv_c(feats: torch.Tensor, k: int) -> torch.Tensor:
    """gen_hypergraph_u_i.py:41-48 (self excluded by -inf on the diagonal) -> knn_ind."""
    sim = build_sim(feats)
    sim.fill_diagonal_(-float('inf'))
    return torch.topk(sim, k, dim=-1)[1]


# ----------------------------------------------------------------------------------------
# Seeded synthetic inputs.  Shared by tests and bench so that the CUDA path
# and the oracle see identical bytes.
# ----------------------------------------------------------------------------------------


def zipf_item_sampler(num_item: int, gen: torch.Generator):
    """Zipf(alpha=1) over a random permutation of the items -> (cdf, perm)."""
    w = 1.0 / torch.arange(1, num_item + 1, dtype=torch.float64)
    cdf = torch.cumsum(w / w.sum(), 0)
    perm = torch.randperm(num_item, generator=gen)
    return cdf, perm


def synth_interactions(num_user: int, num_item: int, n_edges: int, seed: int, min_deg: int = 3):
    """Unique (user, local item) pairs: users uniform with >= min_deg each, items Zipf.

    Returns int64 tensors (user, item) sorted by (user, item), deduplicated.  The edge count
    after dedup is a little below ``n_edges``.
    """
    g = torch.Generator().manual_seed(seed)
    cdf, perm = zipf_item_sampler(num_item, g)
    base = torch.arange(num_user, dtype=torch.int64).repeat_interleave(min_deg)
    extra_n = max(n_edges - base.numel(), 0)
    extra = torch.randint(0, num_user, (extra_n,), generator=g, dtype=torch.int64)
    users = torch.cat((base, extra))
    r = torch.rand(users.numel(), generator=g, dtype=torch.float64)
    items = perm[torch.searchsorted(cdf, r).clamp_(max=num_item - 1)]
    key = torch.unique(users * num_item + items)
    return key // num_item, key % num_item


def synth_holdout(num_user: int, num_item: int, train_u: torch.Tensor, train_i: torch.Tensor, n_pos: int,
                  seed: int):
    """Per user ``n_pos`` held-out items not in train (val/test positives) -> [U, n_pos] local ids
    (a clash with a train item is re-drawn a few times, then moved to the next free id)."""
    g = torch.Generator().manual_seed(seed)
    train_keys = (train_u * num_item + train_i).numpy()
    train_set = np.sort(train_keys)
    out = torch.randint(0, num_item, (num_user, n_pos), generator=g, dtype=torch.int64).numpy()
    uu = np.arange(num_user, dtype=np.int64)[:, None]
    for _ in range(64):
        keys = uu * num_item + out
        pos = np.searchsorted(train_set, keys)
        pos[pos == train_set.size] = 0
        clash = train_set[pos] == keys
        for j in range(1, n_pos):  # no duplicate positives inside a row
            for jj in range(j):
                clash[:, j] |= out[:, j] == out[:, jj]
        if not clash.any():
            break
        out[clash] = (out[clash] + 1) % num_item
    return torch.from_numpy(out)


def to_reference_formats(num_user, train_u, train_i, heldout):
    """-> (train [E,2] int32 global ids, user_item_dict {u: [global ids]}, val_data object array of
    [user, pos...] lists) in the layouts dataload.data_load returns (dataload.py:21-58)."""
    train = torch.stack((train_u, train_i + num_user), 1).to(torch.int32).numpy()
    d: Dict[int, List[int]] = {}
    for u, i in train.tolist():
        d.setdefault(u, []).append(i)
    rows = np.empty(heldout.shape[0], dtype=object)
    held = (heldout + num_user).tolist()
    for u in range(heldout.shape[0]):
        rows[u] = [u] + held[u]
    return train, d, rows


def xavier_embeddings(n: int, d: int, seed: int) -> torch.Tensor:
    """nn.init.xavier_uniform_ on an [n,d] table (Model/LightGCN.py:69-70)."""
    g = torch.Generator().manual_seed(seed)
    bound = math.sqrt(6.0 / (n + d))
    return (torch.rand(n, d, generator=g) * 2 - 1) * bound


def normal_embeddings(n: int, d: int, seed: int, std: float = 0.1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * std


# ---------------------------------------------------------------------------------------------
# Negative sampling (dataload.py:74-79) with the counter-based generator of csrc/bpr.cu
# ---------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def sample_negatives(users, user_item_sets, num_item: int, seed: int, first_draw: int = 0, item_offset: int = 0):
    """Restatement of ``TrainingDataset.__getitem__``'s rejection loop (dataload.py:74-79: draw uniformly
    from all items until the draw is not in ``user_item_dict[user]``) with the stateless generator the CUDA
    kernel uses, so both sides can be compared bit for bit.  ``user_item_sets[u]`` = set of LOCAL train item
    ids of user u.  Pure Python: small cases only."""
    out = []
    for i, u in enumerate(users):
        base = _splitmix64((seed & _M64) ^ _splitmix64((first_draw + i) & _M64))
        c = 0
        for attempt in range(256):
            r = _splitmix64((base + attempt) & _M64)
            c = (r * num_item) >> 64
            if c not in user_item_sets[int(u)]:
                break
        out.append(item_offset + c)
    return out


# ---------------------------------------------------------------------------------------------
# LATTICE learned item graph (Model/LATTICE.py:45-74,137-177), dense restatement
# ---------------------------------------------------------------------------------------------
def lattice_build_sim(context: torch.Tensor) -> torch.Tensor:
    """Model/LATTICE.py:67-74."""
    cn = context.div(torch.norm(context, p=2, dim=-1, keepdim=True))
    return torch.mm(cn, cn.transpose(1, 0))


def lattice_knn_neighbourhood(adj: torch.Tensor, topk: int) -> torch.Tensor:
    """Model/LATTICE.py:45-52."""
    knn_val, knn_ind = torch.topk(adj, topk, dim=-1)
    return torch.zeros_like(adj).scatter_(-1, knn_ind, knn_val)


def lattice_normalized_laplacian(adj: torch.Tensor) -> torch.Tensor:
    """Model/LATTICE.py:56-64."""
    d_inv_sqrt = torch.pow(torch.sum(adj, -1), -0.5)
    d_inv_sqrt = torch.where(torch.isinf(d_inv_sqrt), torch.zeros_like(d_inv_sqrt), d_inv_sqrt)
    return d_inv_sqrt.unsqueeze(1) * adj * d_inv_sqrt.unsqueeze(0)          # == diagflat(d) @ adj @ diagflat(d)


def lattice_forward(p: dict, edge_index: torch.Tensor, num_user: int, num_item: int, topk: int, lambda_coeff: float,
                    mm_layers: int, n_layers: int) -> torch.Tensor:
    """``LATTICE.forward(build_item_graph=True)`` (Model/LATTICE.py:137-177) on plain tensors: ``p`` holds
    user, item, image, text, image_w, image_b, text_w, text_b, modal.  Dense [N, N] like the reference."""
    image_feats = p["image"] @ p["image_w"].t() + p["image_b"]
    text_feats = p["text"] @ p["text_w"].t() + p["text_b"]
    weight = torch.softmax(p["modal"], dim=0)
    image_adj = lattice_knn_neighbourhood(lattice_build_sim(image_feats), topk)
    text_adj = lattice_knn_neighbourhood(lattice_build_sim(text_feats), topk)
    learned = lattice_normalized_laplacian(weight[0] * image_adj + weight[1] * text_adj)
    io = lattice_normalized_laplacian(lattice_knn_neighbourhood(lattice_build_sim(p["image"].detach()), topk))
    to = lattice_normalized_laplacian(lattice_knn_neighbourhood(lattice_build_sim(p["text"].detach()), topk))
    item_adj = (1 - lambda_coeff) * learned + lambda_coeff * (weight[0] * io + weight[1] * to)
    h = p["item"]
    for _ in range(mm_layers):
        h = torch.mm(item_adj, h)
    ego = torch.cat((p["user"], p["item"]), dim=0)
    embs = [ego]
    for _ in range(n_layers):
        ego = pyg_lightgcn_conv(ego, edge_index)                 # LATTICEGCNConv == LightGCNConv (Model/LATTICE.py:18-42)
        embs.append(ego)
    mean = torch.stack(embs, dim=1).mean(dim=1)
    u_g, i_g = torch.split(mean, [num_user, num_item], dim=0)
    return torch.cat((u_g, i_g + torch.nn.functional.normalize(h, p=2, dim=1)), dim=0)


# ---------------------------------------------------------------------------------------------
# User-user co-occurrence graph (dualgnn-gen-u-u-matrix.py:20-50,85-103)
# ---------------------------------------------------------------------------------------------
def user_cooccurrence_matrix(train_edges: np.ndarray, num_user: int) -> np.ndarray:
    """``gen_user_matrix`` (dualgnn-gen-u-u-matrix.py:20-50): C[u][v] = |items(u) & items(v)| for u != v, 0 on the
    diagonal, as R @ R.T of the deduplicated binary interaction matrix."""
    e = np.unique(np.asarray(train_edges, dtype=np.int64), axis=0)
    items, inv = np.unique(e[:, 1], return_inverse=True)
    R = np.zeros((num_user, items.size), dtype=np.int64)
    R[e[:, 0], inv] = 1
    C = R @ R.T
    np.fill_diagonal(C, 0)
    return C


def user_graph_topk(C: np.ndarray, k: int = 200):
    """dualgnn-gen-u-u-matrix.py:85-103 under the canonical order (count descending, ties -> smaller user id;
    ``torch.topk`` leaves the order among equal counts unspecified): {u: [indices, float values]}."""
    out = {}
    for u in range(C.shape[0]):
        row = C[u]
        nz = int((row > 0).sum())
        kk = min(nz, k)
        order = np.lexsort((np.arange(row.size), -row))[:kk]
        out[u] = [order.tolist(), [float(row[v]) for v in order]]
    return out


# ---------------------------------------------------------------------------------------------
# scipy adjacency builders V4 / V5.  The reference's own code cannot run here
# (dok_matrix._update was removed from scipy), so these restate the published formulas: UNPINNED.
# ---------------------------------------------------------------------------------------------
def scipy_norm_adj_v4(train_u: np.ndarray, train_i_local: np.ndarray, num_user: int, num_item: int) -> torch.Tensor:
    """Model/SimGCL.py:64-104: A = [[0,R],[R^T,0]] with ones; diag = ((A>0).sum(1) + 1e-7)^-0.5 in float64;
    L = D @ A @ D in float64, cast to float32 -> torch sparse COO."""
    import scipy.sparse as sp
    n = num_user + num_item
    r = np.concatenate((train_u, train_i_local + num_user))
    c = np.concatenate((train_i_local + num_user, train_u))
    A = sp.coo_matrix((np.ones(r.size, dtype=np.float32), (r, c)), shape=(n, n)).tocsr()
    A.data[:] = 1.0                                            # dok semantics: duplicates collapse to one entry
    diag = np.power(np.asarray((A > 0).sum(axis=1)).flatten() + 1e-7, -0.5)
    L = sp.coo_matrix(sp.diags(diag) @ A @ sp.diags(diag))
    idx = torch.tensor(np.array([L.row, L.col]), dtype=torch.long)
    return torch.sparse_coo_tensor(idx, torch.FloatTensor(L.data), (n, n), dtype=torch.float32)


def scipy_norm_adj_v5(train_u: np.ndarray, train_i_local: np.ndarray, num_user: int, num_item: int) -> torch.Tensor:
    """Model/MGCN.py:164-196: float32 adjacency, rowsum == 0 -> 1e-16, d = rowsum^-0.5 (inf -> 0), D A D."""
    import scipy.sparse as sp
    n = num_user + num_item
    r = np.concatenate((train_u, train_i_local + num_user))
    c = np.concatenate((train_i_local + num_user, train_u))
    A = sp.coo_matrix((np.ones(r.size, dtype=np.float32), (r, c)), shape=(n, n)).tocsr()
    A.data[:] = 1.0
    rowsum = np.array(A.sum(1))
    rowsum[rowsum == 0.] = 1e-16
    d_inv = np.power(rowsum, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.
    L = sp.diags(d_inv).dot(A).dot(sp.diags(d_inv)).tocoo()
    idx = torch.tensor(np.array([L.row, L.col]), dtype=torch.long)
    return torch.sparse_coo_tensor(idx, torch.tensor(L.data, dtype=torch.float32), (n, n))
