"""BPR training step around the propagation.

Reference: ``LightGCN.loss`` = ``bpr_loss`` + ``regularization_loss`` (Model/LightGCN.py:97-135),
driven by ``train_and_evaluate.train`` (train_and_evaluate.py:43-48).  The loss and its gradient
with respect to the propagated table come from one kernel (csrc/bpr.cu); the backward of the
propagation is the fused SpMM chain of propagate.py; the optimiser stays torch's.
"""
from __future__ import annotations

import torch

from . import _lib
from .propagate import graph_from_edge_index, propagate_mean


class _BprLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, users, pos_items, neg_items, reg_weight, item_offset):
        _lib.require_cuda(emb, "embeddings")
        lib = _lib.load()
        emb_c = emb.detach().contiguous()
        dev = emb_c.device
        users = users.to(device=dev, dtype=torch.int64).contiguous()
        pos_items = pos_items.to(device=dev, dtype=torch.int64).contiguous()
        neg_items = neg_items.to(device=dev, dtype=torch.int64).contiguous()
        B = int(users.numel())
        terms = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.zeros_like(emb_c)
        with torch.cuda.device(dev):
            rc = lib.chaorec_bpr_loss_grad_f32(_lib.ptr(emb_c), _lib.ptr(users), _lib.ptr(pos_items), _lib.ptr(neg_items), B,
                                               emb_c.shape[0], emb_c.shape[1], int(item_offset), float(reg_weight),
                                               _lib.ptr(terms), _lib.ptr(grad), _lib.stream_ptr(dev))
        _lib.check(rc, "chaorec_bpr_loss_grad_f32")
        ctx.save_for_backward(grad)
        return terms.sum()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


def bpr_loss(embeddings: torch.Tensor, users: torch.Tensor, pos_items: torch.Tensor, neg_items: torch.Tensor,
             reg_weight: float, num_user: int) -> torch.Tensor:
    """``bpr_loss + regularization_loss`` of Model/LightGCN.py:97-121 on the propagated table.
    ``pos_items`` / ``neg_items`` are LOCAL item ids (the reference subtracts num_user first,
    Model/LightGCN.py:124-125)."""
    return _BprLossFn.apply(embeddings, users, pos_items, neg_items, float(reg_weight), int(num_user))


def lightgcn_loss(user_weight: torch.Tensor, item_weight: torch.Tensor, edge_index: torch.Tensor, n_layers: int,
                  users: torch.Tensor, pos_items_global: torch.Tensor, neg_items_global: torch.Tensor, reg_weight: float):
    """Drop-in body of ``LightGCN.loss(users, pos_items, neg_items)`` (Model/LightGCN.py:123-135):
    returns (total_loss, propagated table); ``total_loss.backward()`` runs the fused backward."""
    num_user = user_weight.shape[0]
    x0 = torch.cat((user_weight, item_weight), dim=0)
    emb = propagate_mean(graph_from_edge_index(edge_index, x0.size(0)), x0, n_layers, True)
    loss = bpr_loss(emb, users, pos_items_global - num_user, neg_items_global - num_user, reg_weight, num_user)
    return loss, emb


def sample_negatives(users: torch.Tensor, train_items, num_item: int, seed: int, first_draw: int = 0,
                     item_offset: int = 0) -> torch.Tensor:
    """Negative items for ``users`` (int64, on the GPU), one each, drawn uniformly from the items the user
    has NOT interacted with -- the rejection loop of ``TrainingDataset.__getitem__`` (dataload.py:74-79) for
    a whole batch in one launch.  ``train_items`` is the ``MaskCSR`` of ``user_item_dict`` (the evaluation
    mask).  Returns ``item_offset + local id`` (pass ``num_user`` for the reference's global ids).  Draw i
    depends only on ``(seed, first_draw + i)``: use ``first_draw = epoch * len(dataset) + batch_start``
    to give every edge of every epoch its own draw.  (Same distribution as the reference; the reference's
    ``random.sample`` stream itself is host state and is not reproduced.)"""
    _lib.require_cuda(users, "users")
    lib = _lib.load()
    users = users.to(torch.int64).contiguous()
    out = torch.empty_like(users)
    with torch.cuda.device(users.device):
        rc = lib.chaorec_sample_negatives(_lib.ptr(users), int(users.numel()), _lib.ptr(train_items.rowptr),
                                          _lib.ptr(train_items.col), int(num_item), int(item_offset),
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_draw) & 0xFFFFFFFFFFFFFFFF,
                                          _lib.ptr(out), _lib.stream_ptr(users.device))
    _lib.check(rc, "chaorec_sample_negatives")
    return out


def training_batches(edge_users: torch.Tensor, edge_pos_items: torch.Tensor, train_items, num_user: int,
                     num_item: int, batch_size: int, epoch: int, seed: int = 0, drop_last: bool = False):
    """One epoch of ``DataLoader(TrainingDataset(...), batch_size, shuffle=True)`` (dataload.py:60-106,
    main.py's loader) without leaving the GPU: yields ``(users, pos_items, neg_items)`` int64 device tensors,
    items as GLOBAL ids like the reference's triples.  ``edge_users`` / ``edge_pos_items`` are the training
    pairs (global item ids) on the device.  The permutation comes from a device generator seeded with
    ``(seed, epoch)``; negatives from :func:`sample_negatives`.  ``drop_last=True`` skips a short final batch (what
    ``GraphedLightGCNStep``, whose captured kernels have the batch size baked in, needs; the reference's DataLoader
    keeps it)."""
    n = int(edge_users.numel())
    gen = torch.Generator(device=edge_users.device)
    gen.manual_seed((int(seed) * 1000003 + int(epoch)) & 0x7FFFFFFFFFFFFFFF)
    perm = torch.randperm(n, device=edge_users.device, generator=gen)
    users_all = edge_users[perm].to(torch.int64)
    pos_all = edge_pos_items[perm].to(torch.int64)
    neg_all = sample_negatives(users_all, train_items, num_item, seed, first_draw=int(epoch) * n, item_offset=num_user)
    for lo in range(0, n, batch_size):
        hi = min(lo + batch_size, n)
        if drop_last and hi - lo < batch_size:
            break
        yield users_all[lo:hi], pos_all[lo:hi], neg_all[lo:hi]


class GraphedLightGCNStep:
    """One LightGCN BPR training step (forward, loss, backward, optimiser) captured in a CUDA graph.

    At the reference's dataset sizes the step is launch-bound (a dozen sub-100-microsecond kernels), so
    replaying a captured graph removes the host overhead.  Indices are copied into static buffers
    before each replay; parameters are updated in place by the (capturable) optimiser.

        step = GraphedLightGCNStep(user_emb.weight, item_emb.weight, edge_index, n_layers, reg, batch, opt)
        loss = step(users, pos_items_global, neg_items_global)     # returns the loss tensor (device)
    """

    def __init__(self, user_weight, item_weight, edge_index, n_layers: int, reg_weight: float, batch: int,
                 optimizer: torch.optim.Optimizer, warmup: int = 3):
        dev = user_weight.device
        self.uw, self.iw, self.ei, self.L, self.reg, self.opt = user_weight, item_weight, edge_index, n_layers, reg_weight, optimizer
        self.users = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.pos = torch.full((batch,), user_weight.shape[0], dtype=torch.int64, device=dev)
        self.neg = torch.full((batch,), user_weight.shape[0], dtype=torch.int64, device=dev)
        graph_from_edge_index(edge_index, user_weight.shape[0] + item_weight.shape[0])   # build + cache the CSR outside capture
        # Warm-up runs real steps on a side stream (it creates .grad, the optimiser state and every cached
        # work buffer outside the capture); parameters are then restored and the optimiser state zeroed, so
        # the first replay is step 1 of a fresh optimiser.
        saved = [user_weight.detach().clone(), item_weight.detach().clone()]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
            with torch.no_grad():
                user_weight.copy_(saved[0])
                item_weight.copy_(saved[1])
                for st in optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()

    def _eager(self):
        self.opt.zero_grad(set_to_none=False)
        loss, _ = lightgcn_loss(self.uw, self.iw, self.ei, self.L, self.users, self.pos, self.neg, self.reg)
        loss.backward()
        self.opt.step()
        return loss

    def __call__(self, users, pos_items_global, neg_items_global):
        # the captured kernels have the batch size baked in (static index buffers, 1/B in the loss): a short last
        # batch of an epoch must go through the eager step or a second captured step of its own size
        if tuple(users.shape) != tuple(self.users.shape) or tuple(pos_items_global.shape) != tuple(self.pos.shape) \
                or tuple(neg_items_global.shape) != tuple(self.neg.shape):
            raise RuntimeError(f"GraphedLightGCNStep was captured for batches of {tuple(self.users.shape)}; got "
                               f"{tuple(users.shape)} (use training_batches(..., drop_last=True) or an eager step for the tail)")
        self.users.copy_(users, non_blocking=True)
        self.pos.copy_(pos_items_global, non_blocking=True)
        self.neg.copy_(neg_items_global, non_blocking=True)
        self.graph.replay()
        return self.loss
