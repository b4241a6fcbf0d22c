"""Single-node multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

How the three parts of the path shard:
  * evaluation  -- users are independent: contiguous user shards, the item table replicated, no
                   collective during compute; at the end the per-row hit masks (8 B/row) are
                   gathered to rank 0, which runs the sequential fp64 metric reduction, so the
                   metrics are bit-identical for every world size;
  * kNN build   -- item rows are independent: each rank owns rows [lo, hi) of the similarity
                   (``knn_topk(..., row_range=...)``), outputs are concatenated by an all-gather;
  * propagation -- one exchange per layer: rows of the adjacency are dealt round-robin to the
                   ranks (node i -> rank i % G, which balances Zipf hubs statistically) and
                   relabelled so that every rank owns one contiguous block of the permuted node
                   numbering; each layer is a local SpMM over the rank's rows followed by one
                   all-gather of the [N/G, d] output blocks into the replicated table.

For the bipartite user-item graph itself there is a cheaper exchange (``BipartiteShard``): users stay on
their rank for the whole propagation and only the ITEM table moves -- every rank adds up its users'
contributions to all item rows and one all-reduce of the [I, d] partials per layer replaces the all-gather of
the whole [U+I, d] table (at the 10M-user / 2M-item config: 0.5 GB instead of 2.7 GB per layer).

Everything here works with the gloo backend on CPU tensors as well (used by the world-size-2
tests); the compute calls themselves are GPU-only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, near-equal split of ``range(n)``; the first ``n % world`` shards get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_to_rank0(local: torch.Tensor, group=None, sizes=None) -> Optional[torch.Tensor]:
    """Concatenates 1-D/2-D per-rank tensors of uneven length along dim 0 on rank 0 (a real gather: only rank 0
    receives).  ``sizes``: the per-rank lengths when the caller knows them (``shard_range``), which saves the
    exchange of the lengths."""
    rank, world = world_info()
    if world == 1:
        return local
    if sizes is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        got = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(got, n, group=group)
        sizes = [int(s.item()) for s in got]
    cap = max(sizes)
    if local.shape[0] == cap:
        pad = local.contiguous()
    else:
        pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def allgather_concat(local: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks get the concatenation (uneven lengths allowed)."""
    rank, world = world_info()
    if world == 1:
        return local
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def knn_topk_sharded(feats: torch.Tensor, k: int, exclude_self: bool = False, group=None, topk_fn=None):
    """Item-row sharded kNN build: every rank ranks rows [lo, hi) of the similarity against ALL rows
    (``knn_topk(..., row_range=...)``), then one all-gather of the [N/G, k] values and indices gives every rank the
    whole [N, k] result -- what ``get_knn_adj_mat`` / ``build_knn_neighbourhood`` consume.  Bit-identical to the
    single-GPU call for any world size (rows are independent).  ``topk_fn(feats, k, exclude_self, row_range)``
    defaults to the CUDA kernel; the gloo tests pass a CPU stand-in."""
    rank, world = world_info()
    lo, hi = shard_range(feats.shape[0], world, rank)
    if topk_fn is None:
        from .knn import knn_topk as topk_fn
    val, idx = topk_fn(feats, k, exclude_self=exclude_self, row_range=(lo, hi))
    if world == 1:
        return val, idx
    return allgather_concat(val.contiguous(), group), allgather_concat(idx.contiguous(), group)


# ---------------------------------------------------------------------------------------------
# propagation: cyclic row deal + relabelling
# ---------------------------------------------------------------------------------------------


@dataclass
class CyclicPartition:
    """node i  <->  permuted id  (i % G) * B + i // G   with B = ceil(N / G) (padded blocks)."""
    n_nodes: int
    world: int

    @property
    def block(self) -> int:
        return (self.n_nodes + self.world - 1) // self.world

    @property
    def n_padded(self) -> int:
        return self.block * self.world

    def to_permuted(self, ids: torch.Tensor) -> torch.Tensor:
        return (ids % self.world) * self.block + ids // self.world

    def from_permuted(self, pid: torch.Tensor) -> torch.Tensor:
        return (pid % self.block) * self.world + pid // self.block

    def permute_rows(self, x: torch.Tensor) -> torch.Tensor:
        """[N, d] in node order -> [n_padded, d] in permuted order (padding rows are zero)."""
        out = torch.zeros((self.n_padded,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        out[self.to_permuted(torch.arange(self.n_nodes, device=x.device))] = x
        return out

    def unpermute_rows(self, xp: torch.Tensor) -> torch.Tensor:
        return xp[self.to_permuted(torch.arange(self.n_nodes, device=xp.device))]

    def local_edges(self, row: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], rank: int):
        """Edges whose destination row belongs to ``rank``: (local row, permuted col, val)."""
        mine = (row % self.world) == rank
        r = row[mine] // self.world
        c = self.to_permuted(col[mine])
        return r, c, (val[mine] if val is not None else None)


def allgather_blocks(local_block: torch.Tensor, out_full: torch.Tensor, group=None) -> torch.Tensor:
    """out_full[g*B:(g+1)*B] = rank g's block: the per-layer exchange of the propagation."""
    rank, world = world_info()
    if world == 1:
        out_full.copy_(local_block)
        return out_full
    dist.all_gather_into_tensor(out_full, local_block.contiguous(), group=group)
    return out_full


def propagate_mean_sharded(local_graph, part: CyclicPartition, x0_perm: torch.Tensor, n_layers: int,
                           alpha: Optional[float] = None, group=None) -> torch.Tensor:
    """Row-partitioned L-layer propagation + layer mean.

    ``local_graph``: CSRGraph of this rank's rows (M = part.block, N = part.n_padded, columns in
    permuted numbering, row/col scales already attached).  ``x0_perm``: the replicated layer-0
    table in permuted order.  Returns this rank's block of the layer mean, [part.block, d].
    """
    from . import propagate as P
    rank, world = world_info()
    B = part.block
    if alpha is None:
        alpha = 1.0 / (n_layers + 1)
    acc = x0_perm[rank * B:(rank + 1) * B] * alpha
    cur = x0_perm
    nxt = torch.empty_like(x0_perm) if n_layers > 1 else None
    y = torch.empty_like(acc)
    for l in range(n_layers):
        last = l == n_layers - 1
        P._spmm_raw(local_graph, cur, None if last else y, acc, alpha, 0)   # acc += alpha * y, fused
        if not last:
            allgather_blocks(y, nxt, group)
            cur, nxt = nxt, (torch.empty_like(x0_perm) if cur is x0_perm else cur)
    return acc


# ---------------------------------------------------------------------------------------------
# propagation on the bipartite user-item graph: users stay local, the item table is all-reduced
# ---------------------------------------------------------------------------------------------


@dataclass
class BipartiteShard:
    """This rank's share of the symmetric user-item graph: users [u_lo, u_hi) and ALL items.

    ``ui``: [U_local, I] block (rows = local users, columns = items), ``iu``: its transpose [I, U_local];
    the degree scales d^-1/2 are attached as row/col scales (item degrees are global)."""
    num_user: int
    num_item: int
    u_lo: int
    u_hi: int
    ui: object
    iu: object
    iu_blocks: Optional[list] = None     # [(item_lo, item_hi, graph)]: ``iu`` cut into row blocks, so that the
                                         # all-reduce of one block runs under the product of the next


def build_bipartite_shard(users: torch.Tensor, local_items: torch.Tensor, num_user: int, num_item: int,
                          group=None, n_blocks: Optional[int] = None) -> BipartiteShard:
    """Builds the shard from (user, local item) training pairs on the GPU.  ``users`` may hold all pairs (every
    rank then keeps the pairs of its own user range) or only this rank's."""
    from . import propagate as P
    import os
    rank, world = world_info()
    if n_blocks is None:
        # measured at the 100M-edge config: cutting the item rows into 4 blocks (all-reduce of block b under the
        # product of block b+1) gives 55 -> 64 G edges/s on 8 GPUs, but costs 12 % on 2 GPUs, where the exchange
        # is short and the smaller launches and NCCL's own CTAs only get in the way
        n_blocks = int(os.environ.get("CHAOREC_BIPARTITE_BLOCKS", "4" if world >= 4 else "1"))
    u_lo, u_hi = shard_range(num_user, world, rank)
    mine = (users >= u_lo) & (users < u_hi)
    u = (users[mine] - u_lo).to(torch.int64)
    it = local_items[mine].to(torch.int64)
    item_deg = torch.bincount(it, minlength=num_item)
    if world > 1:
        dist.all_reduce(item_deg, group=group)
    user_deg = torch.bincount(u, minlength=u_hi - u_lo)
    dis_u = user_deg.to(torch.float32).pow(-0.5)
    dis_i = item_deg.to(torch.float32).pow(-0.5)
    dis_u[torch.isinf(dis_u)] = 0
    dis_i[torch.isinf(dis_i)] = 0
    ui = P.csr_from_coo(u, it, None, u_hi - u_lo, num_item)
    ui.row_scale, ui.col_scale = dis_u, dis_i
    iu = P.csr_from_coo(it, u, None, num_item, u_hi - u_lo)
    iu.row_scale, iu.col_scale = dis_i, dis_u
    blocks = None
    if world > 1 and n_blocks > 1:
        blocks = []
        for b in range(n_blocks):
            lo, hi = shard_range(num_item, n_blocks, b)
            e_lo, e_hi = int(iu.rowptr[lo]), int(iu.rowptr[hi])
            g = P.CSRGraph((iu.rowptr[lo:hi + 1] - iu.rowptr[lo]).contiguous(), iu.col[e_lo:e_hi].contiguous(), None, hi - lo,
                           u_hi - u_lo).plan_hubs()
            g.row_scale, g.col_scale = dis_i[lo:hi].contiguous(), dis_u
            blocks.append((lo, hi, g))
    return BipartiteShard(num_user, num_item, u_lo, u_hi, ui, iu, blocks)


def _cuda_spmm(graph, x: torch.Tensor) -> torch.Tensor:
    from . import propagate as P
    y = torch.empty(graph.M, x.shape[1], dtype=torch.float32, device=x.device)
    P._spmm_raw(graph, x.contiguous(), y, None, 0.0, 0)
    return y


def propagate_mean_bipartite(shard: BipartiteShard, xu_local: torch.Tensor, xi: torch.Tensor, n_layers: int,
                             alpha: Optional[float] = None, spmm=None, group=None):
    """L-layer LightGCN propagation + layer mean on a ``BipartiteShard``.

    ``xu_local``: layer-0 rows of this rank's users [U_local, d]; ``xi``: the replicated layer-0 item table
    [I, d].  Per layer: user rows from the item table (local), partial item rows from the local users, one
    all-reduce (sum) of the item partials.  Returns (mean over layers of the local user rows, of all item rows).
    ``spmm(graph, x)`` defaults to the CUDA kernel; the gloo tests pass a CPU product."""
    rank, world = world_info()
    if alpha is None:
        alpha = 1.0 / (n_layers + 1)
    acc_u, acc_i = xu_local * alpha, xi * alpha
    xu, xi_cur = xu_local, xi
    for l in range(n_layers):
        last = l == n_layers - 1
        # item partials first, block by block: the all-reduce of a block runs under the products that follow
        works = []
        if spmm is None and shard.iu_blocks and world > 1:
            from . import propagate as P
            yi = torch.empty_like(acc_i)
            for lo, hi, g in shard.iu_blocks:
                P._spmm_raw(g, xu.contiguous(), yi[lo:hi], None, 0.0, 0)
                works.append(dist.all_reduce(yi[lo:hi], group=group, async_op=True))
        else:
            yi = (spmm or _cuda_spmm)(shard.iu, xu)
            if world > 1:
                works.append(dist.all_reduce(yi, group=group, async_op=True))
        if spmm is None:
            from . import propagate as P
            yu = None if last else torch.empty_like(acc_u)
            P._spmm_raw(shard.ui, xi_cur.contiguous(), yu, acc_u, alpha, 0)      # acc_u += alpha * yu, fused
        else:
            yu = spmm(shard.ui, xi_cur)
            acc_u.add_(yu, alpha=alpha)
        for w in works:
            w.wait()
        acc_i.add_(yi, alpha=alpha)
        xu, xi_cur = yu, yi
    return acc_u, acc_i


# ---------------------------------------------------------------------------------------------
# the same propagation with the item-table exchange FUSED into the kernels (no NCCL call per layer)
# ---------------------------------------------------------------------------------------------


def _mark(stream) -> "torch.cuda.Event":
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


class FusedItemExchange:
    """Symmetric buffers (every rank holds its peers' addresses: ``torch.distributed._symmetric_memory``, CUDA
    backend = cuMem IPC handles over NVLink) for the fused exchange of ``propagate_mean_bipartite_fused``:

      ``slots``   [world, block, d]  this rank's inbox: slot s receives rank s's partial rows of MY item block, stored
                                     by the EPILOGUE of rank s's item-row SpMM (``chaorec_spmm_csr_push_f32``); rows
                                     without a local edge are not sent at all (the owner re-zeroes its inbox as it
                                     reads it), which at 8 ranks is close to half of the (item, rank) pairs of a
                                     Zipf-shaped graph;
      ``tables``  2 x [world * block, d]  the replicated item table of the current / next layer: every owner sums its
                                     inbox in rank order and stores the rows into all ranks' tables
                                     (``chaorec_reduce_slots_broadcast_f32``).

    Per layer and rank: (world-1)/world of the item partials out and the same amount in, both issued by the compute
    kernels themselves; two device-side barriers (signal pads) order producers and consumers."""

    def __init__(self, num_item: int, d: int, device: torch.device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        rank, world = world_info()
        if world < 2:
            raise RuntimeError("FusedItemExchange needs an initialised process group with 2..8 ranks")
        if world > 8:
            raise RuntimeError("FusedItemExchange: at most 8 ranks (one NVSwitch domain)")
        self.rank, self.world, self.d, self.num_item = rank, world, d, num_item
        self._need = {}
        self.block = (num_item + world - 1) // world
        self.n_pad = self.block * world
        pg = group if group is not None else dist.group.WORLD
        name = pg.group_name
        if hasattr(symm, "is_symm_mem_enabled_for_group") and not symm.is_symm_mem_enabled_for_group(name):
            import warnings
            with warnings.catch_warnings():      # (a no-op on current torch, required by older releases)
                warnings.simplefilter("ignore")
                symm.enable_symm_mem_for_group(name)
        # two inboxes: layer l+1's partial rows may arrive while the owner still reduces layer l's
        self.slots = [symm.empty((world, self.block, d), dtype=torch.float32, device=device) for _ in range(2)]
        self.tables = [symm.empty((self.n_pad, d), dtype=torch.float32, device=device) for _ in range(2)]
        self.h_slots = [symm.rendezvous(t, name) for t in self.slots]
        self.h_tables = [symm.rendezvous(t, name) for t in self.tables]
        # (a symmetric tensor may sit at an offset inside its allocation block: the distance between the local tensor
        # and the local block base is the same on every rank)
        def mk(h, t):
            off = int(t.data_ptr()) - int(h.buffer_ptrs[rank])
            return (C.c_void_p * world)(*[int(p) + off for p in h.buffer_ptrs])
        self.slot_ptrs = [mk(h, t) for h, t in zip(self.h_slots, self.slots)]
        self.table_ptrs = [mk(h, t) for h, t in zip(self.h_tables, self.tables)]
        # Priorities matter: a one-CTA barrier kernel queued behind a 500k-CTA SpMM of an equal-priority stream waits for
        # that grid to drain (measured: 1.2 ms per layer on 4 GPUs), so the exchange chain outranks the products.
        self.side = torch.cuda.Stream(device=device, priority=0)       # user-row SpMM
        self.push = torch.cuda.Stream(device=device, priority=-1)      # item-row SpMM (push) + barrier 0
        self.reduce = torch.cuda.Stream(device=device, priority=-2)    # owner-side reduce + broadcast + barrier 1
        for t in self.slots + self.tables:
            t.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def need_map(self, shard: "BipartiteShard") -> torch.Tensor:
        """uint8 [world, num_item]: need[p][i] != 0 iff some user of rank p has an edge to item i (rank p gathers row i
        in its user-row product).  Built once per shard: one all-gather of 1 byte per item."""
        key = id(shard)
        hit = self._need.get(key)
        if hit is not None and hit[0] is shard:
            return hit[1]
        deg = shard.iu.rowptr[1:] - shard.iu.rowptr[:-1]
        mine = (deg > 0).to(torch.uint8).contiguous()
        out = torch.empty(self.world, mine.numel(), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(out, mine)
        self._need = {key: (shard, out)}
        return out

    def barrier(self, channel: int) -> None:
        """Device-side barrier of all ranks on the CURRENT stream (signal pads of the first inbox; one channel per
        stream that may be inside a barrier at the same time)."""
        self.h_slots[0].barrier(channel=channel)


def propagate_mean_bipartite_fused(shard: BipartiteShard, ex: FusedItemExchange, xu_local: torch.Tensor, xi: torch.Tensor,
                                   n_layers: int, alpha: Optional[float] = None, timing: Optional[list] = None):
    """``propagate_mean_bipartite`` without a collective call: the item partials leave through the epilogue of the
    item-row SpMM into their owners' inboxes, the owners reduce (rank order) and store the rows into every rank's
    item table.  Three streams: the caller's stream runs the item-row products (push), ``ex.reduce`` the owner-side
    reduce + broadcast, ``ex.side`` the user-row products -- layer l+1's push only needs layer l's user rows, so it
    runs beside layer l's reduce + broadcast (two inboxes).  Same results as the NCCL variant up to the order of the
    fp32 partial sums (deterministic for a given world size)."""
    from . import _lib
    from . import propagate as P
    lib = _lib.load()
    rank, world = ex.rank, ex.world
    dev = xu_local.device
    d = xu_local.shape[1]
    I = shard.num_item
    B = ex.block
    lo_i = min(rank * B, I)
    n_rows = max(min((rank + 1) * B, I) - lo_i, 0)
    if alpha is None:
        alpha = 1.0 / (n_layers + 1)
    import os
    need = ex.need_map(shard) if os.environ.get("CHAOREC_FUSED_SPARSE_BCAST", "1") != "0" else None
    caller = torch.cuda.current_stream(dev)
    acc_u = xu_local * alpha
    acc_blk = torch.zeros(B, d, dtype=torch.float32, device=dev)
    acc_blk[:n_rows] = xi[lo_i:lo_i + n_rows] * alpha
    xu, xi_cur = xu_local.contiguous(), xi.contiguous()
    start = _mark(caller)
    main = ex.push
    for st in (ex.side, ex.push, ex.reduce):
        st.wait_event(start)
    table_ready = start            # event after which xi_cur is complete on this rank
    reduced = [None, None]         # per inbox: event after which the owner has consumed (and re-zeroed) it
    keep = []
    for l in range(n_layers):
        last = l == n_layers - 1
        k = l % 2
        # side stream: user rows of this layer from the (complete) item table of the previous one, fused layer mean
        ex.side.wait_event(table_ready)
        with torch.cuda.stream(ex.side):
            yu = None if last else torch.empty_like(acc_u)
            P._spmm_raw(shard.ui, xi_cur, yu, acc_u, alpha, 0)
            done_u = _mark(ex.side)
            if yu is not None:
                keep.append(yu)
        # push stream: item partials of this layer from the local user rows, stored into the owners' inboxes k
        with torch.cuda.stream(main):
            marks = [_mark(main)] if timing is not None else None
            if reduced[k] is not None:
                main.wait_event(reduced[k])             # (all ranks passed barrier 1 of layer l-2: inbox k is free everywhere)
            P._spmm_push(shard.iu, xu, ex.slot_ptrs[k], world, B, rank)
            if marks is not None:
                marks.append(_mark(main))
            ex.barrier(0)                               # every rank's partial rows of this layer have landed
            pushed = _mark(main)
            if marks is not None:
                marks.append(pushed)
        # reduce stream: sum my inbox in rank order, store the rows into every rank's table k
        ex.reduce.wait_event(pushed)
        with torch.cuda.stream(ex.reduce):
            with torch.cuda.device(dev):
                rc = lib.chaorec_reduce_slots_broadcast_f32(_lib.ptr(ex.slots[k]), world, B, n_rows, d, lo_i, ex.table_ptrs[k],
                                                            _lib.ptr(acc_blk), float(alpha), 1 if last else 0,
                                                            _lib.ptr(need), int(need.shape[1]) if need is not None else 0, rank,
                                                            C_stream(ex.reduce))
            _lib.check(rc, "chaorec_reduce_slots_broadcast_f32")
            if marks is not None:
                marks.append(_mark(ex.reduce))
            ex.barrier(1)                               # every owner's rows have landed in every table
            table_ready = _mark(ex.reduce)
            reduced[k] = table_ready
            if marks is not None:
                marks.append(table_ready)
        main.wait_event(done_u)                         # the next push reads this layer's user rows
        if marks is not None:
            marks.append(done_u)
            timing.append(marks)
        xu, xi_cur = yu, ex.tables[k][:I]
    caller.wait_event(table_ready)
    caller.wait_event(done_u)
    for t in [acc_u, acc_blk, xi_cur] + keep:
        for st in (ex.side, ex.push, ex.reduce, caller):
            t.record_stream(st)
    return acc_u, xi_cur


def C_stream(stream: "torch.cuda.Stream"):
    import ctypes as C
    return C.c_void_p(stream.cuda_stream)
