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

Everything here works with the gloo backend on CPU tensors as well (used by the world-size-2
tests); the compute calls themselves are GPU-only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
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


def gather_to_rank0(local: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """Concatenates 1-D/2-D per-rank tensors of uneven length along dim 0 on rank 0."""
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
