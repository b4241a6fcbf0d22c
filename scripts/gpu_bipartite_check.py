"""torchrun --nproc-per-node N scripts/gpu_bipartite_check.py : both sharded propagation variants on CUDA + NCCL --
BipartiteShard (users local, blocked all-reduce of the item partials) and the row-cyclic partition (one all-gather of
the table per layer) -- against the single-GPU fused propagation of the whole graph computed redundantly on every rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chaorec_b200 as cb  # noqa: E402
from chaorec_b200 import dist as cdist  # noqa: E402
from oracle import ref_ops  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
U, I, d, L = 20000, 5000, 64, 3
tu, ti = ref_ops.synth_interactions(U, I, 300000, 55)
tu, ti = tu.to(dev), ti.to(dev)
x0 = ref_ops.normal_embeddings(U + I, d, 56).to(dev)
full = cb.bipartite_graph(tu, ti, U, I, "v1")
want = cb.propagate_mean(full, x0, L)
shard = cdist.build_bipartite_shard(tu, ti, U, I)
au, ai = cdist.propagate_mean_bipartite(shard, x0[shard.u_lo:shard.u_hi].contiguous(), x0[U:].contiguous(), L)
eu = float((au - want[shard.u_lo:shard.u_hi]).abs().max() / want.abs().max())
ei = float((ai - want[U:]).abs().max() / want.abs().max())
# row-cyclic partition of the same adjacency (the north-star's all-gather design), CUDA SpMM + NCCL all-gather
n = U + I
part = cdist.CyclicPartition(n, world)
deg = torch.bincount(torch.cat((tu, ti + U)), minlength=n)
dis = deg.to(torch.float32).pow(-0.5)
dis[torch.isinf(dis)] = 0
rows, cols = [], []
for r_, c_ in ((tu, ti + U), (ti + U, tu)):
    lr, pc, _ = part.local_edges(r_, c_, None, rank)
    rows.append(lr); cols.append(pc)
g = cb.csr_from_coo(torch.cat(rows), torch.cat(cols), None, part.block, part.n_padded)
dis_perm = part.permute_rows(dis.unsqueeze(1)).squeeze(1).contiguous()
g.row_scale = dis_perm[rank * part.block:(rank + 1) * part.block].contiguous()
g.col_scale = dis_perm
blk = cdist.propagate_mean_sharded(g, part, part.permute_rows(x0), L)
want_perm = part.permute_rows(want)
ec = float((blk - want_perm[rank * part.block:(rank + 1) * part.block]).abs().max() / want.abs().max())
# fused exchange (SpMM epilogue stores into the owners' inboxes, owners reduce and store into every table): no NCCL call
ex = cdist.FusedItemExchange(I, d, dev)
ef = 0.0
for rep in range(3):                       # buffers are reused across calls: repeat
    fu, fi = cdist.propagate_mean_bipartite_fused(shard, ex, x0[shard.u_lo:shard.u_hi].contiguous(), x0[U:].contiguous(), L)
    torch.cuda.synchronize()
    ef = max(ef, float((fu - want[shard.u_lo:shard.u_hi]).abs().max() / want.abs().max()),
             float((fi - want[U:]).abs().max() / want.abs().max()))
# phase times of the fused variant at a size where the exchange matters (optional: CHAOREC_CHECK_BIG=1)
if os.environ.get("CHAOREC_CHECK_BIG"):
    Ub, Ib, Eb = 8_388_608, 2_097_152, int(os.environ.get("CHAOREC_CHECK_EDGES", "100000000"))
    g = torch.Generator(device=dev).manual_seed(3)
    w = 1.0 / torch.arange(1, Ib + 1, dtype=torch.float64, device=dev)
    cdf = torch.cumsum(w / w.sum(), 0)
    perm = torch.randperm(Ib, generator=g, device=dev)
    bu = torch.randint(0, Ub, (Eb,), generator=g, device=dev)
    bi = perm[torch.searchsorted(cdf, torch.rand(Eb, generator=g, device=dev, dtype=torch.float64)).clamp_(max=Ib - 1)]
    keys = torch.unique(bu * Ib + bi)
    bu, bi = keys // Ib, keys % Ib
    del keys
    sh = cdist.build_bipartite_shard(bu, bi, Ub, Ib)
    exb = cdist.FusedItemExchange(Ib, d, dev)
    xb = torch.randn(Ub + Ib, d, device=dev, generator=g) * 0.1
    xu_b, xi_b = xb[sh.u_lo:sh.u_hi].contiguous(), xb[Ub:].contiguous()
    want_b = None
    if os.environ.get("CHAOREC_CHECK_BIG_PARITY"):      # the whole graph on every rank: single-GPU reference rows
        fullb = cb.bipartite_graph(bu, bi, Ub, Ib, "v1")
        want_b = cb.propagate_mean(fullb, xb, L)
        del fullb
    del xb
    for rep in range(4):
        tm = [] if rep == 3 else None
        torch.cuda.synchronize(); dist.barrier()
        fub, fib = cdist.propagate_mean_bipartite_fused(sh, exb, xu_b, xi_b, L, timing=tm)
        torch.cuda.synchronize()
    if rank == 0:
        names = ("push SpMM done", "barrier 0 passed", "reduce+broadcast done", "barrier 1 passed", "user rows done")
        t00 = tm[0][0]
        for l, marks in enumerate(tm):
            print(f"  fused layer {l} (ms since start): push starts {t00.elapsed_time(marks[0]):.3f}, " +
                  ", ".join(f"{n} {t00.elapsed_time(marks[i + 1]):.3f}" for i, n in enumerate(names)))
    if want_b is not None:
        sc = float(want_b.abs().max())
        eb = torch.tensor([float((fub - want_b[sh.u_lo:sh.u_hi]).abs().max()) / sc, float((fib - want_b[Ub:]).abs().max()) / sc], device=dev)
        allb = [torch.zeros_like(eb) for _ in range(world)]
        dist.all_gather(allb, eb)
        if rank == 0:
            print(f"  fused exchange on the big graph ({int(bu.numel())} edges), per rank (user rows, item rows) rel err: " +
                  " ".join(f"({float(e[0]):.1e},{float(e[1]):.1e})" for e in allb))
        ef = max(ef, float(eb.max()))
    del sh, exb
err = torch.tensor([max(eu, ei), ec, ef], device=dev)
dist.all_reduce(err, op=dist.ReduceOp.MAX)
ok = bool((err < 1e-5).all())
if rank == 0:
    print(f"sharded propagation on CUDA, world {world}: bipartite/NCCL all-reduce max rel err {float(err[0]):.3e}, "
          f"row-cyclic/NCCL all-gather {float(err[1]):.3e}, bipartite/fused P2P exchange {float(err[2]):.3e}", "OK" if ok else "FAIL")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
