"""torchrun --nproc-per-node N scripts/gpu_bipartite_check.py : BipartiteShard propagation (CUDA + NCCL, blocked
all-reduce) against the single-GPU fused propagation of the whole graph computed redundantly on every rank."""
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
err = torch.tensor([max(eu, ei)], device=dev)
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"bipartite sharded propagation, world {world}: max rel err {float(err):.3e}", "OK" if float(err) < 1e-5 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if float(err) < 1e-5 else 1)
