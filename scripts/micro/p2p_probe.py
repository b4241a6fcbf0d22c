"""torchrun --nproc-per-node N scripts/micro/p2p_probe.py : how fast can this box move bytes between GPUs, by path?
  (1) copy engine:  peer_table.copy_(local)            (cudaMemcpyPeerAsync)
  (2) SM stores:    torch.add(local, 0, out=peer_view) (a plain elementwise kernel whose OUTPUT is peer memory)
  (3) SM loads:     local.copy_ from a peer view through an elementwise kernel (torch.add(peer_view, 0, out=local))
  (4) the owner-side reduce+broadcast kernel with all destinations local / all real
All ranks run the same thing at the same time (every link busy in both directions)."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chaorec_b200 import _lib  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
name = dist.group.WORLD.group_name
d = 64
I = 2_097_152
B = I // world
buf = symm.empty((world * B, d), dtype=torch.float32, device=dev)
h = symm.rendezvous(buf, name)
slots = symm.empty((world, B, d), dtype=torch.float32, device=dev)
hs = symm.rendezvous(slots, name)
local = torch.randn(B, d, device=dev)
acc = torch.zeros(B, d, device=dev)
peer = (rank + 1) % world
pv = h.get_buffer(peer, (B, d), torch.float32, rank * B * d)      # my block inside the next rank's table


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    return e0.elapsed_time(e1) / n


mb = B * d * 4 / 1e6
t = timed(lambda: pv.copy_(local))
r1 = f"copy engine -> 1 peer: {mb / t:.0f} GB/s"
t = timed(lambda: torch.add(local, 0.0, out=pv))
r2 = f"SM stores -> 1 peer: {mb / t:.0f} GB/s"
t = timed(lambda: torch.add(pv, 0.0, out=local))
r3 = f"SM loads <- 1 peer: {mb / t:.0f} GB/s"
views = [h.get_buffer(p, (B, d), torch.float32, rank * B * d) for p in range(world)]


def bcast_ce():
    for p in range(world):
        if p != rank:
            views[p].copy_(local, non_blocking=True)


t = timed(bcast_ce)
r4 = f"copy engine -> {world - 1} peers: {(world - 1) * mb / t:.0f} GB/s out per GPU"


def bcast_sm():
    for p in range(world):
        if p != rank:
            torch.add(local, 0.0, out=views[p])


t = timed(bcast_sm)
r5 = f"SM stores -> {world - 1} peers (one kernel each): {(world - 1) * mb / t:.0f} GB/s out per GPU"
lib = _lib.load()
real = (C.c_void_p * world)(*[int(p) for p in h.buffer_ptrs])
mine = (C.c_void_p * world)(*[int(h.buffer_ptrs[rank])] * world)


def kb(ptrs):
    _lib.check(lib.chaorec_reduce_slots_broadcast_f32(_lib.ptr(slots), world, B, B, d, rank * B, ptrs, _lib.ptr(acc), 0.25, 0, None, 0, rank,
                                                       _lib.stream_ptr(dev)), "kb")


t_local = timed(lambda: kb(mine))
t_real = timed(lambda: kb(real))
r6 = f"reduce+broadcast kernel: all destinations local {t_local:.3f} ms, real peers {t_real:.3f} ms ({(world - 1) * mb / t_real:.0f} GB/s out)"
if rank == 0:
    print(f"world {world}, block {mb:.0f} MB")
    for r in (r1, r2, r3, r4, r5, r6):
        print(" ", r)
dist.destroy_process_group()
