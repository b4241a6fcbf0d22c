"""World-size-2 dry run of bench.py's sharded propagation leg over gloo on CPU tensors: the CUDA entry points are replaced
by torch CPU stand-ins, so that the leg's host logic -- input broadcast + consistency record, collective gates, fall-back
from the fused exchange (unavailable without CUDA) to the NCCL form, the row-cyclic variant, the result dict -- runs
end to end without a GPU.  (The kernels themselves are covered by the -m gpu tests and scripts/gpu_multi_parity.sh.)"""
import os
import socket
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Graph:
    """Stand-in for CSRGraph: a torch sparse matrix plus the scales the kernels apply."""

    def __init__(self, row, col, M, N):
        self.M, self.N = M, N
        self.a = torch.sparse_coo_tensor(torch.stack((row, col)), torch.ones(row.numel()), (M, N)).coalesce()
        self.rowptr = torch.zeros(M + 1, dtype=torch.int64)
        self.rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
        self.col = col
        self.row_scale = self.col_scale = None
        self.n_hub = self.n_hub_chunks = 0

    def plan_hubs(self):
        return self

    def product(self, x):
        cs = self.col_scale if self.col_scale is not None else torch.ones(self.N)
        rs = self.row_scale if self.row_scale is not None else torch.ones(self.M)
        return rs.unsqueeze(1) * torch.sparse.mm(self.a, cs.unsqueeze(1) * x)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        import bench
        import chaorec_b200 as cb
        from chaorec_b200 import dist as cdist, propagate as P

        def csr_from_coo(row, col, val, M, N):
            return _Graph(row.long(), col.long(), M, N)

        def spmm_raw(g, x, y, acc, alpha, mode):
            out = g.product(x)
            if y is not None:
                y.copy_(out)
            if acc is not None:
                if mode == 1:
                    acc.copy_(alpha * x + alpha * out)
                else:
                    acc.add_(out, alpha=alpha)

        def propagate_mean(g, x0, L):
            acc, cur = x0 / (L + 1), x0
            for _ in range(L):
                cur = g.product(cur)
                acc = acc + cur / (L + 1)
            return acc

        def build_full_graph(tu, ti, U, I, dev):
            n = U + I
            g = _Graph(torch.cat((tu, ti + U)), torch.cat((ti + U, tu)), n, n)
            deg = (g.rowptr[1:] - g.rowptr[:-1]).float()
            dis = deg.pow(-0.5)
            dis[torch.isinf(dis)] = 0
            g.row_scale = g.col_scale = dis
            return g, dis

        P.csr_from_coo, P._spmm_raw, P.CSRGraph = csr_from_coo, spmm_raw, lambda rp, col, val, M, N: None
        cb.csr_from_coo, cb.propagate_mean = csr_from_coo, propagate_mean
        bench.build_full_graph = build_full_graph

        class _Event:
            def __init__(self, enable_timing=False):
                import time
                self.t = time.perf_counter()

            def record(self, *a):
                import time
                self.t = time.perf_counter()

            def elapsed_time(self, other):
                return max((other.t - self.t) * 1e3, 1e-3)

        torch.cuda.Event, torch.cuda.synchronize, torch.cuda.empty_cache = _Event, (lambda *a, **k: None), (lambda: None)
        os.environ["CHAOREC_BIPARTITE_BLOCKS"] = "1"
        args = types.SimpleNamespace(scale=2e-4, warmup=1, steps=1, skip_cpu_legs=True, cpu_legs=False)
        peaks = dict(hbm_gbs=6550.0, bf16_tflops=1690.0, bf16_tflops_sustained=1412.0, source="test")
        out = bench.run_propagate_leg_sharded(args, torch.device("cpu"), rank, world, peaks)
        if rank == 0:
            results["out"] = out
    finally:
        dist.destroy_process_group()


def test_sharded_propagation_leg_host_logic_world_2():
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    out = results["out"]
    par = out["parity"]
    assert out["value"] > 0 and "NCCL" in out["exchange"]                      # fused exchange needs CUDA: fell back
    assert out["fused_note"] and "unavailable" in out["fused_note"]
    assert par["bipartite_nccl_rel_err"] <= 1e-5 and par["row_cyclic_nccl_rel_err"] <= 1e-5
    assert par["inputs_identical_across_ranks"] == dict(graph=True, layer0_table=True, note=par["inputs_identical_across_ranks"]["note"])
    assert isinstance(out["bipartite_nccl_allreduce"], dict) and isinstance(out["generic_row_cyclic_allgather"], dict)
