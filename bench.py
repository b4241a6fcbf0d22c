#!/usr/bin/env python
"""bench.py -- headline benchmark of the chaorec_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs eval,propagate,knn]

Headline leg ("eval"): BASELINE.json configs[3], full-ranking top-20 evaluation of 1,000,000 users x
200,000 items, d=64, train mask of 16 Zipf items per user, 2 test positives per user, users sharded
contiguously over the N GPUs (strong scaling: the job is fixed, `value` = 1e6 users / max-over-ranks
time).  One step = one pass over all users: fused score+mask+top-K kernel, hit-mask kernel.
  value  device-resident inputs, CUDA-event time of the K steps (L2 flushed between steps)
  e2e    the same job through the host-buffer API: pinned host -> device copies of every input,
         kernels, device -> host copy of the hit masks, gather to rank 0 and the sequential fp64
         metric reduction on the host (wall clock between synchronised barriers)
Secondary legs (reported inside the same JSON line): "propagate" (configs[4]-shaped 3-layer
normalized-adjacency propagation, edges/s, HBM roofline), "knn" (configs[2], item rows/s) and "train"
(configs[1], LightGCN BPR training step, ms/step).
--impl reference times the oracle's restatement of the reference's CPU path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EVAL_CFG = dict(U=1_000_000, I=200_000, d=64, K=20, mask_per_user=16, pos_per_user=2, seed=20264)
PROP_CFG = dict(U=8_388_608, I=2_097_152, E=100_000_000, d=64, L=3, seed=20265)
KNN_CFG = dict(N=23_033, Dv=4096, Dt=384, k=10, seed=20263)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--legs", default="eval,propagate,knn,train")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every workload (debug only; marks the line)")
    ap.add_argument("--cpu-sample-users", type=int, default=None)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------
# synthetic inputs (host side, seeded; identical for every world size)
# --------------------------------------------------------------------------------------------


def make_eval_job(cfg, scale):
    from oracle import ref_ops  # seeded generators shared with the tests (inputs only)
    U, I, d = int(cfg["U"] * scale), int(cfg["I"] * scale), cfg["d"]
    u = ref_ops.normal_embeddings(U, d, cfg["seed"])
    it = ref_ops.normal_embeddings(I, d, cfg["seed"] + 1)
    g = torch.Generator().manual_seed(cfg["seed"] + 2)
    cdf, perm = ref_ops.zipf_item_sampler(I, g)
    mu = torch.arange(U).repeat_interleave(cfg["mask_per_user"])
    mi = perm[torch.searchsorted(cdf, torch.rand(mu.numel(), generator=g, dtype=torch.float64)).clamp_(max=I - 1)]
    keys = torch.unique(mu * I + mi)
    mu, mi = keys // I, keys % I
    mask_rowptr = torch.zeros(U + 1, dtype=torch.int64)
    mask_rowptr[1:] = torch.cumsum(torch.bincount(mu, minlength=U), 0)
    # positives: pos_per_user random items per user (global ids, sorted inside the row)
    pos = torch.randint(0, I, (U, cfg["pos_per_user"]), generator=g) + U
    pos = torch.sort(pos, dim=1)[0]
    return dict(U=U, I=I, d=d, K=cfg["K"], user=u, item=it, mask_rowptr=mask_rowptr, mask_col=mi.to(torch.int32), pos=pos)


def shard_eval_job(job, lo, hi, n_chunks=1):
    """Pinned host buffers of users [lo, hi); with n_chunks > 1 a list of consecutive chunks (for the
    copy/compute-overlapped end-to-end path)."""
    if n_chunks > 1:
        # a small first chunk (its copies are exposed), the rest in one piece (its copies hide behind
        # the first chunk's kernels)
        n = hi - lo
        bounds = [lo, lo + n // 5, hi]
        return [shard_eval_job(job, bounds[i], bounds[i + 1]) for i in range(2) if bounds[i + 1] > bounds[i]]
    from chaorec_b200.ranking import HostEvalBatch
    rp = job["mask_rowptr"]
    m_lo, m_hi = int(rp[lo]), int(rp[hi])
    P = job["pos"].shape[1]
    n = hi - lo
    return HostEvalBatch(
        user_emb=job["user"][lo:hi].clone(), item_emb=job["item"],
        mask_rowptr=(rp[lo:hi + 1] - m_lo).clone(), mask_col=job["mask_col"][m_lo:m_hi].clone(),
        pos_rowptr=torch.arange(0, (n + 1) * P, P, dtype=torch.int64), pos_col=job["pos"][lo:hi].reshape(-1).clone(),
        n_pos=np.full(n, P, dtype=np.int32), id_offset=job["U"]).pin()


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------


def flush_l2(buf):
    buf.add_(1)


def probe_fp32_peak(dev):
    from chaorec_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    flops = C.c_double(0.0)
    best = 0.0
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.chaorec_probe_fp32_fma(20000, _lib.ptr(sink), C.byref(flops), _lib.stream_ptr(dev)), "probe")
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def probe_tmem_read(dev, warps=16):
    """TB/s at which `warps` warps per SM read fp32 accumulators out of tensor memory (nothing else running)."""
    from chaorec_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    nbytes = C.c_double(0.0)
    best = 0.0
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.chaorec_probe_tmem_read(4000, warps, _lib.ptr(sink), C.byref(nbytes), _lib.stream_ptr(dev)), "probe")
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, nbytes.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_eval_leg(args, dev, rank, world, peaks):
    import torch.distributed as dist
    import chaorec_b200 as cb
    from chaorec_b200 import _lib, dist as cdist
    from chaorec_b200.ranking import MaskCSR, evaluate_host_batch, metrics_from_hits

    job = make_eval_job(EVAL_CFG, args.scale)
    U, I, d, K = job["U"], job["I"], job["d"], job["K"]
    lo, hi = cdist.shard_range(U, world, rank)
    batch = shard_eval_job(job, lo, hi)
    e2e_chunks = shard_eval_job(job, lo, hi, n_chunks=2 if (hi - lo) >= 400_000 else 1)
    n_local = hi - lo
    ks = [5, 10, 20]

    # ---- device-resident leg -----------------------------------------------------------------
    ue, ie = batch.user_emb.to(dev), batch.item_emb.to(dev)
    mask = MaskCSR(batch.mask_rowptr.to(dev), batch.mask_col.to(dev))
    pos_rp, pos_col = batch.pos_rowptr.to(dev), batch.pos_col.to(dev)
    lib = _lib.load()
    hits = torch.empty(n_local, dtype=torch.int64, device=dev)
    l2buf = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    st = {}
    cb.rank_topk(ue, ie, mask, K, 1e-6, U, stats=st)            # which engine, how many rows fell back

    def one_step(timed):
        flush_l2(l2buf)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ranks = cb.rank_topk(ue, ie, mask, K, 1e-6, U)
        e1.record()
        _lib.check(lib.chaorec_topk_hits(_lib.ptr(ranks), None, _lib.ptr(pos_rp), _lib.ptr(pos_col), n_local, K,
                                         _lib.ptr(hits), _lib.stream_ptr(dev)), "hits")
        e2.record()
        return e0, e1, e2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop()
    t_total = sum(e0.elapsed_time(e2) for e0, _, e2 in evs) * 1e-3
    t_rank_kernel = sum(e0.elapsed_time(e1) for e0, e1, _ in evs) * 1e-3 / args.steps
    tt = torch.tensor([t_total, t_rank_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_rank_kernel = float(tt[0]), float(tt[1])
    value = U * args.steps / t_total

    # ---- end-to-end leg (host buffers in, metrics out) ------------------------------------------
    n_pos_all = np.full(U, job["pos"].shape[1], dtype=np.int32)

    def e2e_step():
        h = evaluate_host_batch(e2e_chunks, K, dev)          # pinned host tensor, synchronised
        if world > 1:
            all_h = cdist.gather_to_rank0(h.to(dev))
            if rank != 0:
                return None
            hh = all_h.cpu().numpy()
        else:
            hh = h.numpy()
        return metrics_from_hits(hh, n_pos_all, ks)

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        metrics = e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = U * args.steps / float(te[0])

    fp32_peak = probe_fp32_peak(dev)
    flops = 2.0 * n_local * I * d
    achieved = flops / t_rank_kernel / 1e12
    tensor = st.get("engine") == "tensor"
    if tensor:
        roof = dict(kernel="tc_candidates_kernel (tcgen05 bf16 pre-filter; time includes bf16 conversion, fp32 re-score + "
                           "certification and the exact fallback launch)", bound="tensor", achieved=achieved,
                    peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops"],
                    # dram__bytes_read+write of the main tc_candidates_kernel<true> launch at this exact workload on 1 GPU,
                    # from the ncu --set full capture profiles/r01_final_tc_eval_summary.csv (1.39 GB read + 1.65 GB written)
                    traffic=(3.05e9 if (world == 1 and args.scale == 1.0) else None),
                    tmem_read_tbps_measured=probe_tmem_read(dev),
                    peak_source=peaks["source"] + " (bf16 dense burst)", algorithmic_flops_per_launch=flops,
                    kernel_ms=t_rank_kernel * 1e3, fallback_rows=st.get("fallback_rows"),
                    fp32_fma_peak_tflops_measured=fp32_peak)
    else:
        roof = dict(kernel="rank_topk_kernel<true>", bound="fp32_fma", achieved=achieved, peak=fp32_peak, unit="TFLOP/s",
                    frac=achieved / fp32_peak if fp32_peak else None, traffic=None,
                    peak_source="chaorec_probe_fp32_fma measured live on this GPU (MEASURED_PEAKS.json has no fp32 entry)",
                    frac_of_bf16_tensor_peak=achieved / peaks["bf16_tflops"], algorithmic_flops_per_launch=flops,
                    kernel_ms=t_rank_kernel * 1e3)
    out = dict(
        value=value, ms_per_step=t_total / args.steps * 1e3, clocks=clocks,
        e2e=dict(value=e2e_value, unit="users/s", h2d_bytes_per_step=batch.h2d_bytes(), d2h_bytes_per_step=n_local * 8,
                 includes="H2D of embeddings+mask+positives, rank kernel, hits kernel, D2H hit masks, gather, host fp64 metrics"),
        roofline=roof, engine=st.get("engine"),
        gpu_launches=(11 if tensor else 2) * args.steps,   # to_bf16 x2, max_reduce, gather_rows, tc_candidates x2, rescore, rank_topk x2 + slice_merge (fallback rows), topk_hits
        config=dict(workload="full-rank top-20 eval, 1M users x 200k items, d=64 (BASELINE configs[3])", users=U, items=I,
                    d=d, topk=K, mask_items_per_user=EVAL_CFG["mask_per_user"], sharding=f"users/{world}",
                    l2="flushed between steps (192 MiB write)", scale=args.scale),
        metrics_at_20=metrics[20] if rank == 0 and metrics else None)
    return out, job


def run_cpu_baseline_eval(job, n_users, threads=None):
    """The reference's CPU path (Model/MMGCN.py:204-244 batched form + utils.gene_metrics) on a seeded
    sample of the same job, timed on the host cores."""
    from oracle import ref_ops
    if threads:
        torch.set_num_threads(threads)
    U, I, K = job["U"], job["I"], job["K"]
    n = min(n_users, U)
    g = torch.Generator().manual_seed(7)
    sample = torch.randperm(U, generator=g)[:n].sort()[0]
    rp, mc = job["mask_rowptr"], job["mask_col"]
    uid = {j: (mc[int(rp[u]):int(rp[u + 1])].to(torch.int64) + n).tolist() for j, u in enumerate(sample.tolist())}
    result = torch.cat((job["user"][sample], job["item"]), 0)
    rows = np.empty(n, dtype=object)
    pos = (job["pos"][sample] - U + n).tolist()
    for j in range(n):
        rows[j] = [j] + pos[j]
    t0 = time.perf_counter()
    rank = ref_ops.gene_ranklist_batched(result, n, I, uid, step=2048, topk=K, mask_value=1e-6)
    m = ref_ops.gene_metrics(rows, rank, [5, 10, 20])
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit="users/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{n} seeded users of the same job (all {I} items), batched reference form step=2048 + gene_metrics; "
                       f"{dt:.1f} s of CPU work, users independent so users/s extrapolates linearly"), m


def gpu_zipf_graph(cfg, scale, dev):
    """configs[4]-shaped bipartite graph generated on the device (100M edges do not fit a CPU budget)."""
    U, I, E = int(cfg["U"] * scale), int(cfg["I"] * scale), int(cfg["E"] * scale)
    g = torch.Generator(device=dev).manual_seed(cfg["seed"])
    w = 1.0 / torch.arange(1, I + 1, dtype=torch.float64, device=dev)
    cdf = torch.cumsum(w / w.sum(), 0)
    perm = torch.randperm(I, generator=g, device=dev)
    base = torch.arange(U, device=dev).repeat_interleave(3)
    users = torch.cat((base, torch.randint(0, U, (max(E - base.numel(), 0),), generator=g, device=dev)))
    items = perm[torch.searchsorted(cdf, torch.rand(users.numel(), generator=g, device=dev, dtype=torch.float64)).clamp_(max=I - 1)]
    keys = torch.unique(users * I + items)
    del users, items, base
    return U, I, keys // I, keys % I


def run_propagate_leg(args, dev, rank, world, peaks):
    import chaorec_b200 as cb
    if world > 1:
        return run_propagate_leg_sharded(args, dev, rank, world, peaks)
    cfg = PROP_CFG
    U, I, tu, ti = gpu_zipf_graph(cfg, args.scale, dev)
    n, d, L = U + I, cfg["d"], cfg["L"]
    E = int(tu.numel())
    row = torch.cat((tu, ti + U))
    col = torch.cat((ti + U, tu))
    del tu, ti
    graph = cb.csr_from_coo(row, col, None, n, n)
    del row, col
    from chaorec_b200 import _lib
    lib = _lib.load()
    dis = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.check(lib.chaorec_degree_inv_sqrt(_lib.ptr(graph.rowptr), n, 0.0, _lib.ptr(dis), _lib.stream_ptr(dev)), "deg")
    graph.row_scale = graph.col_scale = dis
    graph.symmetric = True
    x0 = torch.randn(n, d, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 0.1
    for _ in range(args.warmup):
        cb.propagate_mean(graph, x0, L)
    torch.cuda.synchronize()
    times = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cb.propagate_mean(graph, x0, L)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.mean(times))
    nnz = 2 * E
    bytes_layer = nnz * (4 + 4 * d) + n * (4 * d + 8)          # 260 B/edge + 264 B/node
    achieved = bytes_layer * L / t / 1e9
    return dict(metric="GCN propagate edges/s", value=nnz * L / t, unit="edges/s", ms_per_step=t * 1e3,
                config=dict(workload="3-layer normalized-adjacency propagation + layer mean (BASELINE configs[4] shape)",
                            users=U, items=I, edges=E, nnz_sym=nnz, d=d, layers=L, hub_rows=graph.n_hub,
                            hub_chunks=graph.n_hub_chunks, l2="inputs (2.7 GB table, 0.8 GB indices) exceed L2", scale=args.scale),
                roofline=dict(kernel="spmm_rows_kernel<16,1,64,20> (+hub kernels)", bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"],
                              unit="GB/s", frac=achieved / peaks["hbm_gbs"],
                              # ncu dram bytes per layer at this workload (profiles/r01_final_spmm_*_summary.csv): rows
                              # kernel 21.2 GB read + 5.3 GB written, hub-chunk kernel 20.6 GB read + 0.2 GB written
                              # -> 47.3 GB against 52.9 GB algorithmic (L2 absorbs the popular item rows)
                              traffic=(47.3e9 if args.scale == 1.0 else None), peak_source=peaks["source"],
                              algorithmic_bytes_per_layer=bytes_layer),
                gpu_launches=(1 + (2 if graph.n_hub else 0)) * L * args.steps)


def run_propagate_leg_sharded(args, dev, rank, world, peaks):
    """N > 1: users stay on their rank, item partials are all-reduced once per layer (dist.BipartiteShard);
    the generic row-cyclic + all-gather variant (any adjacency) is timed beside it."""
    import torch.distributed as dist
    import chaorec_b200 as cb
    from chaorec_b200 import _lib, dist as cdist
    cfg = PROP_CFG
    U, I, tu, ti = gpu_zipf_graph(cfg, args.scale, dev)          # same seed on every rank -> same graph
    n, d, L = U + I, cfg["d"], cfg["L"]
    E = int(tu.numel())

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        times = []
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            dist.barrier(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        tt = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0])

    # --- bipartite shard -------------------------------------------------------------------------
    shard = cdist.build_bipartite_shard(tu, ti, U, I)
    gen = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.randn(n, d, device=dev, generator=gen) * 0.1
    xu0 = x0[shard.u_lo:shard.u_hi].contiguous()
    xi0 = x0[U:].contiguous()
    del x0
    t = timed(lambda: cdist.propagate_mean_bipartite(shard, xu0, xi0, L))
    del shard, xu0, xi0
    torch.cuda.empty_cache()
    # --- generic row-cyclic partition, one all-gather of the whole table per layer ------------------
    part = cdist.CyclicPartition(n, world)
    deg = torch.bincount(torch.cat((tu, ti + U)), minlength=n)
    dis = deg.to(torch.float32).pow(-0.5)
    rows, cols = [], []
    for r_, c_ in ((tu, ti + U), (ti + U, tu)):
        lr, pc, _ = part.local_edges(r_, c_, None, rank)
        rows.append(lr); cols.append(pc)
    del tu, ti
    graph = cb.csr_from_coo(torch.cat(rows), torch.cat(cols), None, part.block, part.n_padded)
    del rows, cols
    dis_perm = part.permute_rows(dis.unsqueeze(1)).squeeze(1).contiguous()
    dis_perm[torch.isinf(dis_perm)] = 0
    graph.row_scale = dis_perm[rank * part.block:(rank + 1) * part.block].contiguous()
    graph.col_scale = dis_perm
    x0 = torch.randn(part.n_padded, d, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 0.1
    t_cyclic = timed(lambda: cdist.propagate_mean_sharded(graph, part, x0, L))
    nnz = 2 * E
    bytes_layer = nnz * (4 + 4 * d) + n * (4 * d + 8)
    achieved = bytes_layer * L / t / 1e9 / world
    return dict(metric="GCN propagate edges/s", value=nnz * L / t, unit="edges/s", ms_per_step=t * 1e3,
                generic_row_cyclic_allgather=dict(value=nnz * L / t_cyclic, unit="edges/s", ms_per_step=t_cyclic * 1e3),
                config=dict(workload="3-layer propagation, users sharded and kept local, one NCCL all-reduce of the item "
                                     "partials [I, d] per layer",
                            users=U, items=I, edges=E, nnz_sym=nnz, d=d, layers=L, world=world, scale=args.scale,
                            exchanged_bytes_per_layer=I * d * 4),
                roofline=dict(kernel="spmm_rows_kernel<16,1,64,20> (+hub kernels)", bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"],
                              unit="GB/s per GPU", frac=achieved / peaks["hbm_gbs"], traffic=None, peak_source=peaks["source"]),
                gpu_launches=None)


def run_knn_leg(args, dev, rank, world, peaks):
    import chaorec_b200 as cb
    from chaorec_b200 import dist as cdist
    cfg = KNN_CFG
    N = int(cfg["N"] * args.scale)
    out = {}
    lo, hi = cdist.shard_range(N, world, rank)
    for name, D in (("t_feat", cfg["Dt"]), ("v_feat", cfg["Dv"])):
        g = torch.Generator(device=dev).manual_seed(cfg["seed"] + D)
        x = torch.randn(N, D, device=dev, generator=g)
        if name == "v_feat":
            x = torch.relu(x)
        kst = {}
        cb.knn_topk(x, cfg["k"], row_range=(lo, hi), stats=kst)
        for _ in range(max(1, args.warmup - 1)):
            cb.knn_topk(x, cfg["k"], row_range=(lo, hi))
        torch.cuda.synchronize()
        times = []
        for _ in range(max(2, args.steps // 2)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            val, idx = cb.knn_topk(x, cfg["k"], row_range=(lo, hi))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        tt = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
        flops = 2.0 * N * N * D
        cpu_ms = None
        if world == 1 and name == "t_feat":
            # the reference's own form (Model/FREEDOM.py:111-128: normalise, dense sim, topk) on the host cores
            from oracle import ref_ops
            xc = x.cpu()
            t0 = time.perf_counter()
            ref_ops.knn_conv_a(xc, cfg["k"])
            cpu_ms = (time.perf_counter() - t0) * 1e3
        out[name] = dict(value=N / t, unit="item rows/s", ms=t * 1e3, cpu_reference_ms=cpu_ms, D=D, N=N, k=cfg["k"], tflops=flops / t / 1e12,
                         frac_of_bf16_tensor_peak=flops / t / 1e12 / peaks["bf16_tflops"], engine=kst.get("engine"),
                         fallback_rows=kst.get("fallback_rows"))
    return dict(metric="kNN graph item rows/s", config=dict(workload="cosine kNN graph, clothing-shaped (BASELINE configs[2])"), **out)


TRAIN_CFG = dict(U=35_598, I=18_357, E=296_337, d=64, L=3, batch=1024, reg=1e-4, seed=20262)


def run_train_leg(args, dev, rank, world, peaks):
    """BASELINE configs[1]: sports-shaped LightGCN, one BPR training step = fused propagation forward,
    BPR loss + gradient kernel, fused propagation backward, Adam (torch, fused).  Replicas only at N > 1."""
    import chaorec_b200 as cb
    from oracle import ref_ops
    cfg = TRAIN_CFG
    U, I, d, L, B = int(cfg["U"] * args.scale), int(cfg["I"] * args.scale), cfg["d"], cfg["L"], cfg["batch"]
    tu, ti = ref_ops.synth_interactions(U, I, int(cfg["E"] * args.scale), cfg["seed"])
    ei_cpu = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy())
    ei = ei_cpu.to(dev)
    uw = ref_ops.xavier_embeddings(U, d, 1).to(dev).requires_grad_(True)
    iw = ref_ops.xavier_embeddings(I, d, 2).to(dev).requires_grad_(True)
    opt = torch.optim.Adam([uw, iw], lr=1e-3, fused=True)
    g = torch.Generator().manual_seed(5)
    n_batches = 16
    pick = torch.randint(0, tu.numel(), (n_batches, B), generator=g)
    users_h, pos_h = tu[pick], ti[pick] + U
    neg_h = torch.randint(0, I, (n_batches, B), generator=g) + U
    users, pos, neg = users_h.to(dev), pos_h.to(dev), neg_h.to(dev)

    def step(i):
        opt.zero_grad(set_to_none=True)
        loss, _ = cb.lightgcn_loss(uw, iw, ei, L, users[i % n_batches], pos[i % n_batches], neg[i % n_batches], cfg["reg"])
        loss.backward()
        opt.step()
        return loss

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize()
    n = max(20, args.steps * 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nnz = int(ei.shape[1])
    # the same step replayed from a CUDA graph (the step is launch-bound at this size)
    ms_graph = None
    try:
        # fresh leaves: the eager steps above tied uw/iw's gradient accumulators to the default stream
        guw, giw = uw.detach().clone().requires_grad_(True), iw.detach().clone().requires_grad_(True)
        gopt = torch.optim.Adam([guw, giw], lr=1e-3, capturable=True, fused=True)
        gstep = cb.GraphedLightGCNStep(guw, giw, ei, L, cfg["reg"], B, gopt)
        for i in range(3):
            gstep(users[i], pos[i], neg[i])
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            gl = gstep(users[i % n_batches], pos[i % n_batches], neg[i % n_batches])
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / n
    except Exception as exc:  # capture support is environment dependent; the eager number stands
        import traceback
        traceback.print_exc(file=sys.stderr)
        ms_graph = f"unavailable: {type(exc).__name__}: {exc}"[:200]
    out = dict(metric="LightGCN BPR training step", ms_per_step=ms, ms_per_step_cuda_graph=ms_graph, steps_per_s=1e3 / ms, propagate_edges_per_s=nnz * L * 2 / (ms * 1e-3),
               loss=float(loss), config=dict(workload="LightGCN n_layers=3 d=64, sports-shaped (BASELINE configs[1]): propagation fwd+bwd, "
                                             "BPR loss+grad kernel, Adam", users=U, items=I, nnz_sym=nnz, batch=B, scale=args.scale),
               gpu_launches=2 * L + 2)
    if rank == 0 and world == 1:
        # CPU baseline: the oracle's restatement of the same step (reference arithmetic) on the host cores
        cu = ref_ops.xavier_embeddings(U, d, 1).requires_grad_(True)
        ci = ref_ops.xavier_embeddings(I, d, 2).requires_grad_(True)
        copt = torch.optim.Adam([cu, ci], lr=1e-3)
        t0 = time.perf_counter()
        for i in range(3):
            copt.zero_grad(set_to_none=True)
            l = ref_ops.lightgcn_loss(cu, ci, ei_cpu, L, users_h[i], pos_h[i], neg_h[i], cfg["reg"])
            l.backward()
            copt.step()
        out["cpu_baseline"] = dict(value=(time.perf_counter() - t0) / 3 * 1e3, unit="ms/step", cores=torch.get_num_threads(), kind="port",
                                   sample="3 steps of the same model on the host cores (PyG-form gather/scatter propagation + autograd + Adam)")
    return out


def main_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a GPU: chaorec_b200 has no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    legs = [s for s in args.legs.split(",") if s]
    line = dict(metric="eval users/s (full-rank top-20)", unit="users/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic", impl="ours")
    job = None
    if "eval" in legs:
        ev, job = run_eval_leg(args, dev, rank, world, peaks)
        line.update(ev)
    if "propagate" in legs:
        torch.cuda.empty_cache()
        line["propagate"] = run_propagate_leg(args, dev, rank, world, peaks)
    if "knn" in legs:
        torch.cuda.empty_cache()
        line["knn"] = run_knn_leg(args, dev, rank, world, peaks)
    if "train" in legs:
        torch.cuda.empty_cache()
        line["train"] = run_train_leg(args, dev, rank, world, peaks)
    if rank == 0 and world == 1 and job is not None:
        cpu, _ = run_cpu_baseline_eval(job, args.cpu_sample_users or 131072)       # ~13 s of host work on 16 cores
        line["cpu_baseline"] = cpu
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main_reference(args):
    """The reference arm: the oracle's restatement of the reference's CPU evaluation path (the
    reference is pure Python over torch CPU ops; it cannot be pip-installed and /root/reference does
    not travel), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    job = make_eval_job(EVAL_CFG, args.scale)
    n = args.cpu_sample_users or 8192
    for _ in range(min(args.warmup, 1)):
        run_cpu_baseline_eval(job, min(n, 512))
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        cpu, _ = run_cpu_baseline_eval(job, n)
        done += min(n, job["U"])
    dt = time.perf_counter() - t0
    v = done / dt
    print(json.dumps(dict(
        metric="eval users/s (full-rank top-20)", value=v, unit="users/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
        data="synthetic", impl="reference",
        config=dict(workload="full-rank top-20 eval, 1M users x 200k items, d=64 (BASELINE configs[3])", users=job["U"],
                    items=job["I"], d=job["d"], topk=job["K"], scale=args.scale),
        cpu_baseline=dict(value=v, unit="users/s", cores=torch.get_num_threads(), kind="port",
                          sample=f"each step = {n} seeded users of the job against all items (batched reference form + gene_metrics)"),
        e2e=dict(value=v, unit="users/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
