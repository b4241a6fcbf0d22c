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
    ap.add_argument("--legs", default="eval,propagate,knn,train,seam")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every workload (debug only; marks the line)")
    ap.add_argument("--cpu-sample-users", type=int, default=None)
    ap.add_argument("--cpu-legs", action="store_true", help="run the host-side baseline legs even at --scale < 1")
    ap.add_argument("--skip-cpu-legs", action="store_true", help="profiling runs: no host-side baseline legs")
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
# parity gates (outside every timed region; a mismatch aborts the run -- no number without parity)
# --------------------------------------------------------------------------------------------


class ParityError(SystemExit):
    pass


def parity_fail(what, detail):
    raise ParityError(f"PARITY GATE FAILED ({what}): {detail} -- no bench line is printed for a wrong result")


def collective_gate(fn, world, dev, what):
    """Runs a per-rank gate; a failure on any rank raises on EVERY rank (nobody is left waiting in a collective)."""
    msg, out = None, None
    try:
        out = fn()
    except ParityError as exc:
        msg = str(exc)
    if world > 1:
        import torch.distributed as dist
        flag = torch.tensor([1 if msg else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag) and not msg:
            msg = f"PARITY GATE FAILED ({what}) on another rank"
    if msg:
        raise ParityError(msg)
    return out


def gate_eval(job, lo, hi, dev, ue, ie, mask, world, n_users=8192):
    """A seeded sample of this rank's users + EVERY row the tensor engine handed to the exact engine, re-derived by the
    oracle (oracle/canon.c: canonical fmaf chain, mask value, lowest-index tie-break): ids and score bits must be equal;
    the hit masks / metrics of the same users must equal the oracle's restatement of utils.gene_metrics."""
    import chaorec_b200 as cb
    from chaorec_b200 import _lib
    from chaorec_b200.ranking import metrics_from_hits
    from oracle.canon_py import Canon
    from oracle import ref_ops
    canon = Canon()
    U, I, K = job["U"], job["I"], job["K"]
    n_local = hi - lo
    st = {}
    idx, sc = cb.rank_topk(ue, ie, mask, K, 1e-6, U, return_scores=True, stats=st)
    g = torch.Generator().manual_seed(97 + lo)
    take = min(max(n_users // world, 64), n_local)
    sample = torch.randperm(n_local, generator=g)[:take]
    fb = st.get("fallback_row_ids")
    n_fb = int(fb.numel()) if fb is not None else 0
    if fb is not None and n_fb:
        sample = torch.unique(torch.cat((sample, fb[:4096])))
    sample = sample.sort()[0]
    rp, mc = job["mask_rowptr"], job["mask_col"]
    gl = sample + lo
    lens = (rp[gl + 1] - rp[gl])
    sub_rp = torch.zeros(sample.numel() + 1, dtype=torch.int64)
    sub_rp[1:] = torch.cumsum(lens, 0)
    sub_col = torch.cat([mc[int(rp[u]):int(rp[u + 1])] for u in gl.tolist()]) if sample.numel() else mc[:0]
    t0 = time.perf_counter()
    cidx, csc = canon.score_mask_topk(job["user"][gl].numpy(), job["item"].numpy(), sub_rp.numpy(), sub_col.numpy(), 1e-6, K, id_offset=U)
    t_oracle = time.perf_counter() - t0
    sd = sample.to(dev)
    got_idx, got_sc = idx[sd].cpu().numpy(), sc[sd].cpu().numpy()
    bad_rows = int((got_idx != cidx).any(axis=1).sum())
    bad_scores = int((got_sc.view(np.uint32) != csc.view(np.uint32)).any(axis=1).sum())
    if bad_rows or bad_scores:
        parity_fail("eval", f"{bad_rows} of {sample.numel()} sampled users differ in top-{K} ids, {bad_scores} in score bits")
    # metrics of the sampled users: reference python loop (restated) vs hit-mask kernel + host reduction
    rows = np.empty(sample.numel(), dtype=object)
    pos = job["pos"][gl].tolist()
    for j in range(sample.numel()):
        rows[j] = [j] + pos[j]
    ks = [5, 10, 20]
    want = ref_ops.gene_metrics(rows, torch.from_numpy(cidx), ks)
    got = cb.evaluate(None, rows, torch.from_numpy(got_idx), ks, device=dev)
    if any(got[k][n] != want[k][n] for k in ks for n in got[k]):
        parity_fail("eval metrics", f"{got} != {want}")
    return dict(users_checked=int(sample.numel()), fallback_rows=st.get("fallback_rows"), fallback_rows_checked=min(n_fb, 4096),
                ids="bit-exact", scores="bit-exact", metrics="identical (==)", oracle="oracle/canon.c + ref_ops.gene_metrics",
                oracle_s=round(t_oracle, 2)), st


def fp64_reference_chain(graph, dis, x0, L):
    """Independent fp64 product chain + layer mean on the device: torch.sparse (cuSPARSE) CSR @ dense in float64 on the
    SAME CSR pattern with the edge weights d^-1/2[r] d^-1/2[c] formed in float64 (the checker, not the product)."""
    n = graph.M
    counts = graph.rowptr[1:] - graph.rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=x0.device), counts)
    d64 = dis.double()
    vals = d64[rows] * d64[graph.col.long()]
    del rows
    adj = torch.sparse_csr_tensor(graph.rowptr, graph.col.long(), vals, size=(n, graph.N))
    x = x0.double()
    acc = x.clone()
    for _ in range(L):
        x = torch.sparse.mm(adj, x)
        acc += x
    return acc / (L + 1), adj


def gate_propagate(res, ref, what, tol=1e-5):
    err = float((res.double() - ref).abs().max() / ref.abs().max())
    if not (err <= tol):
        parity_fail(what, f"max|delta|/max|ref| = {err:.3e} > {tol:g}")
    return err


def gate_knn(x, k, lo, hi, val, idx, n_windows=3, width=24):
    """A few row windows of this rank's range re-derived by the oracle (canonical normalisation + fmaf chain):
    neighbour ids and similarity bits must be equal."""
    from oracle.canon_py import Canon
    canon = Canon()
    xh = x.cpu().numpy()
    g = torch.Generator().manual_seed(11 + lo)
    n = hi - lo
    starts = sorted({0, max(n - width, 0), *[int(v) for v in torch.randint(0, max(n - width, 1), (n_windows,), generator=g)]})
    checked = 0
    for s0 in starts:
        r0, r1 = lo + s0, min(lo + s0 + width, hi)
        cval, cidx = canon.knn_cosine_topk(xh, k, 0, r0, r1)
        gi, gv = idx[s0:s0 + (r1 - r0)].cpu().numpy(), val[s0:s0 + (r1 - r0)].cpu().numpy()
        if not np.array_equal(gi, cidx) or not np.array_equal(gv.view(np.uint32), cval.view(np.uint32)):
            parity_fail("knn", f"rows [{r0},{r1}) of a {xh.shape} table differ from the oracle")
        checked += r1 - r0
    return dict(rows_checked=checked, ids="bit-exact", values="bit-exact", oracle="oracle/canon.c")


def profile_traffic(key):
    """DRAM bytes per launch of the named kernel from the ncu --set full captures of this round
    (profiles/r02_traffic.json, written by scripts/write_traffic_json.py together with the commit it was taken at)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get(key)
    return (e.get("dram_bytes") if e else None), dict(file="profiles/r02_traffic.json", commit=d.get("commit"), entry=key)


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    """nvidia-smi in loop mode, started BEFORE the warm-up (its own start-up takes longer than a short timed region);
    every line is time-stamped on arrival and `stop(t0, t1)` reports the samples that fall inside the timed window
    (all samples taken under load since the start if the window is shorter than two sampling periods)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1)]
        window = "timed region"
        if len(rows) < 2:
            rows, window = [r for (_, r) in self.rows], "warm-up + timed region (timed region shorter than two samples)"
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


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
    e2e_chunks = shard_eval_job(job, lo, hi, n_chunks=2 if (hi - lo) >= 50_000 else 1)
    n_local = hi - lo
    ks = [5, 10, 20]

    # ---- device-resident leg -----------------------------------------------------------------
    ue, ie = batch.user_emb.to(dev), batch.item_emb.to(dev)
    mask = MaskCSR(batch.mask_rowptr.to(dev), batch.mask_col.to(dev))
    pos_rp, pos_col = batch.pos_rowptr.to(dev), batch.pos_col.to(dev)
    lib = _lib.load()
    hits = torch.empty(n_local, dtype=torch.int64, device=dev)
    l2buf = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    # (aborts on any mismatch -- on EVERY rank together, so that no rank is left waiting in a collective)
    parity, st = collective_gate(lambda: gate_eval(job, lo, hi, dev, ue, ie, mask, world), world, dev, "eval")   # also: engine, fallback rows

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

    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(args.warmup):
        one_step(False)
    barrier()
    t_w0 = time.monotonic()
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop(t_w0, time.monotonic())
    t_total = sum(e0.elapsed_time(e2) for e0, _, e2 in evs) * 1e-3
    t_rank_kernel = sum(e0.elapsed_time(e1) for e0, e1, _ in evs) * 1e-3 / args.steps
    tt = torch.tensor([t_total, t_rank_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_rank_kernel = float(tt[0]), float(tt[1])
    value = U * args.steps / t_total

    # ---- end-to-end leg (host buffers in, metrics out) ------------------------------------------
    n_pos_all = np.full(U, job["pos"].shape[1], dtype=np.int32)

    shard_sizes = [cdist.shard_range(U, world, r)[1] - cdist.shard_range(U, world, r)[0] for r in range(world)]

    def e2e_step():
        h = evaluate_host_batch(e2e_chunks, K, dev, copy=False)   # pinned host tensor, synchronised (consumed before the next call)
        if world > 1:
            all_h = cdist.gather_to_rank0(h.to(dev), sizes=shard_sizes)
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
        traffic, traffic_src = profile_traffic("tc_eval_kernel_main") if (world == 1 and args.scale == 1.0) else (None, None)
        roof = dict(kernel="tc_eval_kernel<false> (tcgen05 fp16 pre-filter); `achieved` divides by the time of the WHOLE ranking "
                           "chain: fp16 conversion, seed pass, main pass, fp32 re-score + certification, exact fallback launch",
                    bound="tensor", achieved=achieved,
                    peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops"],
                    frac_of_sustained_peak=(achieved / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
                    traffic=traffic, traffic_source=traffic_src,
                    tmem_read_tbps_measured=probe_tmem_read(dev),
                    peak_source=peaks["source"] + " (bf16 dense burst; fp16 and bf16 tcgen05 rates are equal)", algorithmic_flops_per_launch=flops,
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
        roofline=roof, engine=st.get("engine"), parity=dict(eval=parity),
        gpu_launches=(10 if tensor else 2) * args.steps,   # absmax, to_f16 x2, tc_eval seed + main, rescore_lists, rank_topk x3 + slice_merge (fallback rows), topk_hits
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


def build_full_graph(tu, ti, U, I, dev):
    """The whole symmetric normalized adjacency on one GPU (the N = 1 leg, and the reference rows of the N > 1 gates)."""
    import chaorec_b200 as cb
    from chaorec_b200 import _lib
    n = U + I
    graph = cb.csr_from_coo(torch.cat((tu, ti + U)), torch.cat((ti + U, tu)), None, n, n)
    lib = _lib.load()
    dis = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.check(lib.chaorec_degree_inv_sqrt(_lib.ptr(graph.rowptr), n, 0.0, _lib.ptr(dis), _lib.stream_ptr(dev)), "deg")
    graph.row_scale = graph.col_scale = dis
    graph.symmetric = True
    return graph, dis


def run_propagate_leg(args, dev, rank, world, peaks):
    import chaorec_b200 as cb
    if world > 1:
        return run_propagate_leg_sharded(args, dev, rank, world, peaks)
    cfg = PROP_CFG
    U, I, tu, ti = gpu_zipf_graph(cfg, args.scale, dev)
    n, d, L = U + I, cfg["d"], cfg["L"]
    E = int(tu.numel())
    graph, dis = build_full_graph(tu, ti, U, I, dev)
    del tu, ti
    x0 = torch.randn(n, d, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 0.1
    # ---- parity gate: the whole table against an fp64 CSR product chain (all rows, not a sample)
    res = cb.propagate_mean(graph, x0, L)
    ref, adj64 = fp64_reference_chain(graph, dis, x0, L)
    err = gate_propagate(res, ref, "propagate")
    parity = dict(rows_checked=n, rel_err=err, tol=1e-5, reference="torch.sparse CSR fp64 chain on the device (cuSPARSE), all rows")
    del res, ref
    # ---- informal library column: the same layer through torch.sparse.mm (fp32 CSR, cuSPARSE) on this GPU
    lib_ms = None
    try:
        adj32 = torch.sparse_csr_tensor(adj64.crow_indices(), adj64.col_indices(), adj64.values().float(), size=adj64.shape)
        del adj64
        for _ in range(2):
            torch.sparse.mm(adj32, x0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch.sparse.mm(adj32, x0)
        e1.record()
        torch.cuda.synchronize()
        lib_ms = e0.elapsed_time(e1) / 3
        # ---- CPU baseline leg (BASELINE.md plan of record): ONE layer of torch.sparse.mm on the host cores, COO as the
        # reference builds it (Model/FREEDOM.py:73-109) and CSR
        cpu_leg = None
        if (rank == 0 and args.scale == 1.0 and not args.skip_cpu_legs) or args.cpu_legs:
            coo = adj32.to_sparse_coo()
            coo = torch.sparse_coo_tensor(coo.indices().cpu(), coo.values().cpu(), tuple(coo.shape), is_coalesced=True)
            xc = x0.cpu()
            t0 = time.perf_counter(); torch.sparse.mm(coo, xc); t_coo = time.perf_counter() - t0
            csr = coo.to_sparse_csr()
            t0 = time.perf_counter(); torch.sparse.mm(csr, xc); t_csr = time.perf_counter() - t0
            cpu_leg = dict(coo_ms_per_layer=t_coo * 1e3, csr_ms_per_layer=t_csr * 1e3, cores=torch.get_num_threads(), kind="reference operator",
                           sample="one layer (of 3) of torch.sparse.mm(adj, x) on the host, same graph and table")
            del coo, csr, xc
        del adj32
    except Exception as exc:       # the library column is informal: never fail the bench for it
        cpu_leg = dict(error=f"{type(exc).__name__}: {exc}"[:200])
    torch.cuda.empty_cache()
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
    prop_traffic, prop_traffic_src = profile_traffic("spmm_layer") if args.scale == 1.0 else (None, None)
    return dict(metric="GCN propagate edges/s", value=nnz * L / t, unit="edges/s", ms_per_step=t * 1e3,
                config=dict(workload="3-layer normalized-adjacency propagation + layer mean (BASELINE configs[4] shape)",
                            users=U, items=I, edges=E, nnz_sym=nnz, d=d, layers=L, hub_rows=graph.n_hub,
                            hub_chunks=graph.n_hub_chunks, l2="inputs (2.7 GB table, 0.8 GB indices) exceed L2", scale=args.scale),
                roofline=dict(kernel="spmm_rows_kernel<16,1,64,20> (+hub kernels)", bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"],
                              unit="GB/s", frac=achieved / peaks["hbm_gbs"],
                              # traffic: ncu dram bytes of one layer's kernels (rows + hub-chunk + hub-finish) at this workload
                              traffic=prop_traffic, traffic_source=prop_traffic_src, peak_source=peaks["source"],
                              algorithmic_bytes_per_layer=bytes_layer),
                parity=parity, library_torch_sparse_mm_csr_ms_per_layer=lib_ms, cpu_baseline=cpu_leg,
                gpu_launches=(1 + (2 if graph.n_hub else 0)) * L * args.steps)


def run_propagate_leg_sharded(args, dev, rank, world, peaks):
    """N > 1: users stay on their rank, item partials are all-reduced once per layer (dist.BipartiteShard);
    the generic row-cyclic + all-gather variant (any adjacency) is timed beside it."""
    import torch.distributed as dist
    import chaorec_b200 as cb
    from chaorec_b200 import _lib, dist as cdist
    cfg = PROP_CFG
    U, I, tu, ti = gpu_zipf_graph(cfg, args.scale, dev)          # same seed on every rank -> (should be) the same graph
    n, d, L = U + I, cfg["d"], cfg["L"]
    # Every rank must work on the SAME graph and the same layer-0 table.  They are generated per rank from the same seeds;
    # whether that really gave identical tensors is recorded (`inputs_identical_across_ranks`), and rank 0's copies are
    # broadcast, so that the gates below cannot be tripped by an input that differs between GPUs.
    chk = torch.stack((torch.tensor(tu.numel(), device=dev), (tu * 31 + ti).sum(), (tu ^ ti).sum())).to(torch.int64)
    all_chk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    graphs_identical = all(bool(torch.equal(c, all_chk[0])) for c in all_chk)
    e_n = all_chk[0][0:1].clone()
    E = int(e_n)
    if rank != 0 and int(tu.numel()) != E:
        tu, ti = torch.empty(E, dtype=torch.int64, device=dev), torch.empty(E, dtype=torch.int64, device=dev)
    tu, ti = tu.contiguous(), ti.contiguous()
    dist.broadcast(tu, 0)
    dist.broadcast(ti, 0)

    def same_on_all_ranks(t):
        """True when the tensor generated on this rank equals rank 0's; the tensor is then replaced by rank 0's copy."""
        mine = t.clone()
        dist.broadcast(t, 0)
        flag = torch.tensor([1 if torch.equal(mine, t) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag))

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
    x0_identical = same_on_all_ranks(x0)
    xu0 = x0[shard.u_lo:shard.u_hi].clone()      # (copies, not views: .contiguous() of a row slice is the slice itself)
    xi0 = x0[U:].clone()
    # ---- parity gates at N > 1 are COLLECTIVE: every rank measures its own rows against the single-GPU product of the
    # same graph, the worst error is all-reduced, and all ranks take the same decision.  The headline variant must pass
    # (else no bench line); a secondary variant that fails is reported as failed and gets no number.
    # (the symmetric buffers of the fused exchange are allocated BEFORE the reference rows are computed, and the reference
    # rows are recomputed for every comparison: the gate must not depend on a 2.7 GB tensor surviving the variants' runs)
    err_f, t_fused, fused_note = None, None, None
    try:
        ex = cdist.FusedItemExchange(I, d, dev)
    except Exception as exc:
        ex, fused_note = None, f"fused exchange unavailable: {type(exc).__name__}: {exc}"[:300]
    full_graph, _ = build_full_graph(tu, ti, U, I, dev)

    def fresh_ref():
        torch.cuda.synchronize()
        r_ = cb.propagate_mean(full_graph, x0, L)
        torch.cuda.synchronize()
        return r_

    ref = fresh_ref()
    ref_scale = float(ref.abs().max())
    del ref

    def shard_err(pairs):
        e = max(float((got.double() - want.double()).abs().max()) / ref_scale for got, want in pairs)
        t_ = torch.tensor([e], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_[0])

    keep_fused = {}

    def nccl_variant():
        au, ai = cdist.propagate_mean_bipartite(shard, xu0, xi0, L)
        if "u" in keep_fused:      # (diagnostic: a failing gate with agreeing variants points at the reference rows)
            keep_fused["agree"] = shard_err([(au, keep_fused.pop("u")), (ai, keep_fused.pop("i"))])
        ref_ = fresh_ref()
        return shard_err([(au, ref_[shard.u_lo:shard.u_hi]), (ai, ref_[U:])])

    # ---- headline: the exchange fused into the kernels (peer stores from the SpMM epilogue, owner-side reduce + store
    # into every table, device-side barriers); falls back to the NCCL form where symmetric memory is unavailable
    if ex is not None:
        fu, fi = cdist.propagate_mean_bipartite_fused(shard, ex, xu0, xi0, L)
        torch.cuda.synchronize()
        ref_ = fresh_ref()
        err_f = shard_err([(fu, ref_[shard.u_lo:shard.u_hi]), (fi, ref_[U:])])
        del ref_
        if not err_f <= 1e-5:
            keep_fused["u"], keep_fused["i"] = fu.clone(), fi.clone()
        del fu, fi
        if err_f <= 1e-5:
            t_fused = timed(lambda: cdist.propagate_mean_bipartite_fused(shard, ex, xu0, xi0, L))
        else:
            fused_note = f"fused exchange FAILED its parity gate (max|delta|/max|ref| over all ranks = {err_f:.3e} > 1e-5): no number for it"
    # ---- NCCL all-reduce form of the same sharding
    err_b = nccl_variant()
    nccl_note = None
    if not err_b <= 1e-5:
        retry = nccl_variant()
        nccl_note = f"gate failed ({err_b:.3e}); second attempt {retry:.3e}"
        err_b = max(err_b, retry)
    t_nccl = timed(lambda: cdist.propagate_mean_bipartite(shard, xu0, xi0, L)) if err_b <= 1e-5 else None
    if t_fused is None and t_nccl is None:
        parity_fail("sharded propagate", f"no bipartite variant passed its gate: fused {err_f}, NCCL {nccl_note}, "
                                         f"fused vs NCCL {keep_fused.get('agree')}")
    t = t_fused if t_fused is not None else t_nccl
    del shard, xu0, xi0, x0, ex
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
    x0_nodes = torch.randn(n, d, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 0.1
    x0_identical = same_on_all_ranks(x0_nodes) and x0_identical
    x0 = part.permute_rows(x0_nodes)       # (padding rows of the permuted numbering are zero: they have no edges)
    # ---- parity gate: this rank's block against the single-GPU product (node order -> permuted order)
    ref = part.permute_rows(cb.propagate_mean(full_graph, x0_nodes, L))
    del x0_nodes
    blk = cdist.propagate_mean_sharded(graph, part, x0, L)
    ref_scale = float(ref.abs().max())
    err_c = shard_err([(blk, ref[rank * part.block:(rank + 1) * part.block])])
    del ref, blk, full_graph
    parity = dict(bipartite_fused_rel_err=err_f, bipartite_nccl_rel_err=err_b, bipartite_nccl_note=nccl_note, row_cyclic_nccl_rel_err=err_c,
                  fused_vs_nccl_rel_err=keep_fused.get("agree"),
                  inputs_identical_across_ranks=dict(graph=graphs_identical, layer0_table=x0_identical,
                                                     note="generated per rank from the same seeds; rank 0's copies are broadcast and used"),
                  tol=1e-5, rows_checked="every rank: all rows of its shard; worst error over ranks",
                  reference="single-GPU propagate_mean of the same graph (itself gated against the fp64 chain at N=1)")
    torch.cuda.empty_cache()
    t_cyclic = timed(lambda: cdist.propagate_mean_sharded(graph, part, x0, L)) if err_c <= 1e-5 else None
    nnz = 2 * E
    bytes_layer = nnz * (4 + 4 * d) + n * (4 * d + 8)
    achieved = bytes_layer * L / t / 1e9 / world
    return dict(metric="GCN propagate edges/s", value=nnz * L / t, unit="edges/s", ms_per_step=t * 1e3,
                exchange=("fused: SpMM epilogue stores item partials into the owners' inboxes over NVLink, owners reduce and store into every "
                          "rank's table (no NCCL call per layer)" if t_fused is not None else "NCCL all-reduce of the item partials"),
                fused_note=fused_note,
                bipartite_nccl_allreduce=(dict(value=nnz * L / t_nccl, unit="edges/s", ms_per_step=t_nccl * 1e3) if t_nccl else "no number: parity gate failed"),
                generic_row_cyclic_allgather=(dict(value=nnz * L / t_cyclic, unit="edges/s", ms_per_step=t_cyclic * 1e3) if t_cyclic else "no number: parity gate failed"),
                config=dict(workload="3-layer propagation, users sharded and kept local, item partials [I, d] exchanged once per layer",
                            users=U, items=I, edges=E, nnz_sym=nnz, d=d, layers=L, world=world, scale=args.scale,
                            exchanged_bytes_per_layer=I * d * 4),
                roofline=dict(kernel="spmm_rows_kernel<16,1,64,20> (+hub kernels)", bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"],
                              unit="GB/s per GPU", frac=achieved / peaks["hbm_gbs"], traffic=None, peak_source=peaks["source"]),
                parity=parity, gpu_launches=None)


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
        val, idx = cb.knn_topk(x, cfg["k"], row_range=(lo, hi), stats=kst)
        parity = collective_gate(lambda: gate_knn(x, cfg["k"], lo, hi, val, idx), world, dev, "knn")   # aborts on any mismatch
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
        cpu_ms = cpu_b_ms = lib_ms = None
        if world == 1 and rank == 0 and not args.skip_cpu_legs:
            # the reference's own forms on the host cores: conv A (Model/FREEDOM.py:111-128: normalise, dense sim, topk)
            # for both modalities, conv B (Model/LATTICE.py:46-75: build_sim, topk, scatter, symmetric Laplacian) for t_feat
            from oracle import ref_ops
            xc = x.cpu()
            t0 = time.perf_counter()
            ref_ops.knn_conv_a(xc, cfg["k"])
            cpu_ms = (time.perf_counter() - t0) * 1e3
            if name == "t_feat":
                t0 = time.perf_counter()
                ref_ops.knn_conv_b(xc, cfg["k"])
                cpu_b_ms = (time.perf_counter() - t0) * 1e3
            # informal library column: the same reference form with torch's CUDA kernels (cuBLAS fp32 GEMM + topk)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it_ in range(2):
                if it_ == 1:
                    e0.record()
                xn = x / torch.norm(x, p=2, dim=-1, keepdim=True)
                torch.topk(torch.mm(xn, xn.t()), cfg["k"], dim=-1)
            e1.record()
            torch.cuda.synchronize()
            lib_ms = e0.elapsed_time(e1)
            del xn
        out[name] = dict(value=N / t, unit="item rows/s", ms=t * 1e3, cpu_reference_ms=cpu_ms, cpu_reference_conv_b_ms=cpu_b_ms,
                         library_torch_cuda_mm_topk_ms=lib_ms, cpu_cores=torch.get_num_threads(),
                         D=D, N=N, k=cfg["k"], tflops=flops / t / 1e12,
                         frac_of_bf16_tensor_peak=flops / t / 1e12 / peaks["bf16_tflops"], engine=kst.get("engine"),
                         fallback_rows=kst.get("fallback_rows"), parity=parity)
    return dict(metric="kNN graph item rows/s", config=dict(workload="cosine kNN graph, clothing-shaped (BASELINE configs[2])"), **out)


SEAM_CFGS = [dict(name="C1 baby-shaped", U=19_445, I=7_050, E=160_000, d=64, L=3, seed=20260),
             dict(name="C2 sports-shaped", U=35_598, I=18_357, E=296_337, d=64, L=3, seed=20261)]


def run_seam_leg(args, dev, rank, world, peaks):
    """The drop-in seam exactly as INTEGRATION.md §1 wires it, on the reference's own small configs (BASELINE configs[0]
    and [1]): python `user_item_dict`, object-array val/test rows, LightGCN forward, then
        rank = types.MethodType(cb.gene_ranklist, model)(topk=50); cb.evaluate(model, val, rank, ks); cb.evaluate(model, test, rank, ks)
    (train_and_evaluate.py:656-659) -- cold (first call: container conversion, H2D, workspace allocation) and warm
    (conversions cached per object), beside the oracle port of the reference doing the same on the host cores.  Parity gate:
    every metric identical (==) and the ranklists equal to the reference's stable-sorted form."""
    import types
    import chaorec_b200 as cb
    from chaorec_b200 import ranking
    from oracle import ref_ops
    out = {}
    ks = [5, 10, 20]
    for cfg in SEAM_CFGS:
        U, I, d, L = int(cfg["U"] * args.scale), int(cfg["I"] * args.scale), cfg["d"], cfg["L"]
        tu, ti = ref_ops.synth_interactions(U, I, int(cfg["E"] * args.scale), cfg["seed"])
        held = ref_ops.synth_holdout(U, I, tu, ti, 4, cfg["seed"] + 1)
        _, uid, val = ref_ops.to_reference_formats(U, tu, ti, held[:, :2])
        _, _, test = ref_ops.to_reference_formats(U, tu, ti, held[:, 2:])
        ei_cpu = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy())
        uw, iw = ref_ops.xavier_embeddings(U, d, 1), ref_ops.xavier_embeddings(I, d, 2)
        # ---- reference port on the host cores: forward + gene_ranklist + evaluate x2
        t0 = time.perf_counter()
        res_cpu = ref_ops.lightgcn_forward(uw, iw, ei_cpu, L)
        t_fwd_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        rank_cpu = ref_ops.gene_ranklist(res_cpu, U, I, uid, topk=50)
        mv = ref_ops.gene_metrics(val, rank_cpu, ks)
        mt = ref_ops.gene_metrics(test, rank_cpu, ks)
        t_eval_cpu = time.perf_counter() - t0
        # ---- ours through the seam
        uwd, iwd, ei = uw.to(dev), iw.to(dev), ei_cpu.to(dev)
        cb.lightgcn_forward(uwd, iwd, ei, L)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = cb.lightgcn_forward(uwd, iwd, ei, L)
        torch.cuda.synchronize()
        t_fwd = time.perf_counter() - t0
        same_fwd = float((res.cpu() - res_cpu).abs().max() / res_cpu.abs().max())
        if not same_fwd <= 1e-5:
            parity_fail("seam forward " + cfg["name"], f"rel err {same_fwd}")
        # the evaluation seam is fed the SAME embedding bits as the reference port (its own forward output), so that
        # the two ranklists are comparable item by item
        model = types.SimpleNamespace(result=res_cpu.to(dev), num_user=U, num_item=I, user_item_dict=uid)
        model.gene_ranklist = types.MethodType(cb.gene_ranklist, model)
        times = []
        for rep in range(3):
            if rep == 0:
                ranking.clear_caches()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rk = model.gene_ranklist(topk=50)
            gv = cb.evaluate(model, val, rk, ks)
            gt = cb.evaluate(model, test, rk, ks)
            times.append(time.perf_counter() - t0)
        # ---- parity: (1) ranklist == the oracle's canonical one, bit for bit; (2) metrics == the reference loop on it;
        # (3) reported, not gated: equality with the torch.matmul/topk port (whose summation order is the BLAS build's)
        from oracle.canon_py import Canon
        rp = np.zeros(U + 1, dtype=np.int64)
        np.cumsum(np.bincount(tu.numpy(), minlength=U), out=rp[1:])
        cidx, _ = Canon().score_mask_topk(res_cpu[:U].numpy(), res_cpu[U:].numpy(), rp, ti.numpy().astype(np.int32), 1e-6, 50, id_offset=U)
        if not np.array_equal(rk.numpy(), cidx):
            parity_fail("seam " + cfg["name"], "gene_ranklist differs from the oracle's canonical ranklist")
        cv, ct = ref_ops.gene_metrics(val, torch.from_numpy(cidx), ks), ref_ops.gene_metrics(test, torch.from_numpy(cidx), ks)
        if any(gv[k][n] != cv[k][n] or gt[k][n] != ct[k][n] for k in ks for n in gv[k]):
            parity_fail("seam " + cfg["name"], f"metrics differ: {gv} / {cv}")
        port_equal = all(gv[k][n] == mv[k][n] and gt[k][n] == mt[k][n] for k in ks for n in gv[k])
        rows_equal = int((rk == rank_cpu).all(dim=1).sum())
        out[cfg["name"]] = dict(users=U, items=I, interactions=int(tu.numel()), topk=50,
                                ours_forward_ms=t_fwd * 1e3, ours_eval_cold_ms=times[0] * 1e3, ours_eval_warm_ms=min(times[1:]) * 1e3,
                                reference_cpu_forward_ms=t_fwd_cpu * 1e3, reference_cpu_eval_ms=t_eval_cpu * 1e3, cpu_cores=torch.get_num_threads(),
                                users_per_s_cold=U / times[0], users_per_s_warm=U / min(times[1:]), users_per_s_reference_cpu=U / t_eval_cpu,
                                parity=dict(ranklist="bit-exact vs oracle/canon.c (top-50, all users)", metrics="identical (==) to the reference loop on that ranklist, val and test, k=5,10,20",
                                            forward_rel_err=same_fwd, metrics_identical_to_torch_matmul_port=port_equal,
                                            ranklist_rows_identical_to_torch_topk_port=f"{rows_equal}/{U}"))
    return dict(metric="drop-in seam: gene_ranklist + evaluate x2 through the reference's python containers", **out)


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
    def secondary(name, fn):
        """A secondary leg whose parity gate fails gets NO number (the failure is reported in its place) but does not
        take the headline line down with it; run on its own (`--legs propagate`) it still aborts.  Every gate inside the
        legs is collective, so all ranks take this branch together."""
        torch.cuda.empty_cache()
        try:
            line[name] = fn(args, dev, rank, world, peaks)
        except ParityError as exc:
            if "eval" not in legs:
                raise
            line[name] = dict(parity_gate_failed=str(exc), value=None)
            torch.cuda.empty_cache()
        except Exception as exc:      # (a broken secondary leg must not cost the headline line; reported, no number)
            if "eval" not in legs:
                raise
            import traceback
            traceback.print_exc(file=sys.stderr)
            line[name] = dict(error=f"{type(exc).__name__}: {exc}"[:400], value=None)
            torch.cuda.empty_cache()

    if "propagate" in legs:
        secondary("propagate", run_propagate_leg)
    if "knn" in legs:
        secondary("knn", run_knn_leg)
    if "train" in legs:
        secondary("train", run_train_leg)
    if "seam" in legs and world == 1:
        secondary("seam", run_seam_leg)
    if rank == 0 and world == 1 and job is not None and not args.skip_cpu_legs:
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
        # (the same keys and values as the other arm's `config`: both describe the one job)
        config=dict(workload="full-rank top-20 eval, 1M users x 200k items, d=64 (BASELINE configs[3])", users=job["U"],
                    items=job["I"], d=job["d"], topk=job["K"], mask_items_per_user=EVAL_CFG["mask_per_user"],
                    sharding=f"users/{args.gpus}", l2="flushed between steps (192 MiB write)", scale=args.scale),
        cpu_baseline=dict(value=v, unit="users/s", cores=torch.get_num_threads(), kind="port",
                          sample=f"each step = {n} seeded users of the job against all items (batched reference form + gene_metrics)"),
        e2e=dict(value=v, unit="users/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
