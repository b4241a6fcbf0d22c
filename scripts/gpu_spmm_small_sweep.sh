#!/bin/bash
# Warm (L2-resident) SpMM time on the sports-shaped graph for several hub plans.
cd "$(dirname "$0")/.."
cat > /tmp/sweep.py <<'PY'
import torch, chaorec_b200 as cb
from oracle import ref_ops
U, I, E, d = 35598, 18357, 296337, 64
tu, ti = ref_ops.synth_interactions(U, I, E, 20262)
ei = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy()).cuda()
g = cb.graph_from_edge_index(ei, U + I)
x = torch.randn(U + I, d, device="cuda")
y = torch.empty_like(x)
from chaorec_b200.propagate import _spmm_raw
def timeit(n=200):
    for _ in range(20): _spmm_raw(g, x, y, None, 0.0, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): _spmm_raw(g, x, y, None, 0.0, 0)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
gr = torch.cuda.CUDAGraph()
def timegraph(n=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        _spmm_raw(g, x, y, None, 0.0, 0)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10): _spmm_raw(g, x, y, None, 0.0, 0)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * 10) * 1e3
for thr, ch in ((1024, 4096), (256, 1024), (128, 512), (64, 256), (64, 128), (32, 128), (32, 64), (16, 64)):
    g.plan_hubs(threshold=thr, chunk=ch)
    t_f = timeit(); tg_f = timegraph()
    g.hub_counters = None
    t_s = timeit(); tg_s = timegraph()
    print(f"thr {thr:5d} chunk {ch:5d} hubs {g.n_hub:6d} chunks {g.n_hub_chunks:6d}  fused {t_f:7.1f} us (graph {tg_f:6.1f})  3-launch {t_s:7.1f} us (graph {tg_s:6.1f})")
PY
PYTHONPATH=. python /tmp/sweep.py
