#!/bin/bash
cd "$(dirname "$0")/.."
PYTHONPATH=. python - <<'PY'
import torch, bench
dev = torch.device("cuda:0")
for w in (4, 8, 12, 16):
    t = bench.probe_tmem_read(dev, w)
    print(f"tmem read, {w:2d} warps/SM: {t:6.2f} TB/s  = {t*1e12/148/1.92e9:6.1f} B/clk/SM @1.92 GHz;  1M x 200k fp32 scores: {8e11/t/1e9:6.1f} ms")
PY
