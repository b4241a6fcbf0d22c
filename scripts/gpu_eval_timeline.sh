#!/bin/bash
# Prints the per-tile time line of CTA 0 (library built with -DCHAOREC_EV_TIMELINE).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/tl.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import chaorec_b200 as cb
dev = torch.device('cuda', 0)
U, I, d, K = 148 * 384, 200_000, 64, 20
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(U, d, device=dev, generator=g) * 0.1
it = torch.randn(I, d, device=dev, generator=g) * 0.1
cb.rank_topk(u, it, None, K, 1e-6, U)
torch.cuda.synchronize()
PY
for dbg in 0; do
  echo "=== CHAOREC_TC_DEBUG=$dbg"
  CHAOREC_TC_DEBUG=$dbg timeout 120 python /tmp/tl.py 2>&1 | grep -A 64 "^timeline" | tail -66
done > gpurun_out/timeline.log 2>&1
cat gpurun_out/timeline.log
