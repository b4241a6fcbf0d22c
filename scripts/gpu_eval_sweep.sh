#!/bin/bash
# Per-kernel times of the 1M x 200k evaluation for a few seed-sample sizes (env CHAOREC_TC_SEED_TILES).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/evalsweep.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import chaorec_b200 as cb
dev = torch.device('cuda', 0)
U, I, d, K = 1_000_000, 200_000, 64, 20
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(U, d, device=dev, generator=g) * 0.1
it = torch.randn(I, d, device=dev, generator=g) * 0.1
mu = torch.arange(U, device=dev).repeat_interleave(16)
import os
U = int(os.environ.get('EVAL_U', U))
u = u[:U]
mu = torch.arange(U, device=dev).repeat_interleave(16)
mi = torch.randint(0, I, (U * 16,), device=dev, generator=g)
mask = cb.mask_csr_from_edges(mu, mi, U, I)
for rep in range(2):
    st = {}
    cb.rank_topk(u, it, mask, K, 1e-6, U, stats=st)
print(st)
PY
for st in ${@:-256 512}; do
  (NCU=1 CHAOREC_TC_SEED_TILES=$st CHAOREC_TC_DEBUG=${DEBUG:-0} timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sweep_$st.csv python /tmp/evalsweep.py) > gpurun_out/sweep_$st.log 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/sweep_$st.csv')) if len(r) > 5 and r[0].isdigit()]
print('seed tiles $st:', [(r[4].split('::')[-1][:22], round(float(r[-1])/1e6, 2)) for r in rows[-6:]], open('gpurun_out/sweep_$st.log').read().strip().splitlines()[-1][:120])
PY
done
