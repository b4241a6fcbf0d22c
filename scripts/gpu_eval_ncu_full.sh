#!/bin/bash
# One `ncu --set full` capture of the main evaluation kernel at the 1M x 200k config (import-source on).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/evalonce.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import chaorec_b200 as cb
dev = torch.device('cuda', 0)
U, I, d, K = 1_000_000, 200_000, 64, 20
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(U, d, device=dev, generator=g) * 0.1
it = torch.randn(I, d, device=dev, generator=g) * 0.1
mu = torch.arange(U, device=dev).repeat_interleave(16)
mi = torch.randint(0, I, (U * 16,), device=dev, generator=g)
mask = cb.mask_csr_from_edges(mu, mi, U, I)
cb.rank_topk(u, it, mask, K, 1e-6, U)
torch.cuda.synchronize()
PY
KPAT=${1:-tc_eval_kernel}
SKIP=${2:-1}
OUT=${3:-prof_eval}
timeout 420 ncu --set full --clock-control none --import-source on -k regex:$KPAT -s $SKIP -c 1 -f -o gpurun_out/$OUT python /tmp/evalonce.py > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
ncu -i gpurun_out/$OUT.ncu-rep --page raw --csv > gpurun_out/${OUT}_raw.csv 2>/dev/null
ncu -i gpurun_out/$OUT.ncu-rep --page source --csv > gpurun_out/${OUT}_source.csv 2>/dev/null
ls -la gpurun_out/$OUT.ncu-rep
