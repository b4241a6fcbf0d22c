#!/bin/bash
# Quick GPU check of the evaluation engine: a tiny case first (a barrier-protocol mistake hangs -> short timeout),
# then the ranking parity tests, then the 1M x 200k timing with a per-kernel launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/tiny.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import chaorec_b200 as cb
from oracle import ref_ops
from oracle.canon_py import Canon
canon = Canon()
dev = torch.device('cuda', 0)
print('imports done', flush=True)
for (U, I, d, K) in ((1500, 70000, 64, 20), (700, 2500, 64, 20), (3000, 5000, 64, 20), (500, 4100, 32, 10), (20000, 200000, 64, 20)):
    u = ref_ops.normal_embeddings(U, d, 100 + U); it = ref_ops.normal_embeddings(I, d, 200 + I)
    tu, ti = ref_ops.synth_interactions(U, I, 8 * U, 3)
    mask = cb.mask_csr_from_edges(tu.to(dev), ti.to(dev), U, I)
    st = {}
    idx, sc = cb.rank_topk(u.to(dev), it.to(dev), mask, K, 1e-6, 0, return_scores=True, stats=st)
    torch.cuda.synchronize()
    cidx, csc = canon.score_mask_topk(u.numpy(), it.numpy(), mask.rowptr.cpu().numpy(), mask.col.cpu().numpy(), 1e-6, K, 0)
    ok = np.array_equal(idx.cpu().numpy(), cidx) and np.array_equal(sc.cpu().numpy().view(np.uint32), csc.view(np.uint32))
    print((U, I, d, K), st, 'MATCH' if ok else 'MISMATCH', flush=True)
    if not ok:
        gi, gs = idx.cpu().numpy(), sc.cpu().numpy()
        bad = np.nonzero((gi != cidx).any(1))[0]
        print('  bad rows', len(bad), bad[:10])
        r = bad[0]
        print('  row', r, 'gpu', gi[r].tolist(), '\n  canon', cidx[r].tolist())
        print('  gpu sc', gs[r][:8], 'canon sc', csc[r][:8])
        miss = [c for c in cidx[r].tolist() if c not in gi[r].tolist()]
        print('  missing cols', miss, 'tiles', [c // 64 for c in miss], 'mask', sorted(mask.col[mask.rowptr[r]:mask.rowptr[r+1]].cpu().tolist()))
PY
timeout 150 python /tmp/tiny.py > gpurun_out/tiny.log 2>&1; echo "tiny rc=$?" >> gpurun_out/tiny.log; tail -8 gpurun_out/tiny.log
if [ "$1" = "tiny" ]; then exit 0; fi
if [ "$1" != "notests" ] && [ "$1" != "ncu" ]; then
  timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider --timeout 300 -k "not exact" > gpurun_out/pytest_rank.log 2>&1; tail -6 gpurun_out/pytest_rank.log
fi
cat > /tmp/evaltime.py <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
import chaorec_b200 as cb
from oracle import ref_ops
dev = torch.device('cuda', 0)
U, I, d, K = 1_000_000, 200_000, 64, 20
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(U, d, device=dev, generator=g) * 0.1
it = torch.randn(I, d, device=dev, generator=g) * 0.1
mu = torch.arange(U, device=dev).repeat_interleave(16)
mi = torch.randint(0, I, (U * 16,), device=dev, generator=g)
mask = cb.mask_csr_from_edges(mu, mi, U, I)
import os
for rep in range(2 if os.environ.get('NCU') else 4):
    st = {}
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    idx = cb.rank_topk(u, it, mask, K, 1e-6, U, stats=None)
    e1.record(); torch.cuda.synchronize()
    print('eval 1M x 200k: %.2f ms' % e0.elapsed_time(e1), flush=True)
st = {}
cb.rank_topk(u, it, mask, K, 1e-6, U, stats=st); print(st)
PY
timeout 200 python /tmp/evaltime.py > gpurun_out/evaltime.log 2>&1; tail -6 gpurun_out/evaltime.log
if [ "$1" != "ncu" ]; then exit 0; fi
(NCU=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eval_launches.csv python /tmp/evaltime.py) > gpurun_out/eval_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/eval_launches.csv')) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-14:]:
    print(r[4][:70], r[-1])
PY
