mkdir -p gpurun_out
cat > /tmp/tcfloor.py <<'PY'
import torch, sys, os
sys.path.insert(0, '.')
import chaorec_b200 as cb
dev = torch.device('cuda', 0)
U, I, d, K = 151_552, 200_000, 64, 20
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(U, d, device=dev, generator=g) * 0.1
it = torch.randn(I, d, device=dev, generator=g) * 0.1
for mode in ('0', '2', '10', '8', '0'):
    os.environ['CHAOREC_TC_DEBUG'] = mode
    cb.rank_topk(u, it, None, K)
    torch.cuda.synchronize()
PY
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_cand -c 8 --csv --log-file gpurun_out/tcfloor.csv python /tmp/tcfloor.py) > gpurun_out/tcfloor.log 2>&1
grep tc_cand gpurun_out/tcfloor.csv | awk -F'","' '{print $NF}'
