mkdir -p gpurun_out
# quick isolated check of the tensor engine first (a hang here must not eat the budget)
(timeout 120 python - <<'PY'
import torch, numpy as np, sys, time
sys.path.insert(0, '.')
import chaorec_b200 as cb
from oracle import ref_ops
from oracle.canon_py import Canon
canon = Canon()
dev = torch.device('cuda', 0)
U, I, d, K = 2000, 4096, 64, 20
u = ref_ops.normal_embeddings(U, d, 1); it = ref_ops.normal_embeddings(I, d, 2)
st = {}
t0 = time.time()
idx, sc = cb.rank_topk(u.to(dev), it.to(dev), None, K, 1e-6, 0, return_scores=True, stats=st)
torch.cuda.synchronize()
print('engine stats', st, 'time', time.time() - t0)
cidx, csc = canon.score_mask_topk(u.numpy(), it.numpy(), None, None, 1e-6, K, 0)
same = np.array_equal(idx.cpu().numpy(), cidx)
print('bit exact:', same, 'rows differing:', int((idx.cpu().numpy() != cidx).any(1).sum()))
PY
) > gpurun_out/tc_quick.log 2>&1; echo "tc_quick rc=$?" >> gpurun_out/tc_quick.log
tail -5 gpurun_out/tc_quick.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
(timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/bench3.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench3.log
tail -3 gpurun_out/bench3.log
