#!/bin/bash
# Quick A/B timing of the eval rank kernel chain (303,104 users x 200k items, masks) + the bit-exactness tests.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_ranking.py -m gpu -q -x -p no:cacheprovider -k "prepass or million or rarely or seeded" 2>&1 | tail -2
cat > /tmp/evtime.py <<'PY'
import torch, sys, os
sys.path.insert(0, '.')
import chaorec_b200 as cb
from oracle import ref_ops
dev = torch.device('cuda', 0)
U, I, d, K = 303_104, 200_000, 64, 20
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(U, d, device=dev, generator=g) * 0.1
it = torch.randn(I, d, device=dev, generator=g) * 0.1
gg = torch.Generator().manual_seed(2)
cdf, perm = ref_ops.zipf_item_sampler(I, gg)
mu = torch.arange(U).repeat_interleave(16)
mi = perm[torch.searchsorted(cdf, torch.rand(mu.numel(), generator=gg, dtype=torch.float64)).clamp_(max=I - 1)]
mask = cb.mask_csr_from_edges(mu.to(dev), mi.to(dev), U, I)
for rep in range(3):
    st = {}
    for _ in range(2): cb.rank_topk(u, it, mask, K, stats=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): cb.rank_topk(u, it, mask, K, stats=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{ms:7.2f} ms per {U} users -> {U/ms/1e3:6.2f} M users/s, fallback rows {st.get('fallback_rows')}")
PY
timeout 600 python /tmp/evtime.py 2>&1 | tail -4
