import sys, torch
sys.path.insert(0, '.')
import chaorec_b200 as cb
from oracle import ref_ops
dev = torch.device('cuda', 0)
U, I, d, K = 5000, 6000, 64, 20
u = ref_ops.normal_embeddings(U, d, 71); it = ref_ops.normal_embeddings(I, d, 72)
tu, ti = ref_ops.synth_interactions(U, I, 8 * U, 73)
mask = cb.mask_csr_from_edges(tu.to(dev), ti.to(dev), U, I)
st = {}
cb.rank_topk(u.to(dev), it.to(dev), mask, K, 1e-6, 0, stats=st)
print('fallback rows', st['fallback_rows'])
