#!/bin/bash
# Per-kernel durations of one sports-shaped LightGCN training step (launch list).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/train_step.py <<'PY'
import torch, chaorec_b200 as cb
from oracle import ref_ops
U, I, E, d, L, B = 35598, 18357, 296337, 64, 3, 1024
tu, ti = ref_ops.synth_interactions(U, I, E, 20262)
ei = ref_ops.symmetrise_edges(torch.stack((tu, ti + U), 1).numpy()).cuda()
uw = ref_ops.xavier_embeddings(U, d, 1).cuda().requires_grad_(True)
iw = ref_ops.xavier_embeddings(I, d, 2).cuda().requires_grad_(True)
opt = torch.optim.Adam([uw, iw], lr=1e-3, fused=True)
g = torch.Generator().manual_seed(5)
pick = torch.randint(0, tu.numel(), (B,), generator=g)
users, pos, neg = tu[pick].cuda(), (ti[pick] + U).cuda(), (torch.randint(0, I, (B,), generator=g) + U).cuda()
gr = cb.graph_from_edge_index(ei, U + I)
deg = (gr.rowptr[1:] - gr.rowptr[:-1])
print("max deg", int(deg.max()), "n_hub", gr.n_hub, "rows>64", int((deg > 64).sum()), "rows>256", int((deg > 256).sum()))
for i in range(4):
    opt.zero_grad(set_to_none=True)
    loss, _ = cb.lightgcn_loss(uw, iw, ei, L, users, pos, neg, 1e-4)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
PY
PYTHONPATH=. ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python /tmp/train_step.py > gpurun_out/train_step.log 2>&1
tail -3 gpurun_out/train_step.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
recs = [(r[ki][:70], float(r[vi].replace(',', '')), r[ui]) for r in rows[1:]]
n = len(recs) // 4
for k, v, u in recs[-n:]:
    print(f"{v:10.1f} {u} {k}")
PY
