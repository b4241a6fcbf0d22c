#!/bin/bash
# Full GPU validation: every -m gpu test, smoke(), the default bench line and the hub-chunk ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1_final.log 2>&1; tail -1 gpurun_out/bench_n1_final.log > gpurun_out/bench_n1_final.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 > gpurun_out/bench_ref_final.json
(timeout 600 ncu --set full --clock-control none -k regex:spmm_hub_chunk -s 3 -c 1 -f -o gpurun_out/prof_spmm_hub python bench.py --legs propagate --steps 1 --warmup 1) > gpurun_out/ncu_spmm_hub.log 2>&1
ncu -i gpurun_out/prof_spmm_hub.ncu-rep --page raw --csv > gpurun_out/prof_spmm_hub_raw.csv 2>/dev/null
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_n1_final.json'))
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'], d['roofline']['frac'], d['clocks'])
for leg in ('propagate', 'knn', 'train'):
    if leg in d: print(leg, json.dumps(d[leg])[:600])
print(open('gpurun_out/bench_ref_final.json').read()[:600])
PY
