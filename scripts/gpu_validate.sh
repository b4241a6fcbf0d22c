#!/bin/bash
# Full GPU validation: every -m gpu test, smoke(), the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1_final.log 2>&1; tail -1 gpurun_out/bench_n1_final.log > gpurun_out/bench_n1_final.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_n1_final.json'))
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'], d['roofline']['frac'], d['clocks'])
print(d['cpu_baseline'])
for leg in ('propagate', 'knn', 'train'):
    if leg in d: print(leg, json.dumps(d[leg])[:500])
PY
