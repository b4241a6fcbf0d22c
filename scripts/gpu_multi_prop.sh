#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --legs propagate) > gpurun_out/bench_prop_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_prop_n$N.log
grep '^{' gpurun_out/bench_prop_n$N.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())['propagate']; print(d['value']/1e9, d['ms_per_step'], d['generic_row_cyclic_allgather'])"
tail -3 gpurun_out/bench_prop_n$N.log | cut -c1-400
