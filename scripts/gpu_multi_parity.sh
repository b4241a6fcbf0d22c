#!/bin/bash
# Parity records of every NCCL path on N GPUs: the sharded-propagation check (both variants) and the multi-GPU pytest.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus_$N.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/gpu_bipartite_check.py) > gpurun_out/sharded_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sharded_check_n$N.log
grep -E "sharded propagation|rc=" gpurun_out/sharded_check_n$N.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_multi_n$N.log
