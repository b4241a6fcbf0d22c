mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus_$N.txt
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3) > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n$N.log
tail -4 gpurun_out/bench_n$N.log | cut -c1-3000
