mkdir -p gpurun_out
(timeout 900 python bench.py --steps 5 --warmup 3 --legs eval) > gpurun_out/bench6.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench6.log
tail -2 gpurun_out/bench6.log | cut -c1-700
