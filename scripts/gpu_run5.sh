mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
(timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/bench5.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench5.log
tail -2 gpurun_out/bench5.log | cut -c1-900
