mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
(timeout 300 python __graft_entry__.py --smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
KREG='regex:rank_topk|spmm_|topk_hits|row_norm|row_scale|fp32_fma'
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1) > gpurun_out/ncu_launches.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rank_topk -s 1 -c 1 -o gpurun_out/prof_rank_r01 python bench.py --legs eval --steps 1 --warmup 1) > gpurun_out/ncu_rank.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 3 -c 1 -o gpurun_out/prof_spmm_r01 python bench.py --legs propagate --steps 1 --warmup 1) > gpurun_out/ncu_spmm.log 2>&1
tail -4 gpurun_out/smoke.log; tail -8 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu_rank.log; tail -3 gpurun_out/ncu_spmm.log; ls -la gpurun_out
