mkdir -p gpurun_out
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_candidates -s 1 -c 1 -o gpurun_out/prof_tc_r01 python bench.py --legs eval --steps 1 --warmup 1) > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
