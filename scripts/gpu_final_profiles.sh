mkdir -p gpurun_out
KREG='regex:rank_topk|spmm_|topk_hits|row_norm|row_scale|tc_cand|rescore|to_bf16|max_reduce|slice_merge|degree_inv'
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 1) > gpurun_out/ncu_launches_final.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_candidates -s 1 -c 1 -o gpurun_out/prof_tc_final python bench.py --legs eval --steps 1 --warmup 1) > gpurun_out/ncu_tc_final.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 3 -c 1 -o gpurun_out/prof_spmm_final python bench.py --legs propagate --steps 1 --warmup 1) > gpurun_out/ncu_spmm_final.log 2>&1
(timeout 900 ncu --set full --clock-control none -k regex:tc_candidates -s 4 -c 1 -o gpurun_out/prof_tc_knn_final python bench.py --legs knn --steps 2 --warmup 1) > gpurun_out/ncu_tcknn_final.log 2>&1
ls -la gpurun_out | tail -8
