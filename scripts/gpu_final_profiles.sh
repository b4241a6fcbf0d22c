#!/bin/bash
# End-of-round evidence: launch list of the default bench command + ncu --set full captures of the dominant
# kernels; raw/source pages are exported to CSV on the box (the .ncu-rep files themselves are dropped when large).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${ROUND:-r02}
KREG='regex:rank_topk|spmm_|sddmm|topk_hits|row_norm|row_scale|tc_cand|tc_eval|rescore|to_bf16|to_f16|absmax|max_reduce|slice_merge|degree_inv|gather_rows|bpr_|sample_neg|cooccur|reduce_slots'
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 900 --csv --log-file gpurun_out/launches_${R}_final.csv python bench.py --steps 2 --warmup 1 --legs eval,propagate,knn,train --skip-cpu-legs) > gpurun_out/ncu_launches_final.log 2>&1
cap() {  # name, kernel regex, skip, extra bench args...
  local name=$1 kre=$2 skip=$3; shift 3
  (timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s "$skip" -c 1 -f -o gpurun_out/prof_$name python bench.py "$@") > gpurun_out/ncu_$name.log 2>&1
  if [ -f gpurun_out/prof_$name.ncu-rep ]; then
    ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_${name}_source.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/prof_$name.ncu-rep)
    if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/prof_$name.ncu-rep; fi
  fi
}
# (bench --legs eval: the parity gate's call comes first; launch 0/1 of tc_eval_kernel = its seed and main pass)
cap ${R}_tc_eval_main tc_eval_kernel 1 --legs eval --steps 1 --warmup 1 --skip-cpu-legs
cap ${R}_rescore_lists rescore_lists 0 --legs eval --steps 1 --warmup 1 --skip-cpu-legs
cap ${R}_spmm_rows spmm_rows 3 --legs propagate --steps 1 --warmup 1 --skip-cpu-legs
cap ${R}_spmm_hub_chunk spmm_hub_chunk 3 --legs propagate --steps 1 --warmup 1 --skip-cpu-legs
ls -la gpurun_out | tail -12
