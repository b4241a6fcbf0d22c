mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_propagate.py -m gpu -q --timeout 600 -p no:cacheprovider -x) > gpurun_out/pytest_prop.log 2>&1; tail -3 gpurun_out/pytest_prop.log
KREG='regex:spmm_'
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREG" -s 9 -c 9 --csv --log-file gpurun_out/spmm_launches.csv python bench.py --steps 2 --warmup 1 --legs propagate) > gpurun_out/ncu_spmm2.log 2>&1
grep -o '"propagate": {"metric": "GCN propagate edges/s", "value": [0-9.]*' gpurun_out/ncu_spmm2.log
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/spmm_launches.csv')))
i=[k for k,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[i+1:]:
    print(r[4].split('(')[0][-32:], r[-3], r[-2], r[-1])
PY
(timeout 600 python bench.py --steps 5 --warmup 3 --legs propagate) 2>&1 | tail -1 | cut -c1-400
