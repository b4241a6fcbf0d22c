mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
(timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/bench4.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench4.log
tail -3 gpurun_out/bench4.log
KREG='regex:rank_topk|spmm_|topk_hits|row_norm|row_scale|tc_cand|rescore|to_bf16|max_reduce|slice_merge'
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 300 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 --legs eval,knn) > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_r01b.csv')))
i = [k for k, r in enumerate(rows) if r and r[0] == 'ID'][0]
agg = collections.OrderedDict()
for r in rows[i + 1:]:
    name = r[4].split('(')[0][-40:]
    agg.setdefault(name, []).append(float(r[-1]) / 1e6)
for k, v in agg.items():
    print(f"{k:42s} n={len(v):3d} ms: " + " ".join(f"{x:.3f}" for x in v[:10]))
PY
