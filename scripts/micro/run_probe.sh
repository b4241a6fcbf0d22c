#!/bin/bash
# Runs scripts/micro/probe.bin on the GPU box under short timeouts.  usage: run_probe.sh [ts|mma|all]
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
OUT=gpurun_out/probe_${1:-all}.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $OUT 2>&1
if [ "${1:-all}" = "ts" ] || [ "${1:-all}" = "all" ]; then
  for m in 4 0 1 3 7; do
    timeout 60 scripts/micro/probe.bin ts $m >> $OUT 2>&1 || echo "ts mode $m FAILED rc=$?" >> $OUT
  done
fi
if [ "${1:-all}" != "ts" ]; then
  (nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 500 > gpurun_out/probe_clocks.csv) &
  SMI=$!
  if [ "${1:-all}" = "mma" ]; then
    for i in ${MMA_IDS:-0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17}; do
      timeout 20 scripts/micro/probe.bin mma $i >> $OUT 2>&1 || echo "mma variant $i FAILED rc=$?" >> $OUT
    done
  else
    timeout 120 scripts/micro/probe.bin ${1:-all} >> $OUT 2>&1 || echo "rates FAILED rc=$?" >> $OUT
  fi
  kill $SMI
  sort gpurun_out/probe_clocks.csv | uniq -c | sort -rn | head -8 >> $OUT
fi
cat $OUT
