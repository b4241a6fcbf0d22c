#!/bin/bash
# Per-kernel times of the 1M x 200k evaluation for the product build and the diagnostic builds under chaorec_b200/dbg/
# (scripts/build_eval_debug.sh): where the main pass's time goes.  Arguments: variant names (default: all built).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
names=${@:-$(ls chaorec_b200/dbg | sed 's/libchaorec_b200_//; s/\.so//')}
for v in product $names; do
  if [ $v = product ]; then unset CHAOREC_LIB_PATH; else export CHAOREC_LIB_PATH=$PWD/chaorec_b200/dbg/libchaorec_b200_$v.so; fi
  DEBUG=0 bash scripts/gpu_eval_sweep.sh 512 2>&1 | sed "s/^/$v: /"
  cp gpurun_out/sweep_512.csv gpurun_out/floor_$v.csv
done
