#!/bin/bash
# Row-kernel variants (CTA size / register cap / hub threshold) at the 100M-edge propagation config.
cd "$(dirname "$0")/.."
run() { echo "variant $1 thr ${2:-default}: $(CHAOREC_SPMM_VARIANT=$1 CHAOREC_SPMM_HUB_THR=$2 python bench.py --legs propagate --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())['propagate']; print(round(d['value']/1e9,2), 'G edges/s', round(d['ms_per_step'],2), 'ms')")"; }
run 0; run 2; run 1; run 3; run 4; run 5; run 6; run 2 256; run 2 128; run 2 4096
