#!/bin/bash
# Row order (sorted by length vs natural) at the 100M-edge propagation config.
cd "$(dirname "$0")/.."
run() { echo "row_order $1: $(CHAOREC_SPMM_ROW_ORDER=$1 python bench.py --legs propagate --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())['propagate']; print(round(d['value']/1e9,2), 'G edges/s', round(d['ms_per_step'],2), 'ms')")"; }
run 0; run 1; run 0; run 1
