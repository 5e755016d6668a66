#!/bin/bash
# usage: tools/variant_sweep.sh "0 1 2" [paths] [steps]  -- parity tests + timing for each fused-kernel variant
P=${2:-4000000}; T=${3:-200}
for v in $1; do
  echo "=== variant $v"
  SDB_FUSED_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fused" 2>&1 | tail -4
  SDB_FUSED_VARIANT=$v timeout 300 python tools/prof_run.py $P $T 5 | tail -1
done
