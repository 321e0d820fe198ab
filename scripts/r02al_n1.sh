#!/bin/bash
# warps per aggregation block vs row width / storage type (scripts/agg_bf16_bench.py: cfg4 graph, d = 256 and 128, fp32 and bf16)
cd "$(dirname "$0")/.."
for w in 8 4 2; do
  CB_AGG_WARPS=$w python scripts/agg_bf16_bench.py > gpurun_out/r02al_agg_w$w.log 2>&1
  echo "== $w warps per block"; cat gpurun_out/r02al_agg_w$w.log
done
