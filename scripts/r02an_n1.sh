#!/bin/bash
# bf16 aggregation kernels with the 64-register bound (scripts/agg_bf16_bench.py), bit-exactness of the bf16 suite
cd "$(dirname "$0")/.."
python scripts/agg_bf16_bench.py > gpurun_out/r02an_agg.log 2>&1; cat gpurun_out/r02an_agg.log
python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -q -x -k "bf16 or aggregate" 2>&1 | tail -3
