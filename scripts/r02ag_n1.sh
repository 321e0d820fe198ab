#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_dropin_trainer.py -q -x 2>&1 | tail -12 > gpurun_out/r02ag_pytest.log


tail -3 gpurun_out/r02ag_pytest.log
