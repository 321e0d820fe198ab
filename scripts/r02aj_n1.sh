#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_relu_mask.py tests/test_gpu_parity.py tests/test_gpu_gemm.py tests/test_gpu_bf16.py -q -x 2>&1 | tail -12 > gpurun_out/r02aj_pytest.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02aj_bench_n1.json 2> gpurun_out/r02aj_bench_n1.err
tail -3 gpurun_out/r02aj_pytest.log
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02aj_bench_n1.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), d['parity'], {k:(v['avg_ms'],v['launches_per_step']) for k,v in d['roofline_kernels'].items()})
P
tail -5 gpurun_out/r02aj_bench_n1.err
