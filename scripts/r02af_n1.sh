#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -q -x -k "row_sparse_hint or fused_aggregate or backward_fusion" 2>&1 | tail -8 > gpurun_out/r02af_pytest.log
python bench.py --dropout 0.5 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/r02af_bench_n1_dropout0.5.json 2> gpurun_out/r02af_bench_n1_dropout0.5.err
tail -3 gpurun_out/r02af_pytest.log
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02af_bench_n1_dropout0.5.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), {k:(v['avg_ms'],v['launches_per_step']) for k,v in d['roofline_kernels'].items()})
P
