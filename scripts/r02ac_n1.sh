#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_row_order.py -q -x 2>&1 | tail -30 > gpurun_out/r02ac_pytest.log
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02ac_bench_$tag.json 2> gpurun_out/r02ac_bench_$tag.err; }
run order CB_X=0
run plain CB_ROW_ORDER=0
tail -4 gpurun_out/r02ac_pytest.log
for t in order plain; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02ac_bench_$t.json') if l.startswith('{')][-1]
    print('$t', round(d['ms_per_step'],2), d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items() if k.startswith('agg')})
except Exception as e:
    print('$t failed', e); print(open('gpurun_out/r02ac_bench_$t.err').read()[-1500:])
P
done
