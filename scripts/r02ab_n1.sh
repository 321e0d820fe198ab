#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_hub_hint.py tests/test_gpu_bulk_staging.py -q 2>&1 | tail -15 > gpurun_out/r02ab_pytest.log
run() { tag=$1; shift; python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/r02ab_bench_$tag.json 2> gpurun_out/r02ab_bench_$tag.err; }
run h0
run h24 --hub-hint-mb 24
run h48 --hub-hint-mb 48
run h80 --hub-hint-mb 80
tail -4 gpurun_out/r02ab_pytest.log
for t in h0 h24 h48 h80; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02ab_bench_$t.json') if l.startswith('{')][-1]
    print('$t', round(d['ms_per_step'],2), d['impl_details'].get('l2_hub_hint'), d['parity']['logits_checksum_initial_weights'][0], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items() if k.startswith('agg')})
except Exception as e:
    print('$t failed', e); print(open('gpurun_out/r02ab_bench_$t.err').read()[-1500:])
P
done
