#!/bin/bash
# one B200: whole -m gpu suite, k_agg with 1 / 2 warps per block
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02z_pytest.log
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02z_bench_$tag.json 2> gpurun_out/r02z_bench_$tag.err; }
run w2 CB_X=0
run w1 CB_AGG_WARPS=1
tail -5 gpurun_out/r02z_pytest.log
for t in w2 w1; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02z_bench_$t.json') if l.startswith('{')][-1]
    print('$t', round(d['ms_per_step'],2), d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items()})
except Exception as e:
    print('$t failed', e)
P
done
