#!/bin/bash
# one B200: whole -m gpu suite, graph-prep timings with caller workspaces, k_agg block-size A/B
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02y_pytest.log
python scripts/graph_prep_bench.py > gpurun_out/r02y_prep_bench.log 2>&1
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02y_bench_$tag.json 2> gpurun_out/r02y_bench_$tag.err; }
run w8 CB_X=0
run w4 CB_AGG_WARPS=4
run w2 CB_AGG_WARPS=2
run bulk8 CB_AGG_BULK=8
run bulk4 CB_AGG_BULK=4
tail -3 gpurun_out/r02y_pytest.log; cat gpurun_out/r02y_prep_bench.log | tail -6
for t in w8 w4 w2 bulk8 bulk4; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02y_bench_$t.json') if l.startswith('{')][-1]
    print('$t', round(d['ms_per_step'],2), d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items()})
except Exception as e:
    print('$t failed', e)
P
done
