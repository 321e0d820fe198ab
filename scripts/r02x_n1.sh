#!/bin/bash
# one B200: graph-prep kernels (tests + timings), then A/B of the weight-gradient overlap and the x0 L2 prefetch
cd "$(dirname "$0")/.."
python -m pytest tests/test_graph_prep.py tests/test_gpu_passes.py -x -q 2>&1 | tail -25 > gpurun_out/r02x_pytest.log
python scripts/graph_prep_bench.py > gpurun_out/r02x_prep_bench.log 2>&1
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02x_bench_$tag.json 2> gpurun_out/r02x_bench_$tag.err; }
run base CB_X=0
run dw CB_DW_OVERLAP=1
run pf CB_AGG_PREFETCH=1
run both CB_DW_OVERLAP=1 CB_AGG_PREFETCH=1
tail -3 gpurun_out/r02x_pytest.log; cat gpurun_out/r02x_prep_bench.log | tail -6
for t in base dw pf both; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02x_bench_$t.json') if l.startswith('{')][-1]
    print('$t', round(d['ms_per_step'],2), d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items()})
except Exception as e:
    print('$t failed', e)
P
done
