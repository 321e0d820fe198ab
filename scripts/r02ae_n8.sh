#!/bin/bash
# 8 GPUs: the default bench (column panels = 4) and the one-launch exchange beside it
cd "$(dirname "$0")/.."
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02ae_bench_n8_$tag.json 2> gpurun_out/r02ae_bench_n8_$tag.err; }
run default
run panels1 --panels 1 --no-e2e
run panels2 --panels 2 --no-e2e
for t in default panels1 panels2; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02ae_bench_n8_$t.json') if l.startswith('{')][-1]
    print('$t', round(d['ms_per_step'],2), (d.get('e2e') or {}).get('ms_per_step'), d['impl_details']['parallelism'], d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items()})
except Exception as e:
    print('$t failed', e); print(open('gpurun_out/r02ae_bench_n8_$t.err').read()[-1500:])
P
done
