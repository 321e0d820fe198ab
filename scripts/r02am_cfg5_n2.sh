#!/bin/bash
cd "$(dirname "$0")/.."
for w in 8 2; do
CB_AGG_WARPS=$w timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02am_bench_cfg5_n2_w$w.json 2> gpurun_out/r02am_bench_cfg5_n2_w$w.err
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02am_bench_cfg5_n2_w$w.json') if l.startswith('{')][-1]
print($w, round(d['ms_per_step'],2), {k:(v['avg_ms'],v['launches_per_step']) for k,v in d['roofline_kernels'].items() if 'agg' in k})
P
done
