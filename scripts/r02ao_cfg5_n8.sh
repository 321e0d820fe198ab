#!/bin/bash
# BASELINE configs[4] on 8 GPUs: 50 M nodes / 10^9 edges / 128-dim bf16, 3-layer GCN + SE
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ao_bench_cfg5_n8.json 2> gpurun_out/r02ao_bench_cfg5_n8.err
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02ao_bench_cfg5_n8.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), d['value'], (d.get('e2e') or {}).get('ms_per_step'), d['impl_details']['parallelism'], d['parity'], {k:(v['avg_ms'],v['launches_per_step']) for k,v in d['roofline_kernels'].items()})
P
tail -2 gpurun_out/r02ao_bench_cfg5_n8.err
