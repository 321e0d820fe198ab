#!/bin/bash
# configs[4] per-GPU size on 2 GPUs (weak-scaling point; the 8-GPU line of this config was measured earlier in the round)
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ak_bench_cfg5_n2.json 2> gpurun_out/r02ak_bench_cfg5_n2.err
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02ak_bench_cfg5_n2.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), d['value'], (d.get('e2e') or {}).get('ms_per_step'), d['parity'], {k:(v['avg_ms'],v['launches_per_step']) for k,v in d['roofline_kernels'].items()})
P
tail -3 gpurun_out/r02ak_bench_cfg5_n2.err
