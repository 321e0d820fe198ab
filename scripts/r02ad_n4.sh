#!/bin/bash
# 4 GPUs: the multi-GPU parity worker at world 4 and the default bench
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30 > gpurun_out/r02ad_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ad_bench_n4.json 2> gpurun_out/r02ad_bench_n4.err
tail -3 gpurun_out/r02ad_pytest.log
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02ad_bench_n4.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), d['e2e']['ms_per_step'], d['impl_details']['parallelism'], d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items()})
P
