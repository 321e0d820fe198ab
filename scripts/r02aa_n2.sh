#!/bin/bash
# 2 GPUs: the multi-GPU parity worker (incl. source-panel passes) and the default bench
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40 > gpurun_out/r02aa_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench_n2.json 2> gpurun_out/r02aa_bench_n2.err
tail -4 gpurun_out/r02aa_pytest.log
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02aa_bench_n2.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), d['e2e']['ms_per_step'], d['parity']['logits_checksum_initial_weights'], d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items()})
P
