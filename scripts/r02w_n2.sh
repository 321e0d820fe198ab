#!/bin/bash
# 2-GPU check of the source-panel passes: parity worker, then the bench with passes on / off / S=4
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_passes.py tests/test_gpu_multi.py -x -q 2>&1 | tail -25 > gpurun_out/r02w_pytest.log
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/r02w_bench_n2_$tag.json 2> gpurun_out/r02w_bench_n2_$tag.err; }
run passes2 --passes on
run off --passes off
run passes4 --passes on --src-panels 4
tail -3 gpurun_out/r02w_pytest.log
for t in passes2 off passes4; do python - <<P
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02w_bench_n2_$t.json') if l.startswith('{')][-1]
    print('$t', d['ms_per_step'], d['parity'], {k:(v['avg_ms'],v['launches_per_step']) for k,v in d['roofline_kernels'].items()})
except Exception as e:
    print('$t failed', e)
P
done
