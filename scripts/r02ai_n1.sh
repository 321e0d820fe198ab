#!/bin/bash
# A/B of one kernel change: scripts/_ab/libcoldbrew_b200_base.so (before) against the in-tree library (after)
cd "$(dirname "$0")/.."
for tag in base new; do
  if [ $tag = base ]; then export CB_LIB=$PWD/scripts/_ab/libcoldbrew_b200_base.so; else unset CB_LIB; fi
  python scripts/grad_epilogue_bench.py > gpurun_out/r02ai_epilogue_$tag.log 2>&1
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02ai_bench_$tag.json 2> gpurun_out/r02ai_bench_$tag.err
done
unset CB_LIB
python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -3
paste gpurun_out/r02ai_epilogue_base.log gpurun_out/r02ai_epilogue_new.log | cut -c1-150
for t in base new; do python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02ai_bench_$t.json') if l.startswith('{')][-1]
print('$t', round(d['ms_per_step'],2), d['parity']['train_nll_after_timed_steps'], {k:v['avg_ms'] for k,v in d['roofline_kernels'].items() if 'gemm' in k})
P
done
