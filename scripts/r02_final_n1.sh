#!/bin/bash
# final single-GPU evidence of the round: default bench line, ncu launch list of the same command, ncu --set full of k_agg
cd "$(dirname "$0")/.."
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
python bench.py --dropout 0.5 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/r02_final_bench_n1_dropout0.5.json 2> gpurun_out/r02_final_bench_n1_dropout0.5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_cfg4.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_agg -s 2 -c 2 -o gpurun_out/r02_prof_agg -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_agg.log 2>&1
ls -la gpurun_out/r02_prof_agg.ncu-rep
python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/r02_final_bench_n1.json') if l.startswith('{')][-1]
print(round(d['ms_per_step'],2), d['value'], d['e2e']['ms_per_step'], d['roofline'], d['cpu_baseline'], d['clocks'])
P
