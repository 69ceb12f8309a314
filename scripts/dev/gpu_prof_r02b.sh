#!/bin/bash
# refresh after the denoiser fusion: launch list of the bench command, warm per-kernel breakdown of a C3 bf16 pass
cd /root/repo; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 420 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/launches_r02b.log 2>&1; echo "launch list rc=$?"
python scripts/kernel_breakdown.py --workload c3 --precision bf16 > gpurun_out/breakdown_r02b_c3_bf16.txt 2>&1; tail -14 gpurun_out/breakdown_r02b_c3_bf16.txt
python scripts/kernel_breakdown.py --workload c2 --precision fp32 > gpurun_out/breakdown_r02b_c2_fp32.txt 2>&1; tail -14 gpurun_out/breakdown_r02b_c2_fp32.txt
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r02b_line.json; python -c "
import json; d = json.loads(open('gpurun_out/bench_r02b_line.json').read()); print(d['value'], d['roofline'])"
