#!/bin/bash
# full GPU suite + the three bench lines
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
line() { python bench.py "$@" --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print(d['config']['workload'][:3], d['dtype'], round(d['value']), round(d['e2e']['value']), d['roofline']['frac'], d['ms_per_step'], d['clocks']['sm_mhz'])"; }
line --workload c3 --precision bf16
line --workload c2 --precision bf16
line --workload c2 --precision fp32
line --workload c3 --precision fp32
