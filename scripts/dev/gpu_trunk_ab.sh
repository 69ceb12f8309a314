#!/bin/bash
# trunk epilogue instruction trimming: full GPU suite + same-box bench lines, with / without the try_wait suspend hint
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
line() { python bench.py "$@" --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print(d['config']['workload'], d['dtype'], round(d['value']), d['roofline']['frac'], d['ms_per_step'], d['clocks']['sm_mhz'])"; }
for h in 0 2000 0 20000; do
  echo "== DGDM_TRUNK_WAIT_HINT=$h"
  DGDM_TRUNK_WAIT_HINT=$h line --workload c3 --precision bf16
  DGDM_TRUNK_WAIT_HINT=$h line --workload c2 --precision bf16
  DGDM_TRUNK_WAIT_HINT=$h line --workload c2 --precision fp32
done
