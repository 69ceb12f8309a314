#!/bin/bash
# same-box bench lines under an env switch of the trunk: VAR=name VALS="0 1 0 1" [TESTS=1 runs the trunk parity tests first]
cd /root/repo; mkdir -p gpurun_out
VAR=${VAR:-DGDM_TRUNK_PRE}
if [ -n "$TESTS" ]; then python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shape.py -m gpu -x -q -k "3d or variants or baseline" 2>&1 | tail -3; fi
line() { python bench.py "$@" --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print(d['config']['workload'][:3], d['dtype'], round(d['value']), d['roofline']['frac'], d['ms_per_step'], d['clocks']['sm_mhz'])"; }
for h in ${VALS:-0 1 0 1}; do
  echo "== $VAR=$h"
  env $VAR=$h bash -c "$(declare -f line); line --workload c3 --precision bf16; line --workload c3 --precision fp32"
done
