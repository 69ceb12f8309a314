#!/bin/bash
# denoiser: parity tests, same-box A/B of the fused conv+GroupNorm launches (DGDM_UNET_FUSED=0/1), launch list of one forward
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "unet or denois or ops or loop or guided" 2>&1 | tail -5
for f in 0 1; do
  for cfg in "bf16 14 16384" "bf16 42 8192" "fp32 14 16384" "fp32 42 8192"; do
    echo -n "fused=$f  "; DGDM_UNET_FUSED=$f python scripts/dev/unet_time.py $cfg 2>&1 | tail -1
  done
done
if [ "$1" = "prof" ]; then
for cfg in "bf16 14 16384" "fp32 42 8192"; do set -- $cfg
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 54 -c 27 --csv \
      --log-file gpurun_out/unet_launches_r02_$1_p$2.csv python scripts/dev/unet_time.py $cfg > /dev/null 2>&1; echo "launch list $cfg rc=$?"
done
fi
python bench.py --workload c3 --precision bf16 --steps 5 --warmup 3 --no-extra 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('c3 bf16', d['value'], d['roofline']['frac'], d['ms_per_step'])"
python bench.py --steps 5 --warmup 3 --no-extra 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('c2 fp32', d['value'], d['roofline']['frac'], d['ms_per_step'])"
