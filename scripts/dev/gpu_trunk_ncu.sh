#!/bin/bash
# ncu --set full of one guidance launch of the trunk kernel in the final tree (text summaries + key raw metrics)
cd /root/repo; mkdir -p gpurun_out
for w in "c2 fp32" "c3 bf16"; do set -- $w
  R=/tmp/tc_trunk_r02c_$1_$2
  ncu --set full --clock-control none --import-source on -k regex:tc_trunk_kernel -s 6 -c 1 -f -o $R \
      python bench.py --workload $1 --precision $2 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1; echo "ncu $w rc=$?"
  ncu -i $R.ncu-rep --page details > gpurun_out/tc_trunk_r02c_$1_$2_details.txt 2>/dev/null
  ncu -i $R.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__inst_executed.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second > gpurun_out/tc_trunk_r02c_$1_$2_raw.csv 2>/dev/null
  tail -1 gpurun_out/tc_trunk_r02c_$1_$2_raw.csv | cut -c1-400
done
