#!/bin/bash
# ncu --set full of the fused conv + GroupNorm launches of one denoiser forward (final tree), at the two bench shapes;
# only the text summaries come back (the reports are ~2.5 MB per launch)
cd /root/repo; mkdir -p gpurun_out
for cfg in "bf16 42 8192" "fp32 14 16384"; do set -- $cfg
  R=/tmp/conv_gn_r02b_$1_p$2
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 63 -c 9 -f -o $R \
      python scripts/dev/unet_time.py $cfg > /dev/null 2>&1; echo "ncu $cfg rc=$?"
  ncu -i $R.ncu-rep --page details > gpurun_out/conv_gn_r02b_$1_p$2_details.txt 2>/dev/null
  ncu -i $R.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__inst_executed.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu.sum > gpurun_out/conv_gn_r02b_$1_p$2_raw.csv 2>/dev/null
done
ls -la gpurun_out/conv_gn_r02b_*
