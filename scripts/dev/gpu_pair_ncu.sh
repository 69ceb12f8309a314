mkdir -p gpurun_out
for m in 0 2; do
DGDM_TRUNK2=$m ncu --metrics lts__t_sectors_srcunit_tex.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_trunk -s 6 -c 1 --csv --log-file gpurun_out/ncu_pair_m$m.csv python bench.py --workload c2 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "m=$m rc=$?"
grep -E "tc_trunk" gpurun_out/ncu_pair_m$m.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
done
