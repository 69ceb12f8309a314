# round-2 evidence: launch list of the bench command + ncu --set full of the trunk kernel in its three headline modes
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 420 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/launches_r02.log 2>&1; echo "launch list rc=$?"
for w in "c2 fp32" "c3 bf16" "c3 fp16"; do set -- $w
  ncu --set full --clock-control none --import-source on -k regex:tc_trunk_kernel -s 6 -c 1 -f -o gpurun_out/tc_trunk_r02_$1_$2 \
      python bench.py --workload $1 --precision $2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_$1_$2.log 2>&1; echo "ncu $w rc=$?"
done
python scripts/kernel_breakdown.py --workload c3 --precision bf16 > gpurun_out/breakdown_r02_c3_bf16.txt 2>&1; tail -12 gpurun_out/breakdown_r02_c3_bf16.txt
ls -la gpurun_out/*.ncu-rep | tail -4
