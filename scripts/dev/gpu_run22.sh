mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shape.py -m gpu -x -q > gpurun_out/r2_tests22.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests22.log | cut -c1-300
for w in "c2 fp32" "c2 bf16" "c3 bf16"; do set -- $w
  timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench22_$1_$2.json 2> gpurun_out/r2_bench22_$1_$2.err; echo "$w rc=$?"; tail -c 300 gpurun_out/r2_bench22_$1_$2.err
  DGDM_TRUNK_SLOTS=1 timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench22s_$1_$2.json 2> gpurun_out/r2_bench22s_$1_$2.err; echo "$w slots rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench22*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), round(d['ms_per_step'],1))
    except Exception as e: print(f, 'ERR', e)
PY
ncu --set full --clock-control none -k regex:tc_trunk_kernel -s 6 -c 1 -f -o gpurun_out/tc_trunk_r02b_c2_fp32 python bench.py --workload c2 --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02b.log 2>&1; echo "ncu rc=$?"
