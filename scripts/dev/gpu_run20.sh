mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shape.py -m gpu -x -q -s -k "fp16" > gpurun_out/r2_tests20.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests20.log
grep -E "^\[(2d|3d) fp16\]|^\[(2d|3d) fp16 " gpurun_out/r2_tests20.log | cut -c1-200 | head -14
for w in "c3 fp16" "c2 fp16" "c3 bf16"; do set -- $w
  timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench20_$1_$2.json 2> gpurun_out/r2_bench20_$1_$2.err; echo "$w rc=$?"; tail -c 300 gpurun_out/r2_bench20_$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench20_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), round(d['ms_per_step'],1))
    except Exception as e: print(f, 'ERR', e)
PY
