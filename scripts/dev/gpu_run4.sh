mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_shape.py -m gpu -q -s > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?"
grep -E "full run|passed|failed|ref " gpurun_out/r2_tests4.log | cut -c1-230
for w in "c2 fp16x3" "c3 fp16x3"; do set -- $w; timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench4_$1_$2.json 2> gpurun_out/r2_bench4_$1_$2.err; echo "$w rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'))
    except Exception as e: print(f, 'ERR', e)
PY
