mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or fp16 or variants" > gpurun_out/r2_tests23.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests23.log | cut -c1-200
for rep in 1 2; do for m in 1 0; do for w in "c2 bf16" "c3 bf16" "c3 fp16"; do set -- $w
  DGDM_TRUNK_MERGE=$m timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench23_m${m}_r${rep}_$1_$2.json 2> /dev/null
done; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench23_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), round(d['ms_per_step'],1))
    except Exception as e: print(f, 'ERR', e)
PY
