mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r2_tests_pair.log 2>&1; echo "variants rc=$?"; tail -6 gpurun_out/r2_tests_pair.log | cut -c1-300
for m in 2 1 0; do for w in "c2 bf16" "c3 bf16"; do set -- $w
  DGDM_TRUNK2=$m timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench_pair_m${m}_$1_$2.json 2> gpurun_out/r2_bench_pair_m${m}_$1_$2.err; echo "m=$m $w rc=$?"; tail -c 200 gpurun_out/r2_bench_pair_m${m}_$1_$2.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_pair_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), round(d['ms_per_step'],1))
    except Exception as e: print(f, 'ERR', e)
PY
