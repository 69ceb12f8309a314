mkdir -p gpurun_out
for w in "c2 fp32" "c2 bf16" "c3 bf16"; do set -- $w
  timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench17_$1_$2.json 2> gpurun_out/r2_bench17_$1_$2.err; echo "$w rc=$?"; tail -c 300 gpurun_out/r2_bench17_$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench17_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), round(d['ms_per_step'],1))
    except Exception as e: print(f, 'ERR', e)
PY
