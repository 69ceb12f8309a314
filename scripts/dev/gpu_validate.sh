mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke19.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r2_smoke19.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests19.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2_tests19.log
DGDM_TRUNK2=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shape.py -m gpu -x -q -k "bf16 or fp16" > gpurun_out/r2_tests19_trunk2.log 2>&1; echo "trunk2 tests rc=$?"; tail -2 gpurun_out/r2_tests19_trunk2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench19.json').read().strip().splitlines()[-1])
print('head', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'], d['gpu_launches'])
print('cpu', {k:(v if not isinstance(v,dict) else v.get('value')) for k,v in d['cpu_baseline'].items() if k!='sample'})
for k,v in d.get('extra',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), round(v['roofline']['frac'],3), round(v['ms_per_step'],1))
PY
