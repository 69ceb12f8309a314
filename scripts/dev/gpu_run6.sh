mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_smi6.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2_tests6.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench6.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench6_ref.json 2> gpurun_out/r2_bench6_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1])
print('head', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'])
print('cpu', d['cpu_baseline'])
for k,v in d.get('extra',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), round(v['roofline']['frac'],3), round(v['ms_per_step'],1))
PY
