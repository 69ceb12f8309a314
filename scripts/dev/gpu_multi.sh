mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > gpurun_out/r2_tests18.log 2>&1; echo "multi-gpu tests rc=$?"
grep -E "^\{|passed|failed|skipped" gpurun_out/r2_tests18.log | cut -c1-330
for n in ${NS:-8 4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_bench18_${n}gpu.json 2> gpurun_out/r2_bench18_${n}gpu.err; echo "bench$n rc=$?"
  tail -c 300 gpurun_out/r2_bench18_${n}gpu.err | grep -v "OMP\|\*\*\*" 
done
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2_bench18_{n}gpu.json') if l.startswith('{')][-1])
        print(n,'head', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'])
        for k,v in d.get('extra',{}).items(): print(n, k, round(v['value']), round(v['e2e']['value']), round(v['roofline']['frac'],3), round(v['ms_per_step'],1))
    except Exception as e: print(n,'ERR',e)
PY
