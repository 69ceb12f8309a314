mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r2_tests10.log 2>&1; echo "multi-gpu tests rc=$?"
grep -E "^\{|passed|failed|skipped" gpurun_out/r2_tests10.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench10_2gpu.json 2> gpurun_out/r2_bench10_2gpu.err; echo "bench2 rc=$?"
tail -c 500 gpurun_out/r2_bench10_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench10_2gpu.json') if l.startswith('{')][-1])
print('head', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'])
for k,v in d.get('extra',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), round(v['roofline']['frac'],3), round(v['ms_per_step'],1), v['e2e']['h2d_bytes_per_step'], v['e2e']['d2h_bytes_per_step'])
PY
