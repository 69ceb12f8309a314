mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/r2_tests1.log
for w in "c2 fp32" "c2 bf16" "c2 fp16" "c3 bf16" "c3 fp16" "c3 fp32"; do set -- $w; timeout 300 python bench.py --workload $1 --precision $2 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench1_$1_$2.json 2> gpurun_out/r2_bench1_$1_$2.err; echo "$w rc=$?"; done
DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f > /dev/null 2>&1
for m in fp32 bf16; do timeout 200 python scripts/dev/trunk_timeline.py $m > gpurun_out/r2_timeline1_$m.txt 2>&1; done
timeout 200 python scripts/dev/trunk_timeline.py fp16 3d > gpurun_out/r2_timeline1_fp16_3d.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'))
    except Exception as e: print(f, 'ERR', e)
PY
