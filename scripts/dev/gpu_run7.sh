mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unet or guided_loop" > gpurun_out/r2_tests7.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2_tests7.log
for a in "fp32 14 16384" "bf16 14 16384" "fp32 42 8192" "bf16 42 8192"; do python scripts/dev/unet_time.py $a; done 2>&1 | tee gpurun_out/r2_unet7.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_unet_launches7.csv python scripts/dev/unet_time.py bf16 42 8192 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_unet_launches7.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0]; v=float(r[-1].replace(',',''))
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print(f"{k[:60]:60s} n={v[0]:4d} {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%")
PY
