#!/bin/bash
# bench matrix used during development: prints "workload precision designs/s ms frac share e2e"
EXTRA="$1"
for w in "c2 fp32" "c2 bf16" "c3 bf16" "c3 fp32" "c2g36 fp32" "c2g36 bf16"; do
  set -- $w
  python bench.py --workload $1 --precision $2 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2>gpurun_out/bench_err.txt || tail -5 gpurun_out/bench_err.txt
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}.json"))
r = d["roofline"]
print(sys.argv[1], sys.argv[2], round(d["value"]), round(d["ms_per_step"], 2), "frac", round(r["frac"], 3), "share",
      round(r["kernel_share_of_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
PY
done
if [ "$EXTRA" = "c5" ]; then
  python bench.py --workload c5 --precision bf16 --no-cpu-baseline --steps 3 > gpurun_out/bench_c5_bf16.json && python -c "
import json;d=json.load(open('gpurun_out/bench_c5_bf16.json'));print('c5 bf16 1gpu',round(d['value']),round(d['ms_per_step'],1),round(d['roofline']['frac'],3),round(d['e2e']['value']))"
fi
