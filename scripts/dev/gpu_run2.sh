mkdir -p gpurun_out
./scripts/microbench/mma_tmem_conflict > gpurun_out/r2_mma_tmem_conflict.txt 2>&1
cat gpurun_out/r2_mma_tmem_conflict.txt
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2_tests2.log
