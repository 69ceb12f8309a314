mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_shape.py -m gpu -q -s -k "fp16 or fp32" > gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?"
grep -E "full run|passed|failed|ref |cond_fn vs" gpurun_out/r2_tests5.log | grep -v simt | cut -c1-200
