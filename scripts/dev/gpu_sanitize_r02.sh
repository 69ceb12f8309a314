mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck, default kernels (fp16 / fp16x3 / bf16 / fp32 trunk modes, 2D + 3D, ops)" > gpurun_out/sanitizer_r02.txt
timeout 1500 $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_cond_fn_2d_golden or test_cond_fn_3d_golden or torch_library or chunking" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> gpurun_out/sanitizer_r02.txt
echo "== memcheck, two-tile kernel (DGDM_TRUNK2=1)" >> gpurun_out/sanitizer_r02.txt
DGDM_TRUNK2=1 timeout 1500 $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(test_cond_fn_2d_golden or test_cond_fn_3d_golden) and (bf16 or fp16)" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> gpurun_out/sanitizer_r02.txt
echo "== racecheck, default kernels, single-pass fp16 + fp16x3 (new in round 2)" >> gpurun_out/sanitizer_r02.txt
timeout 2400 $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_cond_fn_2d_golden and fp16" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6 >> gpurun_out/sanitizer_r02.txt
echo "== racecheck, two-tile kernel (DGDM_TRUNK2=1), bf16" >> gpurun_out/sanitizer_r02.txt
DGDM_TRUNK2=1 timeout 2400 $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_cond_fn_2d_golden and bf16" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6 >> gpurun_out/sanitizer_r02.txt
cat gpurun_out/sanitizer_r02.txt
