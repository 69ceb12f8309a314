#!/bin/bash
# compute-sanitizer over the round-2 denoiser (fused conv + GroupNorm epilogue, gn2 with the first conv, edge zeroing)
# and the trimmed trunk epilogue
cd /root/repo; mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
O=gpurun_out/sanitizer_r02b.txt
echo "== memcheck, denoiser: test_unet_vs_oracle (7 shapes x 5 modes), golden, torch.library ops" > $O
timeout 1500 $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_unet_vs_oracle or test_unet_golden or torch_library" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> $O
echo "== memcheck, trunk (bf16 / fp16 / fp32, 2D + 3D) after the epilogue trim" >> $O
timeout 1500 $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_cond_fn_2d_golden or test_cond_fn_3d_golden" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> $O
echo "== racecheck, denoiser: test_unet_vs_oracle, n = 300 / 130 / 65 / 33 / 7 (P = 14, 42, 8, 48, 2), fp32 + bf16" >> $O
timeout 2400 $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_unet_vs_oracle and (300-14-fp32] or 130-42-bf16 or 65-8-bf16 or 33-48-fp32] or 7-2-bf16)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6 >> $O
echo "== racecheck, trunk bf16 2D" >> $O
timeout 2400 $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_cond_fn_2d_golden and bf16" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6 >> $O
cat $O
