mkdir -p gpurun_out
DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f > /dev/null 2>&1 || echo build failed
python scripts/dev/trunk_timeline.py bf16 2d > gpurun_out/r2_timeline18_bf16.txt 2>&1; grep "tile:" gpurun_out/r2_timeline18_bf16.txt
python scripts/dev/trunk_timeline.py fp32 2d > gpurun_out/r2_timeline18_fp32.txt 2>&1; grep "tile:" gpurun_out/r2_timeline18_fp32.txt
