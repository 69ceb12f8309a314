mkdir -p gpurun_out
DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f > /dev/null 2>&1 || echo build failed
python scripts/dev/trunk2_timeline.py bf16 2d > gpurun_out/r2_timeline9_bf16.txt 2>&1; tail -3 gpurun_out/r2_timeline9_bf16.txt
python scripts/dev/trunk2_timeline.py bf16 3d > gpurun_out/r2_timeline9_bf16_3d.txt 2>&1; tail -3 gpurun_out/r2_timeline9_bf16_3d.txt
