mkdir -p gpurun_out
DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f > /dev/null 2>&1 || echo build failed
DGDM_TRUNK2=2 python scripts/dev/trunk2_timeline.py bf16 2d > gpurun_out/r2_timeline_pair_bf16.txt 2>&1; grep issuer: gpurun_out/r2_timeline_pair_bf16.txt
