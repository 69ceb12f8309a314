mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunking or two_tile or torch_library" > gpurun_out/r2_tests19.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_tests19.log | cut -c1-250
