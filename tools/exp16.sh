set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "philox or fused_noise" 2>&1 | tail -5
PDM_PHILOX_SCALAR=1 timeout 200 python tools/quick_noise.py 2>&1 | tail -4
timeout 200 python tools/quick_noise.py 2>&1 | tail -4
