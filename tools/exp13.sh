set -x
timeout 300 python -m pytest tests/test_gpu_screen.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/quick_screen.py --n 200000 --d 12288 --b 1024 --nt 100 --tmin 1e-4 --tmax 1e8 --iters 1 2>&1 | tail -6
timeout 300 python tools/quick_screen.py --n 100000 --d 16384 --b 1024 --nt 50 --tmin 1e-4 --tmax 1e4 --data sphere --iters 1 2>&1 | tail -6
