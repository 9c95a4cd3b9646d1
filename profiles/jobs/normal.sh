timeout 900 python -m pytest tests/test_gpu_normal_paint.py -m gpu -x -q 2>&1 | tail -25
