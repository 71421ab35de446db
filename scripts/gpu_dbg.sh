timeout 900 python -m pytest tests/test_gpu_f3f4.py -q -m gpu 2>&1 | tail -30
