timeout 900 python -m pytest tests/test_gpu_roi.py tests/test_gpu_dropin.py -q -m gpu 2>&1 | tail -12
