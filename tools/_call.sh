timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -m gpu 2>&1 | tail -5
