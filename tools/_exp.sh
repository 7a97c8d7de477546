python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -15
