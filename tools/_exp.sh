timeout -s KILL 300 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -3
timeout -s KILL 300 python tools/probe_tensor.py 1000000 100 | tail -3
timeout -s KILL 300 python tools/probe_tensor.py 10000000 100 | tail -3
timeout -s KILL 300 python tools/probe_tensor.py 10000000 10 | tail -2
