nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -2 gpurun_out/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 200 --warmup 10 --rows-per-gpu 12500000 > gpurun_out/bench_n8_100m.json 2> gpurun_out/bench_n8_100m.err; tail -2 gpurun_out/bench_n8_100m.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 500 --warmup 10 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -2
