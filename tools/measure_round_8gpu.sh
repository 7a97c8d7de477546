# Round measurement pass on EIGHT B200s of one box (run under `gpurun --gpus 8`): every GPU as a rank in the
# distributed tests, bench.py at N = 8 with both candidate exchanges, the config-5 sweeps.
set -x
nproc; free -g | head -2
timeout -s KILL 420 python -m pytest tests/test_gpu_distributed.py -x -q > gpurun_out/pytest_dist_n8.log 2>&1; tail -6 gpurun_out/pytest_dist_n8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout -s KILL 420 $TR --master-port 29521 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err ); tail -c 1200 gpurun_out/r02_bench_n8.json; tail -4 gpurun_out/r02_bench_n8.err
( time timeout -s KILL 300 $TR --master-port 29522 bench.py --gpus 8 --exchange nccl --steps 100 > gpurun_out/r02_bench_n8_nccl.json 2> gpurun_out/r02_bench_n8_nccl.err ); tail -c 600 gpurun_out/r02_bench_n8_nccl.json
( time timeout -s KILL 420 $TR --master-port 29523 tools/run_sweeps_multi.py > gpurun_out/sweeps_n8.log 2>&1 ); tail -c 1500 gpurun_out/sweeps_n8.log
