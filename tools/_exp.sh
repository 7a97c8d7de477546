python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for v in bulk ldg; do
BSR_SCAN_VARIANT=$v python bench.py --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err || tail -5 gpurun_out/bench_$v.err
done
for k in 50 1024; do python bench.py --steps 200 --warmup 10 --no-cpu-baseline --top-k $k > gpurun_out/bench_k$k.json 2>gpurun_out/bench_k$k.err; done
