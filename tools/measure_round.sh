# Round measurement pass on one B200 (run under gpurun): GPU tests, bench.py (+ reference arm), ncu launch
# lists and --set full captures (exported as CSV pages; the .ncu-rep files stay in /tmp), every single-GPU
# BASELINE config, and the tensor-kernel pipeline timelines.  Outputs land in gpurun_out/.
set -x
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; tail -c 300 gpurun_out/r01_bench_n1.json
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref_n1.json 2> gpurun_out/r01_bench_ref_n1.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk_bulk -s 4 -c 2 -o /tmp/r01b_scan_bulk -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none --csv --log-file gpurun_out/r01b_tensor_launches.csv python tools/probe_tensor_one.py 10000000 100 1024 > gpurun_out/ncu_tl.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk_bf16_tensor -s 4 -c 1 -o /tmp/r01b_tensor_pair -f python tools/probe_tensor_one.py 10000000 100 1024 > gpurun_out/ncu_tf.log 2>&1
timeout -s KILL 900 python tools/run_configs.py > gpurun_out/run_configs.log 2>&1; tail -2 gpurun_out/run_configs.log | cut -c1-300
BSR_LIB_VARIANT=tl BSR_TC_TL=1000 timeout -s KILL 120 python tools/probe_tensor_one.py 10000000 100 256 > gpurun_out/tl_full_1000.log 2>&1
BSR_LIB_VARIANT=tl BSR_TC_TL=0 timeout -s KILL 120 python tools/probe_tensor_one.py 10000000 100 256 > gpurun_out/tl_full_0.log 2>&1
ncu -i /tmp/r01b_scan_bulk.ncu-rep --page raw --csv > gpurun_out/r01b_scan_bulk_rawpage.csv 2>/dev/null
ncu -i /tmp/r01b_tensor_pair.ncu-rep --page raw --csv > gpurun_out/r01b_tensor_pair_rawpage.csv 2>/dev/null
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
