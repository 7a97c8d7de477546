# Round measurement pass on ONE B200 (run under gpurun): GPU tests, bench.py (+ reference arm), ncu launch
# lists and --set full captures (exported as CSV raw pages; the .ncu-rep files stay in /tmp on the box),
# every single-GPU BASELINE config, a clock record.  Outputs land in gpurun_out/ (copy what is to be
# judged into profiles/ with tools/summarize_profiles.py).
set -x
R=r02
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${R}_clocks_bench.csv &
SMI=$!
timeout -s KILL 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 300 gpurun_out/${R}_bench_n1.json
kill $SMI
timeout -s KILL 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_ref_n1.json 2> gpurun_out/${R}_bench_ref_n1.err
# launch list of the bench command (shares, not absolutes)
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
$B > gpurun_out/plain_b.log 2>&1 && timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/ncu_l.log 2>&1
# fp32 scan, 12.5 M rows (the bench kernel)
$B > gpurun_out/plain_b.log 2>&1 && timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk_bulk -s 4 -c 2 -o /tmp/${R}_scan_f32_12m5 -f $B > gpurun_out/ncu_f.log 2>&1
ncu -i /tmp/${R}_scan_f32_12m5.ncu-rep --page raw --csv > gpurun_out/${R}_scan_f32_12m5_raw.csv 2>/dev/null
# bf16 scan, 1 M rows
P="python tools/probe_dtype.py 1000000 bf16 10 10"
$P > gpurun_out/plain_p.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk_bulk -s 6 -c 2 -o /tmp/${R}_scan_bf16_1m -f $P > gpurun_out/ncu_p.log 2>&1
ncu -i /tmp/${R}_scan_bf16_1m.ncu-rep --page raw --csv > gpurun_out/${R}_scan_bf16_1m_raw.csv 2>/dev/null
# fp32 scan, 1 M rows
P="python tools/probe_dtype.py 1000000 f32 10 10"
$P > gpurun_out/plain_p.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk_bulk -s 6 -c 2 -o /tmp/${R}_scan_f32_1m -f $P > gpurun_out/ncu_p2.log 2>&1
ncu -i /tmp/${R}_scan_f32_1m.ncu-rep --page raw --csv > gpurun_out/${R}_scan_f32_1m_raw.csv 2>/dev/null
# tensor pipeline: launch list (bf16 + fp32 stores) and the pair-mode kernel
T="python tools/probe_tensor.py 10000000 bf16 100 1024 2"
$T > gpurun_out/plain_t.log 2>&1 && timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none --csv --log-file gpurun_out/${R}_tensor_bf16_launches.csv $T > gpurun_out/ncu_tl.log 2>&1
T2="python tools/probe_tensor.py 10000000 f32 100 1024 2"
$T2 > gpurun_out/plain_t2.log 2>&1 && timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none --csv --log-file gpurun_out/${R}_tensor_f32_launches.csv $T2 > gpurun_out/ncu_tl2.log 2>&1
$T > gpurun_out/plain_t.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk_bf16_tensor -s 4 -c 1 -o /tmp/${R}_tensor_pair -f $T > gpurun_out/ncu_tf.log 2>&1
ncu -i /tmp/${R}_tensor_pair.ncu-rep --page raw --csv > gpurun_out/${R}_tensor_pair_raw.csv 2>/dev/null
cat gpurun_out/plain_t.log gpurun_out/plain_t2.log
# every single-GPU BASELINE config
timeout -s KILL 1200 python tools/run_configs.py > gpurun_out/run_configs.log 2>&1; tail -2 gpurun_out/run_configs.log | cut -c1-300
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
