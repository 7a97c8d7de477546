python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 300 python tools/probe_tensor.py 10000000 100 > gpurun_out/tensor_10m_k100.txt 2>&1; tail -3 gpurun_out/tensor_10m_k100.txt | cut -c1-140
timeout -s KILL 300 python tools/probe_tensor_one.py 10000000 100 128 > gpurun_out/plain_t.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tensor_kernel -s 2 -c 1 -f -o gpurun_out/r01_tensor_final python tools/probe_tensor_one.py 10000000 100 128 > gpurun_out/ncu_t.log 2>&1
timeout -s KILL 300 python tools/probe_tensor_one.py 10000000 100 128 > gpurun_out/plain_t2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_tensor_launches.csv python tools/probe_tensor_one.py 10000000 100 128 > gpurun_out/ncu_t2.log 2>&1
