python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/probe_batch.py 1000000 10
python tools/probe_batch.py 1000000 100 | head -5
python bench.py --steps 300 --warmup 10 --no-cpu-baseline | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['avg_launch_ms'], j['e2e']['value'])"
