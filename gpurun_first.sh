
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['ctas_per_sm'], d['e2e'])"
