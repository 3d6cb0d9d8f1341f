set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['ctas_per_sm'], d['e2e'])"
