set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sg or SG" 2>&1 | tail -4
for opt in "--ode SG_ODE --deriv cold --rays 262144" "--ode SG_ODE --rays 131072"; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu $opt 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$opt', d['value'], d['ms_per_step'], d['config']['ray_steps_per_fan'], d['roofline']['frac'], d['roofline']['ctas_per_sm'])"
done
