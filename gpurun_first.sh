set -x
RAYS_B200_SLICE=7 RAYS_B200_REGISTER_HOST=0 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "three_species_generic_kernel" 2>&1 | grep -E "^E|assert|Error" | head -20
for sl in 250 125 400 0; do
RAYS_B200_SLICE=$sl timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('slice=$sl', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_kernel_ms'], d['roofline']['resume_pass_ms'], d['gpu_launches'])"
done
