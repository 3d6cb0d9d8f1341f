set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15
RAYS_B200_REGISTER_HOST=0 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tail -3
