RAYS_B200_SLICE=7 RAYS_B200_REGISTER_HOST=0 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
RAYS_B200_REGISTER_HOST=0 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
