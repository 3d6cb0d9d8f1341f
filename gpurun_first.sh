set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25
