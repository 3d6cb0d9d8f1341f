set -x
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e 2>&1 | tail -3
