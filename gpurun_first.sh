
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --rays 131072 --steps 2 --warmup 1 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tail -5
