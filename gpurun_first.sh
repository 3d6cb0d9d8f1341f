set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 examples/deposition_fan.py --grid 1024 --check 2>&1 | tail -2
