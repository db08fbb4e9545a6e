#!/bin/bash
# tuning: device-resident LM it/s against the super-tile size (observations per work item of k_pcg_solve) at N GPUs
N=${1:-4}; shift
for so in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --device-only --super-tile-obs $so
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --device-only --super-tile-obs $so 2>/dev/null
  fi
done
