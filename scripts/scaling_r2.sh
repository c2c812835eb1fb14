#!/bin/bash
# usage: scripts/scaling_r2.sh N   (on the GPU box)  -> gpurun_out/scaling_N.txt
N=$1
mkdir -p gpurun_out
out=gpurun_out/scaling_$N.txt
{
  echo "== N=$N  $(date -u +%FT%TZ)  $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
  echo "-- torchrun -n $N scripts/h2d_ceiling.py"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    scripts/h2d_ceiling.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4
  echo "-- bench.py --gpus $N --steps 5 --warmup 3"
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 2>&1 | tail -1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1
  fi
} > $out 2>&1
cat $out
