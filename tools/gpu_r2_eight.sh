#!/bin/bash
# N GPUs (default 8): headline bench (default exchange), paced ARKit stream, sustained 1000-frame stream (p2p exchange)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # name, args...
  name=$1; shift
  ( BNV_WATCHDOG=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu "$@" ) > gpurun_out/r2_${name}_n$N.json 2> gpurun_out/r2_${name}_n$N.err; echo "$name n$N rc=$?"
  grep '^{' gpurun_out/r2_${name}_n$N.json | tail -1 | cut -c1-900
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r2_${name}_n$N.err | tail -3
}
run bench --steps 20 --warmup 5
if [ "$N" = "8" ]; then
  run paced --paced-fps 60 --exchange p2p
  run sustained --sustained 1000 --exchange p2p
fi
