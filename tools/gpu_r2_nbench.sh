#!/bin/bash
# N GPUs: headline bench only
N=${1:-8}
mkdir -p gpurun_out
( BNV_WATCHDOG=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --no-cpu --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"
grep '^{' gpurun_out/r2_bench_n$N.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v,4) for k,v in r.items() if k.endswith('_ms')}, d['shard_parity'])"
