#!/bin/bash
# 2 GPUs: tile-shard parity tests (NCCL all-gather and peer-memory exchange, both MLP modes), then bench with both exchanges
mkdir -p gpurun_out
nvidia-smi -L | head -4
( timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu ) > gpurun_out/r2_two_gpu_tests.log 2>&1; echo "dist tests rc=$?"; tail -8 gpurun_out/r2_two_gpu_tests.log
for ex in nccl p2p; do
  ( BNV_WATCHDOG=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --exchange $ex ) > gpurun_out/r2_bench_n2_$ex.json 2> gpurun_out/r2_bench_n2_$ex.err; echo "bench n2 $ex rc=$?"
  python - "$ex" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(f"gpurun_out/r2_bench_n2_{sys.argv[1]}.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print(sys.argv[1], "fps", round(d["value"]), "warm", round(d["value_warm"]), "e2e", round(d["e2e"]["value"]), "pre", round(r["prepass_ms"],4), "enc", round(r["kernel_ms"],4), "fin", round(r["finalize_ms"],4), "parity", d.get("shard_parity"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
  tail -3 gpurun_out/r2_bench_n2_$ex.err
done
