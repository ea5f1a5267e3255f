#!/bin/bash
# warp-specialised chain: correctness first (under a short timeout: a protocol bug would hang), then A/B numbers
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py smoke ) > gpurun_out/ws_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/ws_smoke.log
( timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu ) > gpurun_out/ws_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/ws_tests.log
for v in 1 0; do
  ( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --chain $v ) > gpurun_out/ws_bench_$v.json 2> gpurun_out/ws_bench_$v.err; echo "bench chain=$v rc=$?"
  python - $v <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(f"gpurun_out/ws_bench_{sys.argv[1]}.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("chain", sys.argv[1], "fps", round(d["value"]), "warm", round(d["value_warm"]), "e2e", round(d["e2e"]["value"]), "pre", round(r["prepass_ms"],4), "enc", round(r["kernel_ms"],4), "fin", round(r["finalize_ms"],4), "enc_frac", round(r["frac"],3),
          "generic Mq/s", round(d["decode"]["generic"]["value"]), "frac burst", round(d["decode"]["generic"]["roofline"]["frac_of_burst_peak"],3))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/ws_bench_{sys.argv[1]}.err").read()[-600:])
PY
done
