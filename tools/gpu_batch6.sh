#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/batch_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/batch_tests.log
rm -f gpurun_out/batch_ablate.jsonl
for d in 0 1 2; do
  ( BNV_PROBE_DBG=$d timeout 300 python tools/batch_probe.py 7 0 ) 2>> gpurun_out/batch_ablate.err | tee -a gpurun_out/batch_ablate.jsonl
done
tail -3 gpurun_out/batch_ablate.err
