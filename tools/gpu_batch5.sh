#!/bin/bash
mkdir -p gpurun_out
for d in 0 2 1 3 4 31; do
  ( BNV_PROBE_DBG=$d timeout 300 python tools/batch_probe.py 7 0 ) 2>> gpurun_out/batch_ablate.err | tee -a gpurun_out/batch_ablate.jsonl
done
tail -3 gpurun_out/batch_ablate.err
