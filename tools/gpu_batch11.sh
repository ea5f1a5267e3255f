#!/bin/bash
mkdir -p gpurun_out
for sp in 1 2; do
  ( BNV_PREPASS_SPLIT=$sp timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_batch.py -q -m gpu -x ) > gpurun_out/split${sp}_tests.log 2>&1; echo "split=$sp tests rc=$?"; tail -2 gpurun_out/split${sp}_tests.log
done
rm -f gpurun_out/split_probe.jsonl
for sp in 0 1; do
  ( BNV_PREPASS_SPLIT=$sp timeout 300 python tools/batch_probe.py 7 0 ) 2>> gpurun_out/split_probe.err | sed "s/^{/{\"split\": $sp, /" | tee -a gpurun_out/split_probe.jsonl
done
tail -3 gpurun_out/split_probe.err
