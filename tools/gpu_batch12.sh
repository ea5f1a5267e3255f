#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_batch.py -q -m gpu -x ) > gpurun_out/batch_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/batch_tests.log
( timeout 300 python tools/batch_probe.py 7 0 ) 2> gpurun_out/batch_probe.err | tee gpurun_out/batch_probe.jsonl
tail -3 gpurun_out/batch_probe.err
