#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_batch.py -q -m gpu -x ) > gpurun_out/batch_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/batch_tests.log
( timeout 600 python tools/batch_probe.py 0 4 7 8 12 ) > gpurun_out/batch_probe.jsonl 2> gpurun_out/batch_probe.err; echo "probe rc=$?"; cat gpurun_out/batch_probe.jsonl; tail -5 gpurun_out/batch_probe.err
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'frame_prepass_batch|encode_ws|finalize_batch' --launch-skip 30 -c 3 -f \
    -o gpurun_out/r2e_batch python tools/batch_probe.py 7 > gpurun_out/r2e_batch.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2e_batch.log
