#!/bin/bash
# round-2 final single-GPU evidence: smoke, all gpu tests, full default bench (with cpu baseline), ncu launch list + full capture
bash tools/gpu_r2_third.sh
( timeout 900 python bench.py ) > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.err; echo "full bench rc=$?"; tail -c 600 gpurun_out/r2f_bench_full.json
bash tools/gpu_profile.sh r2f
