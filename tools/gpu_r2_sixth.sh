#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_third.sh
( timeout 300 python bench.py --sustained 1000 ) > gpurun_out/r2_sustained_n1.json 2> gpurun_out/r2_sustained_n1.err; echo "sustained rc=$?"; cut -c1-400 gpurun_out/r2_sustained_n1.json
( timeout 200 python bench.py --paced-fps 60 ) > gpurun_out/r2_paced_n1.json 2> gpurun_out/r2_paced_n1.err; echo "paced rc=$?"
bash tools/gpu_profile.sh r2b
