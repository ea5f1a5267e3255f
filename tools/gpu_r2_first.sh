#!/bin/bash
# round-2 first GPU call: baseline check, chain phase timers (profiling build), paced ARKit stream at N=1
mkdir -p gpurun_out
timeout 900 bash tools/gpu_check.sh
( BNV_LIB=$PWD/bnv_fusion_b200/libbnv_b200_prof.so timeout 200 python tools/chain_phase_profile.py ) > gpurun_out/r2_chain_phase.txt 2>&1; echo "phase rc=$?"
cat gpurun_out/r2_chain_phase.txt | tail -14
( timeout 200 python bench.py --paced-fps 60 ) > gpurun_out/r2_paced_n1.json 2> gpurun_out/r2_paced_n1.err; echo "paced rc=$?"
tail -1 gpurun_out/r2_paced_n1.json
