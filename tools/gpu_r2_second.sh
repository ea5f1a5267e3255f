#!/bin/bash
mkdir -p gpurun_out
( BNV_LIB=$PWD/bnv_fusion_b200/libbnv_b200_prof.so timeout 200 python tools/chain_phase_profile.py ) > gpurun_out/r2_chain_phase2.txt 2>&1; echo "phase rc=$?"
( timeout 600 python -m pytest tests -x -q -m gpu ) > gpurun_out/chk_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/chk_tests.log
bash tools/gpu_profile.sh r2a
