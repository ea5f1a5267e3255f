#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_third.sh
( timeout 300 python tools/prepass_ablation.py ) > gpurun_out/r2_prepass_ablation.txt 2>&1; echo "ablation rc=$?"; grep flags gpurun_out/r2_prepass_ablation.txt
( BNV_LIB=$PWD/bnv_fusion_b200/libbnv_b200_prof.so timeout 200 python tools/chain_phase_profile.py ) > gpurun_out/r2_chain_phase3.txt 2>&1; echo "phase rc=$?"
