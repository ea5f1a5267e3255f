#!/bin/bash
# one gpurun call: launch list of the bench command + ncu --set full of the hot kernels (B200_PROFILING.md recipe)
#   usage: tools/gpu_profile.sh <tag> [kernel-regex]
mkdir -p gpurun_out
TAG=${1:-r2a}
KERN=${2:-'frame_prepass|encode_ws|finalize_fused|finalize_batch|decode_ws|gtable_ws|blend_blocks|tsdf_integrate|mesh_'}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-optim > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"$KERN" --launch-skip 15 -c 13 -f \
    -o gpurun_out/${TAG}_hot python tools/profile_workload.py > gpurun_out/${TAG}_hot.log 2>&1; echo "ncu full rc=$?"
tail -2 gpurun_out/${TAG}_hot.log
