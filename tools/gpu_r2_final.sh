#!/bin/bash
# full single-GPU check of the final kernels: smoke, all GPU tests, bench, ncu launch list + full capture
mkdir -p gpurun_out
bash tools/gpu_r2_third.sh
bash tools/gpu_profile.sh r2d
