#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_third.sh
python - <<'PY'
import json
d=json.loads(open("gpurun_out/chk_bench_b200.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "without kernel events", round(d["value_without_kernel_events"]), "local scope", round(d["local_scope"]["value"]))
PY
( timeout 300 python tools/prepass_ablation.py ) > gpurun_out/r2_prepass_ablation2.txt 2>&1; echo "ablation rc=$?"; grep flags gpurun_out/r2_prepass_ablation2.txt
