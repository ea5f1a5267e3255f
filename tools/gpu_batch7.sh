#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/batch_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/batch_tests.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu ) > gpurun_out/chk_bench_b200.json 2> gpurun_out/chk_bench_b200.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/chk_bench_b200.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("B", d["config"]["frames_per_step"], "single", {k:(round(v) if isinstance(v,float) else v) for k,v in d.get("single_frame_calls",{}).items() if k!="what"})
print("fps", round(d["value"]), "warm", round(d["value_warm"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_results_one_frame_behind"]["value"]), "local", round(d["local_scope"]["value"]), "pre_ms", round(r["prepass_ms"],4), "enc_ms", round(r["kernel_ms"],4),
      "fin_ms", round(r["finalize_ms"],4), "enc_frac", round(r["frac"],3), "hbm_frac", round(d["roofline_hbm"]["frac"],3))
PY
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'finalize_batch|frame_prepass_batch' --launch-skip 40 -c 2 -f \
    -o gpurun_out/r2f_batch python tools/batch_probe.py 7 > gpurun_out/r2f_batch.log 2>&1; echo "ncu rc=$?"
