#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/chk_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/chk_smoke.log
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/chk_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/chk_tests.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu ) > gpurun_out/chk_bench_b200.json 2> gpurun_out/chk_bench_b200.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/chk_bench_b200.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("B", d["config"]["frames_per_step"], "single", {k: (round(v) if isinstance(v, float) else v) for k, v in (d.get("single_frame_calls") or {}).items() if k != "what"})
print("mesh", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get("mesh_extraction") or {}).items() if k != "what"})
print("optim", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get("optim_iteration") or {}).items() if k != "what"})
print("fps", round(d["value"]), "warm", round(d["value_warm"]), "e2e", round(d["e2e"]["value"]), "pre_ms", round(r["prepass_ms"],4), "enc_ms", round(r["kernel_ms"],4),
      "fin_ms", round(r["finalize_ms"],4), "enc_frac", round(r["frac"],3), "hbm_frac", round(d["roofline_hbm"]["frac"],3),
      "dec_blocks Mq/s", round(d["decode"]["value"]), "generic Mq/s", round(d["decode"]["generic"]["value"]), "gen frac burst", round(d["decode"]["generic"]["roofline"]["frac_of_burst_peak"],3))
PY
