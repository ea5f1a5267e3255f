#!/bin/bash
# one gpurun call: smoke + GPU tests + bench, with a one-line summary
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/chk_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/chk_smoke.log
( timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/chk_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/chk_tests.log
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/chk_bench_b200.json 2> gpurun_out/chk_bench_b200.err; echo "bench rc=$?"
python - <<'PY'
import json
for v in ("b200",):
    try:
        d=json.loads(open(f"gpurun_out/chk_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "fps", round(d["value"]), "warm", round(d["value_warm"]), "e2e", round(d["e2e"]["value"]), "enc_ms", round(d["roofline"]["kernel_ms"],4),
              "fin_ms", round(d["roofline"]["finalize_ms"],4), "enc_frac", round(d["roofline"]["frac"],3),
              "dec_blocks Mq/s", round(d["decode"]["value"]), "generic Mq/s", round(d["decode"]["generic"]["value"]), "gen frac", round(d["decode"]["generic"]["roofline"]["frac"],3))
    except Exception as e:
        print(v, "failed", e)
PY
