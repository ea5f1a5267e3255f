import os, sys, json, subprocess
for nwg in ("4","3","2"):
    env=dict(os.environ, BNV_TC_NWG=nwg)
    r=subprocess.run([sys.executable,"bench.py","--mlp","tc16","--steps","5","--warmup","3","--no-cpu"],env=env,capture_output=True,text=True)
    try:
        j=json.loads(r.stdout.strip().splitlines()[-1]); print("NWG",nwg,"decode Mq/s",round(j["decode"]["value"]),"ms",round(j["decode"]["ms"],3))
    except Exception as e: print(nwg, r.stdout[-300:], r.stderr[-300:])
