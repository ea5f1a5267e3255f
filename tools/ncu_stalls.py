"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` dump: per kernel section, totals per
stall reason and the top instructions (SASS view).  usage: ncu_stalls.py <src.csv> [top_n] [kernel substring] [section #]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
filt = sys.argv[3] if len(sys.argv) > 3 else ""
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = {}
for si, s in enumerate(starts):
    name = rows[s][1]
    if filt not in name:
        continue
    seen[name] = seen.get(name, -1) + 1
    if seen[name] != which:
        continue
    hdr = rows[s + 1]
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    body = [r for r in rows[s + 2:end] if len(r) == len(hdr)]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    samp = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
    tot = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
    all_s = sum(int(r[samp] or 0) for r in body)
    print(f"## {name[:90]}")
    print("total samples", all_s, " instructions executed (warp-level)", sum(int(r[ex] or 0) for r in body), " SASS lines", len(body))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]:
        print(f"  {k:28s} {v:8d}  {100.0 * v / max(all_s, 1):5.1f} %")
    print("top instructions by samples (line#, samples, %, top reason, executed, SASS):")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:top_n]
    for i in order:
        r = body[i]
        why = max(stall_cols, key=lambda c: int(r[c] or 0))
        print(f"  {i:5d} {int(r[samp]):7d} {100.0 * int(r[samp]) / max(all_s, 1):5.1f}%  {hdr[why][6:]:14s} {int(r[ex] or 0):9d}  {r[src].strip()[:80]}")
