"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` dump:
totals per stall reason, and the top instructions (SASS view).  usage: ncu_stalls.py <src.csv> [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:                      # first kernel section only
    if r and r[0] in ("Address", "Kernel Name"):
        break
    if len(r) == len(hdr):
        body.append(r)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
samp = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
tot = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
all_s = sum(int(r[samp] or 0) for r in body)
print("total samples", all_s, " instructions executed (warp-level)", sum(int(r[ex] or 0) for r in body))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k:28s} {v:8d}  {100.0 * v / max(all_s, 1):5.1f} %")
print("top instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[samp] or 0))[:top_n]:
    why = max(stall_cols, key=lambda i: int(r[i] or 0))
    print(f"  {int(r[samp]):7d} {100.0 * int(r[samp]) / max(all_s, 1):5.1f}%  {hdr[why]:22s} {r[src].strip()[:90]}")
