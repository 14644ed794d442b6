"""Summarise an `ncu -i X.ncu-rep --page source --csv > X.csv` dump: top stall locations and stall-reason totals.
    python tools/ncu_stalls.py X.csv [n_lines]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
col = idx["# Samples"]
src = idx["Source"]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot_reason = {r: 0.0 for r in reasons}
for r in rows[2:]:
    try:
        s = float(r[col])
    except Exception:
        continue
    data.append((s, r[src], r))
    for k in reasons:
        try:
            tot_reason[k] += float(r[idx[k]])
        except Exception:
            pass
tot = sum(d[0] for d in data) or 1
print("stall reasons:", {k: round(v / tot * 100, 1) for k, v in sorted(tot_reason.items(), key=lambda kv: -kv[1])[:8]})
data.sort(key=lambda d: -d[0])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s, t, r in data[:n]:
    top = sorted(((float(r[idx[k]] or 0), k) for k in reasons), reverse=True)[:2]
    print(f"{s / tot * 100:5.1f}%  {t[:95]:95s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
