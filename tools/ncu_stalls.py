"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples + stall-reason totals.
Usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 | python tools/ncu_stalls.py [N]"""
import csv
import sys

n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:          # first table only (a report page may hold several kernels)
    if not r or r[0] in ("Address", "Kernel Name"):
        break
    if len(r) == len(hdr):
        body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
samp = ci["# Samples"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[samp] or 0) for r in body)
print(f"total samples {tot}, instructions {len(body)}")
agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100 * v / max(tot, 1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print(f"top {n_top} instructions by samples:")
for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][samp] or 0))[:n_top]:
    reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"  #{idx:5d} {int(r[samp] or 0):6d} ({100 * int(r[samp] or 0) / max(tot, 1):4.1f}%)  {r[ci['Source']].strip()[:70]:70s} "
          + " ".join(f"{n}:{c}" for c, n in reasons if c))
