"""Top stalled SASS lines of an ncu source-page CSV with their dominant stall reasons.
    python tools/ncu_hot_lines.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot)
agg = {}
for r in body:
    for c in stall_cols:
        agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} exec {r[ie]:>9s} samp {r[si]:>6s}  {st}  {r[src][:70]}")
