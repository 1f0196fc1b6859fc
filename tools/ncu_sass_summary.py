"""Summarise the SASS page of an ncu report: instructions executed per region / opcode.
    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_sass_summary.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared Excessive")}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
print("total warp instructions", tot, "lines", len(body))
ops = collections.Counter()
for r in body:
    s = r[ci["Source"]].split()
    op = s[1] if s and s[0].startswith("@") else (s[0] if s else "?")
    ops[op.split(".")[0]] += int(r[ci["Instructions Executed"]] or 0)
for op, n in ops.most_common(top):
    print(f"{op:12s} {n:12d} {100*n/tot:6.2f}%")
if len(sys.argv) > 3:
    for i, r in enumerate(body):
        print(i, r[ci["Instructions Executed"]], r[ci["# Samples"]], r[ci["L1 Wavefronts Shared Excessive"]], r[ci["Source"]])
