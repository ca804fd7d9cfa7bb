"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` (SASS view), with the dominant stall reasons.
usage: ncu -i rep --page source --csv | python tools/ncu_sass_hot.py [N]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
data = rows[hi + 1:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:N]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}%  exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{n}:{v}" for v, n in st if v))
