"""Hottest SASS lines (by stall samples) of launch #idx in an .ncu-rep, grouped with source line info."""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
si, sc, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
def num(x):
    try: return int(x)
    except Exception: return 0
tot = sum(num(r[si]) for r in body)
print("total samples", tot, "sass lines", len(body))
order = sorted(range(len(body)), key=lambda i: -num(body[i][si]))[:topn]
for i in sorted(order):
    r = body[i]
    st = {k: num(r[h.index(k)]) for k in h if k.startswith('stall_') and 'Not Issued' not in k}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(i, f"{100*num(r[si])/max(tot,1):.1f}%", r[ie], r[sc].strip()[:80], '|', ' '.join(f"{k[6:]}={v}" for k, v in top if v))
