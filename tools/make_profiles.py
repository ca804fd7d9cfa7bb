"""Condense gpurun_out/ ncu captures into the tracked summaries under profiles/ (round tag as argv[1])."""
import csv, json, sys, collections
from pathlib import Path
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
out = Path("profiles"); out.mkdir(exist_ok=True)
g = Path("gpurun_out")

# 1. launch list -> per-launch rows of the first full sub-batch + per-family shares
rows = [r for r in csv.reader(open(g / f"launches_{mode}.csv")) if len(r) > 10 and r[0].isdigit()]
def kname(full):
    full = full.replace("void ", "")
    i = full.find(">(")   # templated kernel: keep the template arguments, drop the parameter list
    return full[: i + 1] if i >= 0 else full.split("(")[0]
launches = [(int(r[0]), kname(r[4]), r[7], r[8], float(r[14]) / 1e3) for r in rows]
def family(name):
    # pw_tc_kernel<T, GATED, RELU, POOL, TS>: ungated = expand / head layers, gated = project layers, pool = head conv + pool;
    # TS = A operands through tensor memory
    if "pw_tc_kernel" in name:
        args = [a.strip() for a in name[name.find("<") + 1: name.rfind(">")].split(",")]
        args += ["0"] * (5 - len(args))
        ts = " TS" if args[4] in ("1", "true") else ""
        if args[3] in ("1", "true"): return "pw_tc(head conv + pool)" + ts
        if args[2] in ("1", "true"): return "pw_tc(mlp head)" + ts
        return ("pw_tc(project, gated)" if args[1] in ("1", "true") else "pw_tc(expand)") + ts
    for k in ("synth_image", "stem_tc", "stem", "mbconv_fused", "dw_tma", "dw_reg", "se_kernel", "avgpool", "pw_simt", "head_rows"):
        if k in name: return k
    return "other(torch)"
# one sub-batch = from the first stem launch to the launch before the second stem (or head_rows)
stems = [i for i, l in enumerate(launches) if "stem" in l[1]]
lo = stems[0]; hi = stems[1] if len(stems) > 1 else len(launches)
sub = launches[lo:hi]
with open(out / f"{tag}_launches_{mode}.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : one 500-patch sub-batch of `python bench.py --images 5 --batch 500 --steps 1 --no-cpu-baseline --no-sub --mode %s` (cold-cache, serialised: compare shares)\n" % mode)
    f.write("launch,kernel,block,grid,us\n")
    for i, n, b, gr, us in sub: f.write(f"{i},\"{n}\",\"{b}\",\"{gr}\",{us:.1f}\n")
fam = collections.OrderedDict()
for _, n, _, _, us in sub: fam[family(n)] = fam.get(family(n), 0.0) + us
tot = sum(fam.values())
shares = {k: round(v / tot, 4) for k, v in fam.items()}

# 2. per-kernel full-set summaries (already condensed on the GPU box by tools/ncu_capture.sh)
summ = []
kf = g / f"kernels_{mode}.csv"
if kf.exists():
    lines = [l for l in kf.read_text().splitlines() if l.strip()]
    hdr = lines[0].split(",")
    for l in lines[1:]:
        if l.startswith("capture,"): continue
        summ.append(dict(zip(hdr, l.split(","))))
    with open(out / f"{tag}_kernels_{mode}.csv", "w") as f:
        f.write("# ncu --set full --clock-control none, one launch each (500 patches); dram_*_MB = dram__bytes_{read,write}.sum per launch\n")
        f.write("\n".join(lines) + "\n")
import subprocess
try:
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except OSError:
    head = ""
json.dump({"mode": mode, "patches_per_launch": 500, "captured_at_commit": {"others": head}, "family_time_shares_ncu": shares, "sub_batch_us_ncu": round(tot, 1),
           "traffic_MB_per_launch": {d["capture"]: round(float(d["dram_rd_MB"]) + float(d["dram_wr_MB"]), 1) for d in summ}},
          open(out / f"{tag}_summary_{mode}.json", "w"), indent=1)
print(json.dumps(shares, indent=1)); print(tot)
