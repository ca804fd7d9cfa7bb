"""SASS evidence that the kernels are Blackwell-native: per kernel of libmermaid_b200.so, the count of tcgen05 MMA (UTC*MMA),
TMEM load/store (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP), packed FP32 (FFMA2) and legacy tensor (HMMA) instructions.
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, re, subprocess, sys
from pathlib import Path
lib = Path(__file__).resolve().parents[1] / "mermaid_classifier_b200" / "libmermaid_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
pats = {"UTCMMA(tcgen05.mma)": r"\bUTC[A-Z]*MMA", "LDTM(tcgen05.ld)": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG(TMA load)": r"\bUTMALDG",
        "UTMASTG/UBLKCP": r"\bUTMASTG|\bUBLKCP", "SYNCS(mbarrier)": r"\bSYNCS", "FFMA2": r"\bFFMA2", "MUFU": r"\bMUFU", "HMMA(legacy)": r"\bHMMA"}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        if ">" in subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout:
            full = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            i = full.find(">(")
            cur = full[: i + 1].replace("void ", "") if i >= 0 else cur
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {lib.name}: instruction counts per kernel (sm_100a)")
print("kernel," + ",".join(pats))
tot = collections.Counter()
for k, c in counts.items():
    print(f"\"{k}\"," + ",".join(str(c[p]) for p in pats))
    tot.update(c)
print("\"TOTAL\"," + ",".join(str(tot[p]) for p in pats))
