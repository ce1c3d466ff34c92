"""Static evidence from the built library (no GPU): per kernel, registers / spills from `cuobjdump -res-usage`
and the count of the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, STAS = st.async (DSMEM),
FFMA2 = packed fp32 FMA, HMMA = legacy mma.sync.   python scripts/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "amuse_b200" / "lib" / "libamuse_b200.so"
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                               ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("STAS", r"\bSTAS"), ("FFMA2", r"\bFFMA2"),
                               ("FFMA", r"\bFFMA\b"), ("HMMA", r"\bHMMA"), ("BAR.SYNC", r"\bBAR\.SYNC"),
                               ("UCGABAR", r"\bUCGABAR")])


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n


res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
usage = {}
cur = None
for l in res.splitlines():
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in l:
        usage[cur] = l.strip()
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for l in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        counts[cur]["instr"] += 1
        for k, p in PAT.items():
            if re.search(p, l):
                counts[cur][k] += 1
print(f"# {LIB.relative_to(ROOT)}: {len(counts)} kernels (nvcc -gencode arch=compute_100a,code=sm_100a)")
for fn, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(fn).replace("(anonymous namespace)::", ""))
    print(f"\n{name}")
    print(f"  {usage.get(fn, '')}")
    print("  instr=%d  " % c["instr"] + "  ".join(f"{k}={c[k]}" for k in PAT if c[k]))
