"""Per-source-line share of executed instructions and warp-state samples of the tcgen05 sampler loop, from an
`ncu --set full --import-source on` capture:  ncu -i <rep> --page source --csv > src.csv ; python scripts/ncu_lines_tc.py src.csv
(joins the SASS page with `nvdisasm -g` of the same cubin extracted from the built library)."""
import collections
import csv
import glob
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), r[isrc], int(r[iex]), int(r[ismp])))
    except Exception:
        pass
base = min(d[0] for d in data)
tot, tots = sum(d[2] for d in data), sum(d[3] for d in data)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "denoise_tc", str(ROOT / "amuse_b200/lib/libamuse_b200.so")], cwd=tmp, capture_output=True)
cub = [c for c in glob.glob(tmp + "/*.cubin") if "denoise_tc.sm" in c or c.endswith("denoise_tc.sm_100a.cubin")] or glob.glob(tmp + "/*.cubin")
dis = subprocess.run(["nvdisasm", "-g", "-c", cub[0]], capture_output=True, text=True).stdout.splitlines()
line, amap = None, {}
inside = True   # the cubin holds two instantiations of the kernel (PROF = false / true): map the product one
for l in dis:
    ms = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if ms:
        inside = "denoise_tc_kernelILb0E" in ms.group(1)
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        line = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/", l)
    if m:
        amap[int(m.group(1), 16)] = line
agg, aggs = collections.Counter(), collections.Counter()
for a, src, ex, smp in data:
    ln = amap.get(a - base)
    agg[ln] += ex
    aggs[ln] += smp
src_lines = (ROOT / "amuse_b200/csrc/denoise_tc.cu").read_text().splitlines()
print(f"# {tot} warp-instructions executed, {tots} warp-state samples; share per source line (top 45 by samples)")
for ln, c in aggs.most_common(45):
    text = src_lines[ln[1] - 1].strip()[:90] if ln and ln[0] == "denoise_tc.cu" and ln[1] <= len(src_lines) else ""
    print(f"{str(ln[0]) + ':' + str(ln[1]) if ln else '?':28s} samples {c / tots * 100:5.1f}%  instr {agg[ln] / tot * 100:5.1f}%  | {text}")
