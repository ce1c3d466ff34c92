"""Attribute the warp-state samples of an `ncu --set full --import-source on` capture of the denoise loop to the
source lines of the KERNEL BODY (outermost inlining frame), so that every stage of the step gets its share.

    python scripts/ncu_lines.py gpurun_out/prof_denoise.ncu-rep [mangled-kernel-substring] [inside-line]

ncu's CSV source page has no per-line view for inlined code, so the SASS page (address -> samples per stall
reason) is joined with `nvdisasm -gi` of the same cubin (address -> inlining chain).  All 10 warps of a CTA are
resident for the whole launch, so a line's share of the samples is (warps inside it) x (time) / (10 x total).
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "denoise_loop_kernelILi2ELb1ELb0"
inside = int(sys.argv[3]) if len(sys.argv) > 3 else 0
SRC = ROOT / "amuse_b200" / "csrc" / "denoise_loop.cu"

tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "denoise_loop", str(ROOT / "amuse_b200/lib/libamuse_b200.so")], cwd=tmp, check=True,
               capture_output=True)
cubin = next(tmp.glob("*.cubin"))
dis = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
addr2, pending, cur, on = {}, [], [], False
for l in dis:
    if l.startswith(".text."):
        on = kname in l
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        if pending:
            cur, pending = list(pending), []
        addr2[int(m.group(1), 16)] = (cur, m.group(2))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ix["Address"]], 16)
    per[a]["samples"] += int(r[ix["# Samples"]] or 0)
    per[a]["inst"] += int(r[ix["Instructions Executed"]] or 0)
    for c in stall_cols:
        if r[ix[c]]:
            per[a][c] += int(r[ix[c]])
addrs = sorted(per)
base = addrs[0]
src = SRC.read_text().splitlines()
agg = collections.defaultdict(collections.Counter)
for a in addrs:
    info = addr2.get(a - base)
    if info is None or not info[0]:
        agg[(-1, "?")].update(per[a])
        continue
    chain = info[0]
    if not chain[-1][0].endswith("denoise_loop.cu"):
        agg[(-1, "?")].update(per[a])
        continue
    if inside:
        if chain[-1][1] != inside:
            continue
        k = chain[-2] if len(chain) > 1 else chain[-1]
        agg[(k[1], Path(k[0]).name)].update(per[a])
    else:
        agg[(chain[-1][1], "denoise_loop.cu")].update(per[a])
total = sum(v["samples"] for v in per.values())
print(f"# {rep}: {total} warp-state samples, kernel *{kname}*" + (f", inside the call at line {inside}" if inside else ""))
print("# line  share  warp-instructions  top stall reasons | source")
for (line, f), c in sorted(agg.items()):
    if c["samples"] < total * 0.002:
        continue
    top = sorted(((k, v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    reasons = " ".join(f"{k[6:]}={100 * v / max(1, c['samples']):.0f}%" for k, v in top)
    text = src[line - 1].strip()[:88] if (line > 0 and f == "denoise_loop.cu") else f
    print(f"{line:5d} {100 * c['samples'] / total:5.1f}%  {c['inst']:10d}  {reasons:44s} | {text}")
