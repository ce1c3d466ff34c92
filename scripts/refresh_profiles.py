"""Turn the outputs of scripts/gpu_round2.sh (gpurun_out/) into the round-2 artefacts under profiles/:
bench lines, test log, launch list, the loop kernel's full capture + per-line shares + stage stamps, the decoder
kernels' capture, SASS evidence.   python scripts/refresh_profiles.py"""
import csv
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"


def sh(cmd, **kw):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True, cwd=ROOT, **kw).stdout


for src, dst in (("bench_ours.json", "r02_bench_n1.json"), ("bench_ref.json", "r02_bench_reference_arm.json"),
                 ("pytest_gpu.log", "r02_pytest_gpu.txt"), ("quick64.log", "r02_denoise_tc_quick64.txt")):
    if (G / src).is_file():
        shutil.copy(G / src, P / dst)
(P / "r02_launch_list.txt").write_text(
    "# ncu --metrics gpu__time_duration.sum --clock-control none -c 600  python bench.py --gpus 1 --steps 1 --warmup 1 "
    "--no-baselines   (final round-2 build)\n# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n"
    + sh("python scripts/launch_agg.py gpurun_out/launches.csv"))
(P / "r02_sass_evidence.txt").write_text(sh("python scripts/sass_evidence.py"))

# the loop kernel: full capture (same metric list as the committed file), per-line shares
rows = list(csv.reader(sh("ncu -i gpurun_out/prof_dtc.ncu-rep --page raw --csv").splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
old = (P / "r02_denoise_tc_full.txt").read_text().splitlines()
keys = [l.split()[0] for l in old if l and not l.startswith("#")]
head = [l for l in old if l.startswith("#")]
out = head + [""] + [f"{k:90s} {vals[hdr.index(k)]:>12s} {units[hdr.index(k)]}" for k in keys if k in hdr]
(P / "r02_denoise_tc_full.txt").write_text("\n".join(out) + "\n")
sh("ncu -i gpurun_out/prof_dtc.ncu-rep --page source --csv > /tmp/dtc_src.csv")
(P / "r02_denoise_tc_lines.txt").write_text(sh("python scripts/ncu_lines_tc.py /tmp/dtc_src.csv"))

# the decoder's kernels
rows = list(csv.reader(sh("ncu -i gpurun_out/prof_decode.ncu-rep --page raw --csv").splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
old = (P / "r02_decode_kernels_full.txt").read_text().splitlines()
head = [l for l in old if l.startswith("#")]
metrics = [l.split(" [")[0] for l in old if " = " in l and not l.startswith("#") and not l.startswith("kernels")]
ik = hdr.index("Kernel Name")
out = head + ["kernels: " + str([d[ik].split("(")[0][-40:] for d in data])]
out += [f"{m} [{units[hdr.index(m)]}] = " + " | ".join(d[hdr.index(m)] for d in data) for m in metrics if m in hdr]
(P / "r02_decode_kernels_full.txt").write_text("\n".join(out) + "\n")
for f in sorted(P.glob("r02_*")):
    print(f.name, f.stat().st_size)
