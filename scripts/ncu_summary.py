"""Turn gpurun_out/*.ncu-rep + launches.csv into the small text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles"
G = ROOT / "gpurun_out"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_active.avg"]


def launches():
    f = G / "launches.csv"
    if not f.is_file():
        return
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
        k = d["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(OUT / f"{TAG}_launch_list.txt", "w") as o:
        o.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400  python bench.py --gpus 1 --steps 1 --warmup 1\n")
        o.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"{v[1] / 1e6:10.3f} ms  {100 * v[1] / tot:5.1f}%  launches={v[0]:4d}  {k}\n")
        o.write(f"total {tot / 1e6:.3f} ms\n")


def full(rep, name, n_show=1):
    f = G / rep
    if not f.is_file():
        return
    raw = subprocess.run(["ncu", "-i", str(f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(OUT / f"{TAG}_{name}.txt", "w") as o:
        o.write(f"# ncu --set full --clock-control none --import-source on ... ({rep})\n")
        for vals in rows[2:2 + n_show]:
            d = dict(zip(hdr, vals))
            o.write(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
            for i, h in enumerate(hdr):
                if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                    o.write(f"{h} = {vals[i]} {units[i]}\n")
            o.write("\n")


OUT.mkdir(exist_ok=True)
launches()
full("prof_denoise.ncu-rep", "denoise_loop_full")
full("prof_decode_gemm.ncu-rep", "decode_gemm_full", n_show=3)
full("prof_tc_gemm_ast.ncu-rep", "tc_gemm_ast_full", n_show=4)
for p in sorted(OUT.glob(f"{TAG}_*")):
    print(p.name, p.stat().st_size)
