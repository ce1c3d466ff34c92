"""Same-box A/B of two builds of the library (scripts/build_variant.sh NAME): loop time at B=64 x DDPM-1000 and the
difference of the latents.   python scripts/experiments/ab_lib.py prev [other ...]   ('' = the product library)"""
import os
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
CHILD = (ROOT / "scripts" / "experiments" / "ab_ilv.py").read_text().split("CHILD = r'''")[1].split("''' % str(ROOT)")[0] % str(ROOT)
names = [""] + sys.argv[1:]
ref = None
for name in names + names:
    env = dict(os.environ)
    if name:
        env["AMUSE_B200_LIB"] = str(ROOT / f"amuse_b200/lib/libamuse_b200_{name}.so")
    out = f"/tmp/ab_{name or 'product'}.pt"
    r = subprocess.run([sys.executable, "-c", CHILD, out], env=env, capture_output=True, text=True, timeout=300)
    print(f"lib={name or 'product'}: {r.stdout.strip()} {r.stderr.strip()[-300:] if r.returncode else ''}", flush=True)
    if r.returncode == 0:
        d = torch.load(out)
        if ref is None:
            ref = d
        else:
            print("   max|dz| ddpm1000 %.3e  ddim50 %.3e" % ((d["z"] - ref["z"]).abs().max(), (d["z50"] - ref["z50"]).abs().max()))
