"""A/B of the interleaved MMA issue (AMUSE_DN2_DEBUG bits 8..10): loop time at B=64 x DDPM-1000 and the difference of the
latents against the default order.  Each variant runs in its own process (the flag is read per launch, but keep it clean)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from amuse_b200.engine import Engine
from oracle import weights as W
eng = Engine("cuda:0")
eng.load_state_dict("denoiser", W.denoiser_state_dict()); eng.load_state_dict("vae", W.motionprior_state_dict()); eng.finalize()
g = torch.Generator().manual_seed(0)
B = 64
l0, con, emo, sty = (torch.randn(B, d, generator=g).cuda() for d in (128, 256, 256, 256))
ts = []
for i in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); z = eng.denoise(l0, con, emo, sty, n_steps=1000, sampler="ddpm", seed=7); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
z50 = eng.denoise(l0, con, emo, sty, n_steps=50, sampler="ddim")
torch.save({"z": z.cpu(), "z50": z50.cpu()}, sys.argv[1])
print("loop ms:", " ".join("%%.2f" %% t for t in ts), " median %%.2f" %% sorted(ts)[2])
''' % str(ROOT)

ref = None
import torch
for flags in (0, 1024, 0, 1024):
    env = dict(os.environ, AMUSE_DN2_DEBUG=str(flags))
    out = f"/tmp/ab_{flags}.pt"
    r = subprocess.run([sys.executable, "-c", CHILD, out], env=env, capture_output=True, text=True, timeout=300)
    print(f"flags={flags}: {r.stdout.strip()} {r.stderr.strip()[-300:] if r.returncode else ''}", flush=True)
    if r.returncode == 0:
        d = torch.load(out)
        if ref is None:
            ref = d
        else:
            print("   max|dz| ddpm1000 %.3e  ddim50 %.3e" % ((d["z"] - ref["z"]).abs().max(), (d["z50"] - ref["z50"]).abs().max()))
