"""Short workloads for `ncu --set full` captures (one kernel replayed ~40x: keep them small)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amuse_b200.engine import Engine          # noqa: E402
from oracle import weights as W              # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "denoise"
eng = Engine("cuda:0")
eng.load_state_dict("denoiser", W.denoiser_state_dict())
eng.load_state_dict("vae", W.motionprior_state_dict())
eng.finalize()
g = torch.Generator().manual_seed(0)
B = 64
l0, con, emo, sty = (torch.randn(B, d, generator=g).cuda() for d in (128, 256, 256, 256))
if what == "denoise":
    z = eng.denoise(l0, con, emo, sty, n_steps=50, sampler="ddim")       # the shipped 50-step schedule, B=64
else:
    z = eng.denoise(l0, con, emo, sty, n_steps=2, sampler="ddim")
    eng.decode(z)
torch.cuda.synchronize()
print("done", what)
