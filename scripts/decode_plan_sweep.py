"""Decode (K3+K4) time for 64 clips vs clips-per-pass x lanes.  python scripts/decode_plan_sweep.py"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amuse_b200.engine import Engine          # noqa: E402
from oracle import weights as W               # noqa: E402

eng = Engine("cuda:0")
eng.load_state_dict("vae", W.motionprior_state_dict())
eng.finalize()
z = torch.randn(64, 128, generator=torch.Generator().manual_seed(0)).cuda()
ref = None
for chunk, lanes in [(32, 1), (32, 2), (22, 3), (16, 2), (16, 4), (64, 1), (11, 4), (8, 4)]:
    eng._check(eng.lib.amuse_debug_set_decode_plan(eng._h, chunk, lanes))
    for _ in range(3):
        p, t = eng.decode(z)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        p, t = eng.decode(z)
    b.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = p.clone()
    print(f"clips/pass {chunk:3d} lanes {lanes}: {a.elapsed_time(b) / 10:.3f} ms   same={torch.equal(p, ref)}")
