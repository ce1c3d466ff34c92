"""Small workloads for compute-sanitizer (scripts/gpu_sanitize.sh): every kernel family of the library once, sized so
that memcheck / synccheck / racecheck finish in minutes.   python scripts/sanitize_target.py denoise|decode|ast|fbank"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amuse_b200.engine import Engine          # noqa: E402
from oracle import weights as W              # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "denoise"
eng = Engine("cuda:0")
g = torch.Generator().manual_seed(0)
if what in ("denoise", "decode"):
    eng.load_state_dict("denoiser", W.denoiser_state_dict())
    eng.load_state_dict("vae", W.motionprior_state_dict())
    eng.finalize()
    B = 5
    l0, con, emo, sty = (torch.randn(B, d, generator=g).cuda() for d in (128, 256, 256, 256))
    if what == "denoise":
        z = eng.denoise(l0, con, emo, sty, n_steps=10, sampler="ddpm", seed=3)        # 5 tokens, in-kernel Philox
        z3 = eng.denoise(l0, con, None, None, n_steps=4, sampler="ddim")               # 3-token ablation
        print("latents", float(z.abs().max()), float(z3.abs().max()))
    else:
        poses, trans = eng.decode(l0)
        print("poses", float(poses.abs().max()))
        feats = eng.motion_to_feats(poses[:2].contiguous(), trans[:2].contiguous())   # the encoder: 302-token attention
        mu, logvar = eng.encode(feats)
        print("mu", float(mu.abs().max()))
elif what == "ast":
    eng.load_state_dict("ast", W.ast_state_dict(depth=1))
    eng.finalize()
    fb = torch.randn(2, 1024, 128, generator=g).cuda() * 0.5
    con, emo, sty = eng.ast_features(fb)
    print("ast", float(con.abs().max()))
elif what == "fbank":
    eng.load_state_dict("denoiser", W.denoiser_state_dict())
    eng.load_state_dict("vae", W.motionprior_state_dict())
    eng.finalize()
    fb = eng.fbank(0.1 * torch.randn(2, 160000, generator=g).cuda())
    print("fbank", float(fb.abs().max()))
torch.cuda.synchronize()
print("done", what)
