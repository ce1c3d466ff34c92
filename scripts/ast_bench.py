"""AST (K5) timing + full-depth parity probe.  python scripts/ast_bench.py [B]   (AST_DEPTH=n for a shallower stack)"""
import os, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amuse_b200.engine import Engine          # noqa: E402
from oracle import weights as W, ast_ref as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
DEPTH = int(os.environ.get("AST_DEPTH", "12"))
sd = W.ast_state_dict(depth=DEPTH)
eng = Engine("cuda:0")
eng.load_state_dict("ast", sd)
eng.finalize()
fb = torch.randn(B, 1024, 128, generator=torch.Generator().manual_seed(0)) * 0.5
fbd = fb.cuda()
for _ in range(2):
    out = eng.ast_features(fbd)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = eng.ast_features(fbd); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(f"AST 3 branches depth {DEPTH}, B={B}: {ms:.1f} ms  -> {B * 783.08 * DEPTH / 12 / ms:.1f} TFLOP/s algorithmic, {B * 300 / ms * 1000:.0f} frames/s")
if "--parity" in sys.argv:
    torch.set_num_threads(16)
    t0 = time.time()
    ref = A.ast_features(sd, fb[:1])
    print(f"cpu oracle 1 clip: {time.time() - t0:.1f} s")
    ref64 = A.ast_features({k: v.double() for k, v in sd.items()}, fb[:1].double())
    for n, g, r, r64 in zip(("con", "emo", "sty"), out, ref, ref64):
        print(f"[parity] ast depth=12 {n}: |cuda-f64|={(g[:1].cpu().double() - r64).abs().max():.3e} |f32ref-f64|={(r.double() - r64).abs().max():.3e} |feat|max={r64.abs().max():.2f}")
