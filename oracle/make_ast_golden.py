"""Golden AST features at the SHIPPED depth (12 DeiT-base blocks x 3 branches) for one clip, computed by the CPU
restatement (oracle/ast_ref.py) in float64:

    python -m oracle.make_ast_golden        ->  tests/golden/ast_depth12_b1.npz      (a few minutes of CPU time)

TEST INFRASTRUCTURE (see oracle/__init__.py).  The GPU suite compares the engine with this fixture, so the driver-run
tests cover the configuration bench.py times (12 blocks) and not only the 1-2-block stacks that are cheap to evaluate on
the CPU inside a test.  Parity stays "unpinned" in the sense of SURVEY.md section 8c: timm 0.4.5 is not installable, the
restatement is cross-checked against HuggingFace ASTModel (tests/test_oracle.py)."""
from pathlib import Path

import numpy as np
import torch

from oracle import ast_ref as A
from oracle import weights as W


def main():
    torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
    sd = W.ast_state_dict(depth=12)
    fb = torch.randn(1, 1024, 128, generator=torch.Generator().manual_seed(12)) * 0.5
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        c64, e64, s64 = A.ast_features(sd64, fb.double())
        c32, e32, s32 = A.ast_features(sd, fb)
    out = dict(fbank_seed=np.array(12), fbank_probe=fb[0, [0, 511, 1023], :4].numpy(),
               weights_sha1=np.array(W.checksum({k: sd[k] for k in list(sd)[:8]})),
               con_f64=c64.numpy(), emo_f64=e64.numpy(), sty_f64=s64.numpy(),
               con_f32=c32.numpy(), emo_f32=e32.numpy(), sty_f32=s32.numpy())
    dst = Path(__file__).resolve().parents[1] / "tests" / "golden" / "ast_depth12_b1.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "fp32-vs-fp64 of the restatement:",
          max((c32.double() - c64).abs().max().item(), (e32.double() - e64).abs().max().item(), (s32.double() - s64).abs().max().item()))


if __name__ == "__main__":
    main()
