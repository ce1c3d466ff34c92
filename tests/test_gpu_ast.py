"""AST encoders (a13) and process_single_seq (a12) on the GPU against the CPU restatement
(oracle/ast_ref.py; parity unpinned -- timm is not installable, see the oracle header)."""
import numpy as np
import pytest
import torch

from oracle import ast_ref as A
from oracle import weights as W

pytestmark = pytest.mark.gpu


def _engine_with_ast(depth):
    from amuse_b200.engine import Engine
    sd = W.ast_state_dict(depth=depth)
    eng = Engine("cuda:0")
    eng.load_state_dict("ast", sd)
    eng.finalize()
    return eng, sd


@pytest.mark.parametrize("depth,B", [(1, 3), (2, 2)])
def test_ast_features_vs_oracle(depth, B):
    eng, sd = _engine_with_ast(depth)
    fb = torch.randn(B, 1024, 128, generator=torch.Generator().manual_seed(depth)) * 0.5
    con, emo, sty = eng.ast_features(fb)
    torch.set_num_threads(min(16, torch.get_num_threads()))
    rc, re, rs = A.ast_features(sd, fb)
    sd64 = {k: v.double() for k, v in sd.items()}
    dc, de, ds = A.ast_features(sd64, fb.double())
    for name, got, ref, ref64 in (("con", con, rc, dc), ("emo", emo, re, de), ("sty", sty, rs, ds)):
        e64 = (got.cpu().double() - ref64).abs().max().item()
        r = (ref.double() - ref64).abs().max().item()
        print(f"[parity] ast depth={depth} {name}: |cuda-f64|={e64:.3e} |f32ref-f64|={r:.3e} |feat|max={ref64.abs().max():.2f}")
        assert e64 < 3e-4      # measured 4-5e-5 (tensor-core RZ accumulation over K = 768..3072)
    eng.close()


def test_process_single_seq_shapes_and_fbank():
    """The mirror class end to end on a synthetic 10 s / 16 kHz waveform: host kaldi fbank (as the
    reference), normalisation after zero padding, AST on the device."""
    from amuse_b200.infer_ldm import PretrainedLPDM_v1
    den, vae, ast = W.denoiser_state_dict(), W.motionprior_state_dict(), W.ast_state_dict(depth=1)
    m = PretrainedLPDM_v1.from_state_dicts(den, vae, ast, device="cuda:0")
    wav = 0.1 * torch.randn(1, 160000, generator=torch.Generator().manual_seed(0))
    con, emo, sty = m.process_single_seq(wav - wav.mean(), framerate=16000)
    assert con.shape == emo.shape == sty.shape == (1, 256) and con.is_cuda
    fb = A.fbank_features(wav - wav.mean())
    assert fb.shape == (1024, 128) and abs(float(fb[1000:, :].mean()) - 0.906) < 1e-2     # padded rows after normalisation
    rc, re, rs = A.ast_features(ast, fb[None])
    assert (con.cpu() - rc).abs().max().item() < 3e-4
    out = m.diffusion_backward(1, con, emo, sty)
    assert out["poses"].shape == (1, 300, 55, 3) and torch.isfinite(out["poses"]).all()
    m.engine.close()
