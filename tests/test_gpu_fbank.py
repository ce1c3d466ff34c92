"""On-device Kaldi fbank (+ pad + normalise) against torchaudio, the library the reference itself calls
(infer_ldm.py:182-190) -- so this row IS pinned by the reference's own dependency."""
import pytest
import torch

from oracle import ast_ref as A

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,seed", [(160000, 0), (159744, 1), (443117, 2), (8000, 3), (400, 4)])
def test_fbank_matches_torchaudio(engine, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 16000.0
    wav = 0.1 * torch.randn(1, n, generator=g) + 0.05 * torch.sin(2 * torch.pi * 440 * t)[None] + 0.01
    wav = wav - wav.mean()
    ref = A.fbank_features(wav)                                   # [1024,128] normalised, CPU torchaudio
    got = engine.fbank(wav)[0].cpu()
    err = (got - ref).abs().max().item()
    print(f"[parity] fbank n={n}: max|d| (normalised units) = {err:.3e}")
    assert got.shape == (1024, 128)
    assert err < 2e-4


def test_fbank_batch_and_channel(engine):
    g = torch.Generator().manual_seed(9)
    wav = 0.1 * torch.randn(3, 160000, generator=g)
    got = engine.fbank(wav).cpu()
    for b in range(3):
        assert (got[b] - A.fbank_features(wav[b:b + 1])).abs().max().item() < 2e-4
