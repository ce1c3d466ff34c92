"""CPU check of the branch-free erf used by the kernels' GELU (amuse_b200/csrc/common.cuh
erf_fast): the same coefficients evaluated in float32 numpy against float64 scipy."""
import numpy as np
from scipy.special import erf


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def erf_fast_np(a):
    a = a.astype(np.float32)
    t, s = np.abs(a), (a * a).astype(np.float32)
    c = lambda v: np.full_like(a, v, dtype=np.float32)
    r = _fma(c(-1.72853470e-5), t, c(3.83197126e-4))
    u = _fma(c(-3.88396438e-3), t, c(2.42546219e-2))
    r = _fma(r, s, u)
    for k in (-1.06777877e-1, -6.34846687e-1, -1.28717512e-1):
        r = _fma(r, t, c(k))
    r = _fma(r, t, -t)
    big = np.copysign((1.0 - np.exp(r.astype(np.float64))).astype(np.float32), a)
    q = c(-5.96761703e-4)
    for k in (4.99119423e-3, -2.67681349e-2, 1.12819925e-1, -3.76125336e-1, 1.28379166e-1):
        q = _fma(q, s, c(k))
    small = _fma(q, a, a)
    return np.where(t > 0.927734375, big, small)


def test_erf_fast_is_sub_ulp():
    rng = np.random.default_rng(0)
    x = np.concatenate([np.linspace(-6, 6, 400001), rng.standard_normal(200000) * 1.5]).astype(np.float32)
    ref = erf(x.astype(np.float64))
    err = np.abs(erf_fast_np(x).astype(np.float64) - ref)
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    assert err.max() < 1e-7
    assert (err / ulp).max() < 1.5


def test_check_poses_is_conditioning_aware():
    """oracle/lpdm_ref.py::check_poses (used by smoke() and the setup test): the feature error explains the pose error of
    ill-conditioned 6D pairs, a well-conditioned rotation that is off by more than the tolerance fails."""
    import torch
    from oracle import lpdm_ref as R
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(2, 4, 333, generator=g)
    feats[0, 0, :6] *= 0.01                                     # one ill-conditioned joint (|a| ~ 0.01)
    ref = R.feats_to_motion(feats)
    ok, geo, well = R.check_poses(ref["poses"], ref["poses"], feats, 1e-6)
    assert ok and geo < 1e-3 and well < 1e-3
    noisy = feats + 1e-5 * torch.randn(feats.shape, generator=g)
    got = R.feats_to_motion(noisy)
    ok, geo, well = R.check_poses(got["poses"], ref["poses"], feats, 4e-5)
    assert ok and well < 0.05
    bad = feats.clone()
    bad[1, 1, 6:12] += 0.05                                     # a well-conditioned joint moved by degrees
    got = R.feats_to_motion(bad)
    assert not R.check_poses(got["poses"], ref["poses"], feats, 1e-5)[0]
