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
