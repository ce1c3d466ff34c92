"""tcgen05 3xTF32 GEMM (csrc/tc_gemm.cu) against float64 torch on the shapes the decoder / AST use,
including ragged M / N tails, the two-source K concat and every fused epilogue."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

EPI_PLAIN, EPI_QKV, EPI_PLANES, EPI_GELU, EPI_RES_LN, EPI_RES_LN_CROSS_LN, EPI_RES = range(7)


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (300, 384, 128), (19200, 512, 128), (1000, 128, 512),
                                    (600, 333, 128), (130, 2304, 768), (1214, 768, 3072)])
def test_plain(engine, M, N, K):
    A, W, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=1 / math.sqrt(K)), _rand(N, seed=3)
    got = engine.debug_tc_gemm(EPI_PLAIN, A, W, b).cpu().double()
    ref = A.double() @ W.double().T + b.double()
    ref32 = (A @ W.T + b).double()
    err, err32 = (got - ref).abs().max().item(), (ref32 - ref).abs().max().item()
    print(f"[tc_gemm] {M}x{N}x{K}: 3xTF32 max|err|={err:.2e}  (fp32 torch CPU: {err32:.2e})")
    # tensor-core accumulation truncates (RZ) each partial sum, so the error grows ~linearly with K
    assert err < 4e-8 * K * max(1.0, ref.abs().max().item() / 4) + 1e-5


def test_concat_k(engine):
    M, N = 700, 128
    A, A2 = _rand(M, 128, seed=1), _rand(M, 128, seed=4)
    W, b = _rand(N, 256, seed=2, scale=1 / 16), _rand(N, seed=3)
    got = engine.debug_tc_gemm(EPI_PLAIN, A, W, b, A2=A2).cpu().double()
    ref = torch.cat([A, A2], 1).double() @ W.double().T + b.double()
    assert (got - ref).abs().max().item() < 2e-5


def test_epilogues(engine):
    M, K = 900, 128
    A, b = _rand(M, K, seed=1), _rand(128, seed=3)
    W = _rand(128, K, seed=2, scale=1 / math.sqrt(K))
    R, ln, cvec = _rand(M, 128, seed=5), torch.cat([1 + 0.1 * _rand(128, seed=6), 0.1 * _rand(128, seed=7),
                                                   1 + 0.1 * _rand(128, seed=8), 0.1 * _rand(128, seed=9)]), _rand(3, 128, seed=10)
    y = A.double() @ W.double().T + b.double()
    g1, b1, g2, b2 = (t.double() for t in ln.view(4, 128))
    # planes / gelu / residual
    assert (engine.debug_tc_gemm(EPI_PLANES, A, W, b).cpu().double() - y).abs().max().item() < 2e-5
    assert (engine.debug_tc_gemm(EPI_GELU, A, W, b).cpu().double() - F.gelu(y)).abs().max().item() < 2e-5
    assert (engine.debug_tc_gemm(EPI_RES, A, W, b, R=R).cpu().double() - (y + R.double())).abs().max().item() < 2e-5
    # residual + LayerNorm
    ln1 = F.layer_norm(y + R.double(), (128,), g1, b1, 1e-5)
    got = engine.debug_tc_gemm(EPI_RES_LN, A, W, b, R=R, ln=ln).cpu().double()
    assert (got - ln1).abs().max().item() < 3e-5
    # ... + cross vector + second LayerNorm (300 rows per clip)
    cv = cvec.double()[torch.arange(M) // 300]
    ln2 = F.layer_norm(ln1 + cv, (128,), g2, b2, 1e-5)
    got = engine.debug_tc_gemm(EPI_RES_LN_CROSS_LN, A, W, b, R=R, ln=ln, cvec=cvec, rows_per_clip=300).cpu().double()
    assert (got - ln2).abs().max().item() < 3e-5
    # q scaling of the packed in_proj
    W3, b3 = _rand(384, K, seed=11, scale=1 / math.sqrt(K)), _rand(384, seed=12)
    ref = A.double() @ W3.double().T + b3.double()
    ref[:, :128] *= 0.17677669529663687
    assert (engine.debug_tc_gemm(EPI_QKV, A, W3, b3).cpu().double() - ref).abs().max().item() < 2e-5


PAIR = 0x100   # route debug_tc_gemm to the CTA-pair kernel (csrc/tc_gemm2.cu)


@pytest.mark.parametrize("M,N,K", [(2424, 768, 256), (1214, 2304, 768), (300, 768, 3072), (19424, 256, 96)])
def test_pair_kernel_plain(engine, M, N, K):
    """cta_group::2 persistent GEMM: ragged M tails, several tiles per pair (double-buffered accumulators)."""
    A, W, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=1 / math.sqrt(K)), _rand(N, seed=3)
    got = engine.debug_tc_gemm(EPI_PLAIN | PAIR, A, W, b).cpu().double()
    ref = A.double() @ W.double().T + b.double()
    err = (got - ref).abs().max().item()
    print(f"[tc_gemm2] {M}x{N}x{K}: 3xTF32 max|err|={err:.2e}")
    assert err < 4e-8 * K * max(1.0, ref.abs().max().item() / 4) + 1e-5


def test_pair_kernel_epilogues(engine):
    M, N, K = 40000, 768, 128      # 157 row tiles x 3: more tiles than CTA pairs -> persistent loop + both buffers
    A, W, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=1 / math.sqrt(K)), _rand(N, seed=3)
    R = _rand(M, N, seed=5)
    y = A.double() @ W.double().T + b.double()
    assert (engine.debug_tc_gemm(EPI_GELU | PAIR, A, W, b).cpu().double() - F.gelu(y)).abs().max().item() < 2e-5
    assert (engine.debug_tc_gemm(EPI_RES | PAIR, A, W, b, R=R).cpu().double() - (y + R.double())).abs().max().item() < 2e-5
