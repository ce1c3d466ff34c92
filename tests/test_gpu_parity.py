"""GPU parity tests: the CUDA path (through the C ABI) against
  (1) the committed golden vectors produced by the REFERENCE'S OWN modules (oracle/make_golden.py),
  (2) the CPU oracle (oracle/lpdm_ref.py) run live on odd batch sizes / ablations.

Tolerances.  All arithmetic is fp32.  The yardstick is the fp64 run of the same reference
modules: the reference's own fp32 path is `ref_err` away from it, and the CUDA path must be within
max(ABS_FLOOR, K * ref_err) of the fp64 result and within TOL32 of the reference fp32 result.
"""
import numpy as np
import pytest
import os

import torch

from oracle import lpdm_ref as R
from oracle import weights as W

pytestmark = pytest.mark.gpu

TOL_EPS = 2e-5          # one Denoiser.forward, |eps| ~ 3
TOL_LAT_DDIM = 2e-4     # 50 recurrent steps with clamp; reference fp32-vs-fp64 itself is ~3e-5
TOL_FEATS = 2e-4        # 6D features, |feats| ~ 3
TOL_GEO_DEG = 0.05      # rotation geodesic, degrees


def _t(a):
    return torch.from_numpy(np.asarray(a))


_gram_schmidt_cond = R.gram_schmidt_cond


def _assert_poses(name, poses, poses32, poses64, feats64, feats_err):
    """Pose parity through the rotation geodesic.  A rotation whose 6D vectors are short is ill-conditioned
    (synthetic weights produce |a| down to ~0.01), so the bound per rotation is the allowed 6D feature error
    carried through the conditioning of Gram-Schmidt; well-conditioned rotations must meet TOL_GEO_DEG."""
    geo = R.geodesic_deg(poses.cpu(), _t(poses64))
    geo_ref = R.geodesic_deg(_t(poses32), _t(poses64))          # the reference's own fp32 path vs fp64
    cond = _gram_schmidt_cond(_t(feats64))
    bound = torch.rad2deg(3.0 * feats_err * 6 ** 0.5 * cond) + 0.01
    worst = int(geo.argmax())
    print(f"[parity] {name} poses geodesic max {geo.max().item():.4f} deg at conditioning {cond.flatten()[worst].item():.1f} "
          f"(reference fp32 vs fp64: {geo_ref.max().item():.4f}); well-conditioned max "
          f"{geo[cond < 4].max().item():.4f} deg")
    assert bool((geo <= bound).all()), "pose error not explained by the 6D feature error"
    assert geo[cond < 4].max().item() < max(TOL_GEO_DEG, 3 * geo_ref[cond < 4].max().item())


def _report(name, got, ref32, ref64):
    got = got.double().cpu()
    e64 = (got - _t(ref64).double()).abs().max().item()
    e32 = (got - _t(ref32).double()).abs().max().item()
    r = (_t(ref32).double() - _t(ref64).double()).abs().max().item()
    print(f"[parity] {name}: |cuda-f64|={e64:.3e} |cuda-f32ref|={e32:.3e} |f32ref-f64|={r:.3e}")
    return e64, e32, r


def test_weights_regenerate(golden_dir, synthetic_weights):
    g = np.load(golden_dir / "denoiser_step.npz")
    assert W.checksum(synthetic_weights["denoiser"]) == str(g["denoiser_sha1"])
    assert W.checksum(synthetic_weights["vae"]) == str(g["motionprior_sha1"])


def test_denoiser_eps_golden(engine, golden_dir):
    g = np.load(golden_dir / "denoiser_step.npz")
    x, con, emo, sty = (_t(g[k]) for k in ("x", "con", "emo", "sty"))
    for t in (981, 500, 1):
        got = engine.denoiser_eps(x, t, con, emo, sty)
        e64, e32, r = _report(f"eps t={t}", got, g[f"eps_t{t}_f32"], g[f"eps_t{t}_f64"])
        assert e64 < max(TOL_EPS, 4 * r) and e32 < TOL_EPS
    got = engine.denoiser_eps(x, 981, con, None, None)
    e64, e32, r = _report("eps no emo/sty", got, g["eps_t981_noemo_nosty_f32"], g["eps_t981_noemo_nosty_f64"])
    assert e64 < max(TOL_EPS, 4 * r) and e32 < TOL_EPS
    got = engine.denoiser_eps(x, 981, con, emo, None)
    e64, e32, r = _report("eps no sty", got, g["eps_t981_nosty_f32"], g["eps_t981_nosty_f64"])
    assert e64 < max(TOL_EPS, 4 * r) and e32 < TOL_EPS


@pytest.mark.parametrize("name", ["ddim50_b4", "ddim1_b1", "ddim50_b1"])
def test_ddim_golden(engine, golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    got = engine.denoise(_t(g["latents0"]), _t(g["con"]), _t(g["emo"]), _t(g["sty"]), n_steps=int(g["n_steps"]),
                         sampler="ddim")
    e64, e32, r = _report(name, got, g["z_f32"], g["z_f64"])
    assert e64 < max(TOL_LAT_DDIM, 4 * r) and e32 < TOL_LAT_DDIM


@pytest.mark.parametrize("name", ["ddpm100_b2", "ddpm1000_b2"])
def test_ddpm_golden(engine, golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    n, B = int(g["n_steps"]), g["latents0"].shape[0]
    noise = torch.randn(n, B, 128, generator=torch.Generator().manual_seed(int(g["noise_seed"])))
    if "noise_probe" in g:
        assert torch.equal(noise[[0, 499, 999]], _t(g["noise_probe"])), "torch CPU RNG stream differs from the fixture's"
    got = engine.denoise(_t(g["latents0"]), _t(g["con"]), _t(g["emo"]), _t(g["sty"]), n_steps=n, sampler="ddpm",
                         step_noise=noise)
    e64, e32, r = _report(name, got, g["z_f32"], g["z_f64"])
    scale = float(np.abs(g["z_f64"]).max())          # unclipped ancestral chains of a random net grow large
    assert e64 < max(2e-5 * scale, 4 * r) and e32 < max(2e-5 * scale, 4 * r)


def test_decode_golden(engine, golden_dir):
    g = np.load(golden_dir / "decode_b2.npz")
    idx = g["frame_idx"]
    poses, trans, feats = engine.decode(_t(g["z"]), want_feats=True)
    e64, e32, r = _report("decode feats", feats[:, idx], g["feats_f32"], g["feats_f64"])
    assert e64 < max(TOL_FEATS, 4 * r) and e32 < TOL_FEATS
    geo = R.geodesic_deg(poses[:, idx].cpu(), _t(g["poses_f64"]))
    print(f"[parity] decode poses geodesic max {geo.max().item():.4f} deg")
    assert geo.max().item() < TOL_GEO_DEG
    assert torch.equal(trans.cpu(), feats[:, :, -3:].cpu())


def test_backward_golden(engine, golden_dir):
    g = np.load(golden_dir / "backward_ddim50_b2.npz")
    idx = g["frame_idx"]
    out = engine.diffusion_backward(_t(g["latents0"]), _t(g["con"]), _t(g["emo"]), _t(g["sty"]), n_steps=50,
                                    sampler="ddim", want_latents=True, want_feats=True)
    e64, e32, r = _report("backward latents", out["latents"], g["z_f32"], g["z_f64"])
    assert e64 < max(TOL_LAT_DDIM, 4 * r)
    e64, e32, r = _report("backward feats", out["feats"][:, idx], g["feats_f32"], g["feats_f64"])
    assert e64 < max(TOL_FEATS, 4 * r)
    _assert_poses("backward", out["poses"][:, idx], g["poses_f32"], g["poses_f64"], g["feats_f64"], e64)
    # SMPL-X pose L2 (north_star): on the 6D rotation features, per value RMS
    l2 = (out["feats"][:, idx].double().cpu() - _t(g["feats_f64"])).pow(2).mean().sqrt().item()
    print(f"[parity] backward 6D feats RMS error {l2:.3e}")
    assert l2 < 1e-4


def test_rot6d_cases(engine, golden_dir):
    g = np.load(golden_dir / "rot6d_cases.npz")
    aa = engine.rot6d_to_axis_angle(_t(g["d6"])).cpu()
    geo = R.geodesic_deg(aa, _t(g["aa_f64"]))
    print(f"[parity] rot6d geodesic max {geo.max().item():.5f} deg")
    assert geo.max().item() < 0.02
    d = (aa - _t(g["aa_f32"])).abs()
    ok = d.max(dim=-1).values < 1e-4          # identical up to 2*pi-wrap sign flips near pi
    assert ok.float().mean().item() > 0.9


@pytest.mark.parametrize("B,emo,sty", [(1, True, True), (3, True, True), (5, True, False), (7, False, False),
                                        (13, True, True), (16, True, True), (17, True, True), (64, True, True)])
def test_ddim_vs_oracle_live(engine, synthetic_weights, B, emo, sty):
    """Odd batch sizes exercise every clips-per-cluster split (1..4 clips, uneven halves, partial
    last cluster); None conditions exercise 3- and 4-token sequences."""
    g = torch.Generator().manual_seed(1000 + B)
    l0, con = torch.randn(B, 128, generator=g), torch.randn(B, 256, generator=g)
    ze = torch.randn(B, 256, generator=g) if emo else None
    zs = torch.randn(B, 256, generator=g) if sty else None
    n = 10
    ref = R.sample_latents(synthetic_weights["denoiser"], l0, con, ze, zs, n, "ddim")
    got = engine.denoise(l0, con, ze, zs, n_steps=n, sampler="ddim").cpu()
    err = (got - ref).abs().max().item()
    print(f"[parity] live ddim10 B={B} emo={emo} sty={sty}: max|d|={err:.3e}")
    assert err < 1e-4


def test_eta_and_philox(engine, synthetic_weights):
    """DDIM with eta > 0 against the oracle with injected noise; Philox path: reproducible,
    seed-dependent, and independent of how clips are packed into clusters."""
    B, n, eta = 6, 8, 0.5
    g = torch.Generator().manual_seed(77)
    l0, con = torch.randn(B, 128, generator=g), torch.randn(B, 256, generator=g)
    ze, zs = torch.randn(B, 256, generator=g), torch.randn(B, 256, generator=g)
    noise = torch.randn(n, B, 128, generator=g)
    ts, coef = engine.schedule(n, "ddim", eta)
    plan = {"timesteps": ts, "coef": coef, "clip": True, "sampler": "ddim"}
    x = l0
    for i, t in enumerate(ts):
        e = R.denoiser_forward(synthetic_weights["denoiser"], x, t, con, ze, zs)
        x = R.scheduler_step(plan, i, x, e, None) + coef[i, 4] * noise[i]
    got = engine.denoise(l0, con, ze, zs, n_steps=n, sampler="ddim", eta=eta, step_noise=noise).cpu()
    assert (got - x).abs().max().item() < 1e-4
    a = engine.denoise(l0, con, ze, zs, n_steps=n, sampler="ddpm", seed=5).cpu()
    b = engine.denoise(l0, con, ze, zs, n_steps=n, sampler="ddpm", seed=5).cpu()
    c = engine.denoise(l0, con, ze, zs, n_steps=n, sampler="ddpm", seed=6).cpu()
    assert torch.equal(a, b) and not torch.equal(a, c)
    a2 = engine.denoise(l0[:2], con[:2], ze[:2], zs[:2], n_steps=n, sampler="ddpm", seed=5).cpu()
    assert torch.allclose(a[:2], a2, atol=1e-5)


def test_schedule_matches_oracle(engine):
    for n in (1, 50):
        ts, coef = engine.schedule(n, "ddim")
        plan = R.ddim_coeffs(n)
        assert ts == plan["timesteps"]
        print("[parity] ddim coef max rel diff", ((coef - plan["coef"]).abs() / plan["coef"].abs().clamp_min(1e-30)).max().item())
        assert torch.allclose(coef, plan["coef"], rtol=5e-7, atol=0)
    for n in (100, 1000):
        ts, coef = engine.schedule(n, "ddpm")
        plan = R.ddpm_coeffs(n)
        assert ts == plan["timesteps"]
        print("[parity] ddpm coef max rel diff", ((coef - plan["coef"]).abs() / plan["coef"].abs().clamp_min(1e-30)).max().item())
        assert torch.allclose(coef, plan["coef"], rtol=5e-7, atol=0)
    with pytest.raises(Exception):
        engine.schedule(1000, "ddim")          # alphas_cumprod[1000]: IndexError in the reference too


def test_host_entry_point(engine, golden_dir):
    g = np.load(golden_dir / "backward_ddim50_b2.npz")
    idx = g["frame_idx"]
    out = engine.diffusion_backward_host(_t(g["latents0"]).pin_memory(), _t(g["con"]).pin_memory(),
                                         _t(g["emo"]).pin_memory(), _t(g["sty"]).pin_memory(), n_steps=50)
    _assert_poses("host entry point", out["poses"][:, idx], g["poses_f32"], g["poses_f64"], g["feats_f64"], TOL_FEATS)


def test_decode_chunking_and_full_size(engine, synthetic_weights):
    """B larger than the decoder's chunk (32 clips) and not a multiple of it; compared clip-wise
    with a B=1 run (batch invariance: every clip is independent)."""
    B = 70
    z = torch.randn(B, 128, generator=torch.Generator().manual_seed(5))
    poses, trans, feats = engine.decode(z, want_feats=True)
    for b in (0, 31, 32, 69):
        p1, t1, f1 = engine.decode(z[b:b + 1], want_feats=True)
        assert torch.equal(f1[0], feats[b])
    ref = R.vae_decode(synthetic_weights["vae"], z[69:70])
    assert (feats[69].cpu() - ref[0]).abs().max().item() < TOL_FEATS


def test_encode_golden(engine, golden_dir):
    """axis-angle -> 6D features -> MotionPrior.encode (vae.py:154-214) against the
    reference's own modules (tests/golden/encode_b2.npz, oracle/make_golden.py G8) on closed-form motion."""
    from oracle.make_golden import synthetic_motion
    g = np.load(golden_dir / "encode_b2.npz")
    idx = g["frame_idx"]
    poses, trans = synthetic_motion(2)
    feats = engine.motion_to_feats(poses, trans)
    e = (feats[:, idx].cpu().double() - _t(g["feats_f64"])).abs().max().item()
    r = (_t(g["feats_f32"]).double() - _t(g["feats_f64"])).abs().max().item()
    print(f"[parity] motion_to_feats: |cuda-f64|={e:.3e}  |ref32-f64|={r:.3e}")
    assert e < max(2e-6, 4 * r)
    mu, logvar = engine.encode(feats)
    std = logvar.exp().pow(0.5)                                   # vae.py:210
    e_mu, _, r_mu = _report("encode mu", mu, g["mu_f32"], g["mu_f64"])
    e_sd, _, r_sd = _report("encode std", std, g["std_f32"], g["std_f64"])
    assert e_mu < max(TOL_FEATS, 4 * r_mu) and e_sd < max(2 * TOL_FEATS, 4 * r_sd)   # |mu|, |std| up to ~4


def test_encode_batch_and_chunking(engine, synthetic_weights):
    """40 clips (two decoder-sized passes of 32 + 8) against the oracle; clip order must not matter."""
    from oracle.make_golden import synthetic_motion
    poses, trans = synthetic_motion(40)
    feats = engine.motion_to_feats(poses, trans)
    mu, logvar = engine.encode(feats)
    ref_mu, ref_lv = R.vae_encode(synthetic_weights["vae"], R.motion_to_feats(poses[[0, 33, 39]], trans[[0, 33, 39]]))
    assert (mu[[0, 33, 39]].cpu() - ref_mu).abs().max().item() < TOL_FEATS
    assert (logvar[[0, 33, 39]].cpu() - ref_lv).abs().max().item() < 2 * TOL_FEATS
    mu2, _ = engine.encode(feats.flip(0))
    assert torch.equal(mu2.flip(0), mu)


def test_full_size_edit_permutation_equivariance(engine):
    """BASELINE configs[4] (edit_gesture, B=256) through a size-independent property: clips are independent
    (SURVEY 8e), so permuting the (content, emotion, style, noise) rows -- what style_Xemo_transfer does with the
    emotion/style banks -- permutes the poses bit for bit, and any clip equals its own B=1 run."""
    B = 256
    g = torch.Generator().manual_seed(4)
    l0, con, emo, sty = (torch.randn(B, d, generator=g) for d in (128, 256, 256, 256))
    perm = torch.randperm(B, generator=g)
    a = engine.diffusion_backward(l0, con, emo, sty, n_steps=50, sampler="ddim", want_latents=True)
    b = engine.diffusion_backward(l0[perm], con[perm], emo[perm], sty[perm], n_steps=50, sampler="ddim", want_latents=True)
    assert torch.equal(a["latents"][perm], b["latents"]) and torch.equal(a["poses"][perm], b["poses"])
    assert torch.isfinite(a["poses"]).all() and torch.equal(a["trans"][perm], b["trans"])
    # swapping only the emotion/style rows changes exactly the clips whose rows changed
    emo2 = emo.clone()
    emo2[:128] = emo[perm[:128]]
    c = engine.diffusion_backward(l0, con, emo2, sty, n_steps=50, sampler="ddim", want_latents=True)
    same_row = (emo2 == emo).all(dim=1)
    assert torch.equal(c["latents"][same_row], a["latents"][same_row])
    assert not torch.equal(c["latents"][~same_row], a["latents"][~same_row])
    one = engine.diffusion_backward(l0[200:201], con[200:201], emo[200:201], sty[200:201], n_steps=50, sampler="ddim",
                                    want_latents=True)
    assert torch.allclose(one["latents"][0], a["latents"][200], atol=1e-5)


def test_full_size_ddpm1000_b64_properties(engine):
    """BASELINE configs[2] at full size (the bench workload): 64 clips x 1000 ancestral steps with in-kernel
    Philox noise -- reproducible for a seed, seed-dependent, finite, and each half of the batch (what a
    2-GPU shard would own) reproduces the full run's latents when given its global element offset."""
    B = 64
    g = torch.Generator().manual_seed(9)
    l0, con, emo, sty = (torch.randn(B, d, generator=g) for d in (128, 256, 256, 256))
    a = engine.denoise(l0, con, emo, sty, n_steps=1000, sampler="ddpm", seed=11)
    b = engine.denoise(l0, con, emo, sty, n_steps=1000, sampler="ddpm", seed=11)
    c = engine.denoise(l0, con, emo, sty, n_steps=1000, sampler="ddpm", seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.isfinite(a).all()
    # GPU-count invariance: the two halves of the batch, run as separate calls with their
    # global clip offsets (what ranks 0 and 1 of a 2-GPU run do), draw the same Philox noise as the full run.
    # Without the offset the second half would reuse the first half's noise stream and differ grossly.
    lo = engine.denoise(l0[:32], con[:32], emo[:32], sty[:32], n_steps=1000, sampler="ddpm", seed=11, clip_offset=0)
    hi = engine.denoise(l0[32:], con[32:], emo[32:], sty[32:], n_steps=1000, sampler="ddpm", seed=11, clip_offset=32)
    halves = torch.cat([lo, hi])
    if os.environ.get("AMUSE_DENOISE_FFMA", "0") == "1":
        # the FFMA fallback kernel packs 1 or 2 clips per cluster with different K-split orders: same noise, fp32 reordering
        assert (halves - a).abs().max().item() < 5e-3
    else:
        assert torch.equal(halves, a)     # one clip per chain whatever the batch: bit-identical
    wrong = engine.denoise(l0[32:], con[32:], emo[32:], sty[32:], n_steps=1000, sampler="ddpm", seed=11, clip_offset=0)
    assert (wrong - a[32:]).abs().max().item() > 1.0
    poses, trans = engine.decode(a)
    assert torch.isfinite(poses).all() and poses.shape == (B, 300, 55, 3)
    assert poses.abs().max().item() <= 3.1416 + 1e-3          # axis-angle magnitude is an angle in [0, pi]


def test_philox_path_against_oracle_full_size(engine, synthetic_weights):
    """The benchmarked configuration itself -- B = 64, 1000 ancestral steps, noise drawn IN the kernel -- against the CPU
    oracle: the kernel's Philox normals are exported through the debug ABI (same device function, philox.cuh), checked
    for N(0,1) moments and stream independence, and handed to the oracle as ``step_noise``.  Tolerance as for the
    injected-noise golden (ddpm1000_b2): within max(2e-5 |z|max, 4 x the oracle's own fp32-vs-fp64 distance) of fp64."""
    B, n, seed = 64, 1000, 2024
    g = torch.Generator().manual_seed(31)
    l0, con, emo, sty = (torch.randn(B, d, generator=g) for d in (128, 256, 256, 256))
    noise = engine.philox_normals(seed, B, n).cpu()
    x = noise.double().flatten()
    N = x.numel()
    mean, var = x.mean().item(), x.var().item()
    kurt = ((x - mean) ** 4).mean().item() / var ** 2
    tail = (x.abs() > 3).double().mean().item()
    print(f"[philox] N={N} mean={mean:.2e} var={var:.5f} kurtosis={kurt:.4f} P(|z|>3)={tail:.5f} max|z|={x.abs().max().item():.2f}")
    assert abs(mean) < 4 / N ** 0.5 and abs(var - 1) < 4 * (2 / N) ** 0.5 and abs(kurt - 3) < 4 * (24 / N) ** 0.5
    assert abs(tail - 0.0026998) < 2e-4 and x.abs().max().item() < 7.0
    c01 = torch.corrcoef(torch.stack([noise[:, 0].flatten(), noise[:, 1].flatten()]))[0, 1].item()
    lag = torch.corrcoef(torch.stack([noise[:-1, 0].flatten(), noise[1:, 0].flatten()]))[0, 1].item()
    assert abs(c01) < 0.02 and abs(lag) < 0.02             # clips and consecutive steps draw independent streams
    # the same draws through the offset argument: clips 32.. of this stream are what a second shard would see
    assert torch.equal(engine.philox_normals(seed, 32, 4, clip_offset=32).cpu(), noise[:4, 32:])
    got = engine.denoise(l0, con, emo, sty, n_steps=n, sampler="ddpm", seed=seed)
    same = engine.denoise(l0, con, emo, sty, n_steps=n, sampler="ddpm", step_noise=noise)
    assert torch.equal(got, same)                          # in-kernel draws == the exported array fed back
    sd = synthetic_weights["denoiser"]
    z32 = R.sample_latents(sd, l0, con, emo, sty, n, "ddpm", noise)
    sd64 = {k: v.double() for k, v in sd.items()}
    z64 = R.sample_latents(sd64, l0.double(), con.double(), emo.double(), sty.double(), n, "ddpm", noise.double())
    scale = z64.abs().max().item()
    r = (z32.double() - z64).abs().max().item()
    e64 = (got.cpu().double() - z64).abs().max().item()
    print(f"[parity] philox ddpm1000 B=64: CUDA vs fp64 {e64:.3e}  oracle fp32 vs fp64 {r:.3e}  |z|max {scale:.1f}")
    assert e64 < max(2e-5 * scale, 4 * r)


def test_empty_batch(engine):
    """bsz = 0: the reference's modules return empty tensors for an empty batch; nothing is launched here."""
    z = torch.empty(0, 256)
    n0 = engine.launch_count()
    out = engine.diffusion_backward(torch.empty(0, 128), z, z, z, n_steps=50, want_latents=True, want_feats=True)
    assert tuple(out["poses"].shape) == (0, 300, 55, 3) and tuple(out["trans"].shape) == (0, 300, 3)
    assert tuple(out["latents"].shape) == (0, 128) and tuple(out["feats"].shape) == (0, 300, 333)
    assert out["poses"].device.type == "cuda" and out["poses"].dtype == torch.float32
    assert tuple(engine.denoise(torch.empty(0, 128), z, None, None).shape) == (0, 128)
    poses, trans = engine.decode(torch.empty(0, 128))
    assert tuple(poses.shape) == (0, 300, 55, 3) and tuple(trans.shape) == (0, 300, 3)
    mu, logvar = engine.encode(torch.empty(0, 300, 333))
    assert tuple(mu.shape) == (0, 128) and tuple(logvar.shape) == (0, 128)
    host = engine.diffusion_backward_host(torch.empty(0, 128), z, z, z, n_steps=50)
    assert tuple(host["poses"].shape) == (0, 300, 55, 3) and host["poses"].device.type == "cpu"
    assert engine.launch_count() == n0


def test_ffma_fallback_kernels_match(engine, synthetic_weights, monkeypatch):
    """The round-1 fp32 FFMA2 loop kernel (AMUSE_DENOISE_FFMA=1, the A/B switch next to the tcgen05 kernel) in its default
    and WIDE layouts (AMUSE_WIDE_ROWS=1: 10-row GEMM warps over 1/8 of K, partials parked in the idle exchange buffer),
    and the tcgen05 kernel, compute the same sampler: all three against the oracle, on batch sizes with full and
    half-filled clusters, a 4-token ablation, and an ancestral run on the shared Philox stream."""
    from amuse_b200.engine import Engine
    engines = {}
    for name, env in (("ffma", {"AMUSE_DENOISE_FFMA": "1"}), ("ffma-wide", {"AMUSE_DENOISE_FFMA": "1", "AMUSE_WIDE_ROWS": "1"})):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        engines[name] = Engine("cuda:0")
        for k_ in env:
            monkeypatch.delenv(k_)
        engines[name].load_state_dict("denoiser", synthetic_weights["denoiser"])
        engines[name].finalize()
    try:
        for B, sty in ((64, True), (35, False)):
            g = torch.Generator().manual_seed(77 + B)
            l0, con, emo = torch.randn(B, 128, generator=g), torch.randn(B, 256, generator=g), torch.randn(B, 256, generator=g)
            zs = torch.randn(B, 256, generator=g) if sty else None
            ref = R.sample_latents(synthetic_weights["denoiser"], l0, con, emo, zs, 10, "ddim")
            errs = {"tcgen05": (engine.denoise(l0, con, emo, zs, n_steps=10, sampler="ddim").cpu() - ref).abs().max().item()}
            for name, e in engines.items():
                errs[name] = (e.denoise(l0, con, emo, zs, n_steps=10, sampler="ddim").cpu() - ref).abs().max().item()
            print(f"[parity] loop kernels B={B}: " + "  ".join(f"{k_} max|d|={v_:.3e}" for k_, v_ in errs.items()))
            assert all(v_ < 1e-4 for v_ in errs.values())
        # the same stateless Philox draws in both kernel families: an ancestral run agrees to fp32 reordering
        l0, con = torch.randn(5, 128, generator=g), torch.randn(5, 256, generator=g)
        a = engine.denoise(l0, con, None, None, n_steps=20, sampler="ddpm", seed=9).cpu()
        b = engines["ffma"].denoise(l0, con, None, None, n_steps=20, sampler="ddpm", seed=9).cpu()
        assert (a - b).abs().max().item() < 2e-4 * max(1.0, a.abs().max().item())
    finally:
        for e in engines.values():
            e.close()


def test_attention_kernels_match(engine, synthetic_weights, monkeypatch):
    """MotionPrior's self-attention: the tcgen05 kernel (attn_tc.cu, the default) and the fp32 CUDA-core kernel
    (AMUSE_ATTN_FFMA=1) against the oracle's decode and encode -- 300 decoder frames (3 query tiles, the last with 44
    rows; keys padded 300 -> 304) and 302 encoder tokens."""
    from amuse_b200.engine import Engine
    from oracle.make_golden import synthetic_motion
    monkeypatch.setenv("AMUSE_ATTN_FFMA", "1")
    ffma = Engine("cuda:0")
    monkeypatch.delenv("AMUSE_ATTN_FFMA")
    ffma.load_state_dict("vae", synthetic_weights["vae"])
    ffma.finalize()
    try:
        z = torch.randn(3, 128, generator=torch.Generator().manual_seed(11))
        ref = R.vae_decode(synthetic_weights["vae"], z)
        errs = {}
        for name, e in (("tcgen05", engine), ("ffma", ffma)):
            _, _, feats = e.decode(z, want_feats=True)
            errs[name] = (feats.cpu() - ref).abs().max().item()
        print("[parity] decode, attention kernels: " + "  ".join(f"{k_} max|d|={v_:.3e}" for k_, v_ in errs.items()))
        assert all(v_ < TOL_FEATS for v_ in errs.values())
        poses, trans = synthetic_motion(2)
        feats = engine.motion_to_feats(poses, trans)
        mu_a, lv_a = engine.encode(feats)
        mu_b, lv_b = ffma.encode(feats)
        d = max((mu_a - mu_b).abs().max().item(), (lv_a - lv_b).abs().max().item())
        print(f"[parity] encode, attention kernels: max|d mu, logvar| = {d:.3e}")
        assert d < 2 * TOL_FEATS
    finally:
        ffma.close()


def test_profiling_instantiation_matches_product(engine):
    """`amuse_profile_arm` launches the second instantiation of the sampler loop (denoise_tc_kernel<true>, with the in-kernel
    clock stamps); the product launch compiles them out.  Same latents bit for bit, and the stamps are filled."""
    g = torch.Generator().manual_seed(21)
    l0, con, emo, sty = (torch.randn(6, d, generator=g) for d in (128, 256, 256, 256))
    ref = engine.denoise(l0, con, emo, sty, n_steps=6, sampler="ddpm", seed=5).cpu()
    engine.profile_arm(2)
    got = engine.denoise(l0, con, emo, sty, n_steps=6, sampler="ddpm", seed=5).cpu()
    st = engine.profile_read(512)
    assert torch.equal(got, ref)
    if os.environ.get("AMUSE_DENOISE_FFMA") != "1":
        assert st[92] > st[0] > 0 and 20_000 < st[92] - st[0] < 2_000_000     # one denoiser step, in cycles
        assert all(v > 0 for v in st[300:380])                                 # the issuer's per-stage stamps
    assert torch.equal(engine.denoise(l0, con, emo, sty, n_steps=6, sampler="ddpm", seed=5).cpu(), ref)   # disarmed again
