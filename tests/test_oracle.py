"""CPU tests of the oracle itself: (1) the restatement against the committed golden vectors
(produced by the reference's own modules), everywhere; (2) against the reference modules imported
live, in the build container only (/root/reference does not exist on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import lpdm_ref as R
from oracle import reference_loader as L
from oracle import weights as W


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def sds(synthetic_weights):
    d, v = synthetic_weights["denoiser"], synthetic_weights["vae"]
    return d, v, R.cast_sd(d, torch.float64), R.cast_sd(v, torch.float64)


def test_golden_checksums(golden_dir, synthetic_weights):
    g = np.load(golden_dir / "denoiser_step.npz")
    assert W.checksum(synthetic_weights["denoiser"]) == str(g["denoiser_sha1"])
    assert W.checksum(synthetic_weights["vae"]) == str(g["motionprior_sha1"])


def test_denoiser_step_golden(golden_dir, sds):
    d32, _, d64, _ = sds
    g = np.load(golden_dir / "denoiser_step.npz")
    x, con, emo, sty = (_t(g[k]) for k in ("x", "con", "emo", "sty"))
    for t in (981, 500, 1):
        assert (R.denoiser_forward(d32, x, t, con, emo, sty) - _t(g[f"eps_t{t}_f32"])).abs().max() < 1e-5
        e64 = R.denoiser_forward(d64, x.double(), t, con.double(), emo.double(), sty.double())
        assert (e64 - _t(g[f"eps_t{t}_f64"])).abs().max() < 1e-12
    e = R.denoiser_forward(d64, x.double(), 981, con.double(), None, None)
    assert (e - _t(g["eps_t981_noemo_nosty_f64"])).abs().max() < 1e-12
    e = R.denoiser_forward(d64, x.double(), 981, con.double(), emo.double(), None)
    assert (e - _t(g["eps_t981_nosty_f64"])).abs().max() < 1e-12


@pytest.mark.parametrize("name", ["ddim50_b4", "ddim1_b1", "ddim50_b1"])
def test_ddim_golden(golden_dir, sds, name):
    d32, _, d64, _ = sds
    g = np.load(golden_dir / f"{name}.npz")
    n = int(g["n_steps"])
    a = [_t(g[k]) for k in ("latents0", "con", "emo", "sty")]
    z64 = R.sample_latents(d64, *[t.double() for t in a], n, "ddim")
    assert (z64 - _t(g["z_f64"])).abs().max() < 1e-10
    z32 = R.sample_latents(d32, *a, n, "ddim")
    assert (z32 - _t(g["z_f32"])).abs().max() < 2e-4


def test_ddpm_golden(golden_dir, sds):
    _, _, d64, _ = sds
    g = np.load(golden_dir / "ddpm100_b2.npz")
    n, B = int(g["n_steps"]), 2
    noise = torch.randn(n, B, 128, generator=torch.Generator().manual_seed(int(g["noise_seed"])))
    a = [_t(g[k]).double() for k in ("latents0", "con", "emo", "sty")]
    z64 = R.sample_latents(d64, *a, n, "ddpm", noise.double())
    assert (z64 - _t(g["z_f64"])).abs().max() < 1e-9


def test_decode_and_rot_golden(golden_dir, sds):
    _, v32, _, v64 = sds
    g = np.load(golden_dir / "decode_b2.npz")
    idx = g["frame_idx"]
    f64 = R.vae_decode(v64, _t(g["z"]).double())
    assert (f64[:, idx] - _t(g["feats_f64"])).abs().max() < 1e-11
    f32 = R.vae_decode(v32, _t(g["z"]))
    assert (f32[:, idx] - _t(g["feats_f32"])).abs().max() < 2e-5
    m = R.feats_to_motion(f64)
    assert R.geodesic_deg(m["poses"][:, idx], _t(g["poses_f64"])).max() < 1e-4   # acos near 1: ~1e-6 deg is rounding noise
    c = np.load(golden_dir / "rot6d_cases.npz")
    aa = R.rot6d_to_axis_angle(_t(c["d6"]).double())
    assert R.geodesic_deg(aa, _t(c["aa_f64"])).max() < 1e-4
    assert torch.allclose(R.rot6d_to_axis_angle(_t(c["d6"])), _t(c["aa_f32"]), atol=1e-5)


def test_scheduler_tables():
    assert R.ddim_timesteps(50)[:3] == [981, 961, 941] and R.ddim_timesteps(50)[-1] == 1
    assert R.ddim_timesteps(1) == [1]
    with pytest.raises(IndexError):
        R.ddim_timesteps(1000)                       # alphas_cumprod[1000] -- invalid in the reference too
    assert R.ddpm_timesteps(1000)[0] == 999 and R.ddpm_timesteps(1000)[-1] == 0
    c = R.ddpm_coeffs(1000)["coef"]
    assert c[-1, 4] == 0 and (c[:-1, 4] > 0).all()   # no noise at t = 0
    ac = R.alphas_cumprod()
    assert abs(float(ac[0]) - (1 - 0.00085)) < 1e-6 and float(ac[-1]) < 0.01
    # DDIM eta=0 is deterministic: direction coefficient is sqrt(1 - a')
    d = R.ddim_coeffs(50)["coef"]
    assert torch.allclose(d[:, 2] ** 2 + d[:, 3] ** 2, torch.ones(50), atol=1e-6)


@pytest.mark.skipif(not L.available(), reason="/root/reference not present (GPU box)")
def test_restatement_matches_reference_modules(sds):
    d32, v32, _, _ = sds
    den, vae, tf = L.load_denoiser(d32), L.load_motionprior(v32), L.load_transforms()
    g = torch.Generator().manual_seed(5)
    B = 3
    x, con, emo, sty = (torch.randn(B, n, generator=g) for n in (128, 256, 256, 256))
    with torch.no_grad():
        for t, e, s in ((981, emo, sty), (1, emo, None), (500, None, None)):
            u = lambda z: None if z is None else z[:, None, :]
            ref = den(sample=x[:, None, :], timestep=torch.tensor(t), con_hidden=u(con), emo_hidden=u(e),
                      sty_hidden=u(s), lengths=[300] * B)[0][:, 0, :]
            assert (R.denoiser_forward(d32, x, t, con, e, s) - ref).abs().max() < 1e-5
        z = torch.randn(2, 128, generator=g)
        ref = vae.decode(z[None], [300, 300])
        assert (R.vae_decode(v32, z) - ref).abs().max() < 2e-5
        d6 = ref[:, :, :-3].reshape(2, 300, 55, 6)
        assert torch.equal(R.rot6d_to_axis_angle(d6), tf.matrix_to_axis_angle(tf.rotation_6d_to_matrix(d6)))


def test_encode_golden(golden_dir, sds):
    """MotionPrior.encode restatement + axis-angle -> 6D against the reference-generated fixture."""
    from oracle.make_golden import synthetic_motion
    g = np.load(golden_dir / "encode_b2.npz")
    poses, trans = synthetic_motion(2)
    feats = R.motion_to_feats(poses, trans)
    assert np.abs(feats[:, g["frame_idx"]].numpy() - g["feats_f32"]).max() < 1e-6
    _, v32, _, v64 = sds
    mu, lv = R.vae_encode(v32, feats)
    assert np.abs(mu.numpy() - g["mu_f64"]).max() < 2e-5
    assert np.abs(lv.exp().pow(0.5).numpy() - g["std_f64"]).max() < 2e-5
    mu64, lv64 = R.vae_encode(v64, R.motion_to_feats(poses.double(), trans.double()))
    assert np.abs(mu64.numpy() - g["mu_f64"]).max() < 1e-10
    assert np.abs(lv64.exp().pow(0.5).numpy() - g["std_f64"]).max() < 1e-10


def _timm_to_hf(sd, prefix):
    """timm-0.4.5 DeiT key names (oracle/weights.py) -> HuggingFace ``ASTModel`` key names."""
    v = f"{prefix}.v"
    out = {"embeddings.cls_token": sd[f"{v}.cls_token"], "embeddings.distillation_token": sd[f"{v}.dist_token"],
           "embeddings.position_embeddings": sd[f"{v}.pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd[f"{v}.patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd[f"{v}.patch_embed.proj.bias"],
           "layernorm.weight": sd[f"{v}.norm.weight"], "layernorm.bias": sd[f"{v}.norm.bias"]}
    i = 0
    while f"{v}.blocks.{i}.norm1.weight" in sd:
        b, h = f"{v}.blocks.{i}", f"encoder.layer.{i}"
        D = sd[f"{b}.attn.proj.weight"].shape[0]
        for j, nm in enumerate(("query", "key", "value")):        # timm packs q|k|v along the output axis
            out[f"{h}.attention.attention.{nm}.weight"] = sd[f"{b}.attn.qkv.weight"][j * D:(j + 1) * D]
            out[f"{h}.attention.attention.{nm}.bias"] = sd[f"{b}.attn.qkv.bias"][j * D:(j + 1) * D]
        for wb in ("weight", "bias"):
            out[f"{h}.attention.output.dense.{wb}"] = sd[f"{b}.attn.proj.{wb}"]
            out[f"{h}.layernorm_before.{wb}"] = sd[f"{b}.norm1.{wb}"]
            out[f"{h}.layernorm_after.{wb}"] = sd[f"{b}.norm2.{wb}"]
            out[f"{h}.intermediate.dense.{wb}"] = sd[f"{b}.mlp.fc1.{wb}"]
            out[f"{h}.output.dense.{wb}"] = sd[f"{b}.mlp.fc2.{wb}"]
        i += 1
    return out, i


def test_ast_restatement_matches_hf_transformers():
    """Independent pin of the AST restatement (oracle/ast_ref.py).  timm 0.4.5 (what the reference
    asserts, audio_main_new.py:52) is not installable, but the image's ``transformers`` ships its own
    implementation of the same published AST/DeiT encoder (``ASTModel``: 16x16 patches, stride 10,
    cls + distillation tokens, pre-LN blocks).  Same weights through both must agree; the reference's
    frame-based head (audio_main_new.py:196-204: mean of tokens 2:, LayerNorm, Linear) is applied to
    HF's ``last_hidden_state`` here."""
    tr = pytest.importorskip("transformers")
    from oracle import ast_ref as A
    sd = W.ast_state_dict(depth=2)
    g = torch.Generator().manual_seed(7)
    fb = torch.randn(2, 1024, 128, generator=g)
    for prefix in ("con_enc", "emo_enc"):
        hf_sd, depth = _timm_to_hf(sd, prefix)
        cfg = tr.ASTConfig(num_hidden_layers=depth, layer_norm_eps=1e-6, hidden_dropout_prob=0.0,
                           attention_probs_dropout_prob=0.0)
        model = tr.ASTModel(cfg).eval()
        missing = model.load_state_dict(hf_sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        with torch.no_grad():
            tokens = model(input_values=fb).last_hidden_state                       # after the final LayerNorm
            feat = tokens[:, 2:].mean(dim=1)
            feat = torch.nn.functional.layer_norm(feat, (768,), sd[f"{prefix}.feature_head.0.weight"],
                                                  sd[f"{prefix}.feature_head.0.bias"], 1e-5)
            want = torch.nn.functional.linear(feat, sd[f"{prefix}.feature_head.1.weight"],
                                              sd[f"{prefix}.feature_head.1.bias"])
            got = A.ast_branch(sd, prefix, fb)
        assert got.shape == (2, 256)
        assert float((got - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))


def test_scheduler_tables_against_published_formulas():
    """The restated scheduler tables against the closed forms of the papers, evaluated independently in float64
    numpy: scaled-linear betas (latent-diffusion convention), DDIM eq. 12 with eta = 0 (Song et al.),
    the DDPM posterior mean / fixed-small variance (Ho et al. eq. 6-7).  diffusers itself is not installable here
    (parity unpinned); this pins the restatement to the published algorithm instead."""
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas)
    assert np.allclose(R.alphas_cumprod().double().numpy(), ac, rtol=2e-6, atol=0)
    # DDIM, 50 steps, leading-offset timesteps t = 1 + 20 i, final alpha = alphas_cumprod[0]
    plan = R.ddim_coeffs(50)
    ts = np.array(plan["timesteps"])
    assert np.array_equal(ts, (np.arange(50) * 20)[::-1] + 1)
    a, ap = ac[ts], np.where(ts - 20 >= 0, ac[np.maximum(ts - 20, 0)], ac[0])
    want = np.stack([a ** 0.5, (1 - a) ** 0.5, ap ** 0.5, (1 - ap) ** 0.5, np.zeros(50)], 1)
    # fp32 library order: 1 - a is rounded in fp32 before the sqrt, a relative 4e-5 on sqrt(1 - a) where a -> 1
    assert np.allclose(plan["coef"].double().numpy(), want, rtol=3e-6, atol=3e-6)
    # DDPM, 1000 steps: x' = c0 x0 + cx x + sigma z with the posterior coefficients of Ho et al.
    plan = R.ddpm_coeffs(1000)
    ts = np.array(plan["timesteps"])
    assert np.array_equal(ts, np.arange(1000)[::-1])
    a, ap = ac[ts], np.where(ts > 0, ac[np.maximum(ts - 1, 0)], 1.0)
    beta_t = 1.0 - a / ap
    c0 = ap ** 0.5 * beta_t / (1 - a)
    cx = (a / ap) ** 0.5 * (1 - ap) / (1 - a)
    sigma = np.where(ts > 0, np.sqrt(np.maximum((1 - ap) / (1 - a) * beta_t, 1e-20)), 0.0)
    want = np.stack([a ** 0.5, (1 - a) ** 0.5, c0, cx, sigma], 1)
    got = plan["coef"].double().numpy()
    assert np.allclose(got[:, :2], want[:, :2], rtol=3e-6, atol=3e-6)
    assert np.allclose(got[:, 2:], want[:, 2:], rtol=3e-4, atol=3e-6)   # beta_t = 1 - a/a' and 1 - a cancel in fp32
    # one full step of each sampler against the formula applied directly
    g = torch.Generator().manual_seed(3)
    x, e, z = (torch.randn(4, 128, generator=g, dtype=torch.float64) for _ in range(3))
    i = 500
    got = R.scheduler_step(plan, i, x, e, z).numpy()
    x0 = (x.numpy() - (1 - a[i]) ** 0.5 * e.numpy()) / a[i] ** 0.5
    assert np.allclose(got, c0[i] * x0 + cx[i] * x.numpy() + sigma[i] * z.numpy(), rtol=1e-5, atol=1e-6)
