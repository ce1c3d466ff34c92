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


def test_ast_depth12_golden(golden_dir):
    """The shipped depth (12 blocks x 3 branches, what bench.py's scope E runs) against the committed float64 output of the
    CPU restatement for one clip (oracle/make_ast_golden.py).  Tolerance: the tensor cores accumulate with truncation,
    so the error grows with depth; measured 4-5e-5 on |feat| ~ 3, asserted at 3e-4 like the shallow stacks."""
    g = np.load(golden_dir / "ast_depth12_b1.npz")
    eng, sd = _engine_with_ast(12)
    assert W.checksum({k: sd[k] for k in list(sd)[:8]}) == str(g["weights_sha1"]), "synthetic AST weights differ from the fixture's"
    fb = torch.randn(1, 1024, 128, generator=torch.Generator().manual_seed(int(g["fbank_seed"]))) * 0.5
    assert torch.equal(fb[0, [0, 511, 1023], :4], torch.from_numpy(g["fbank_probe"])), "torch CPU RNG stream differs from the fixture's"
    con, emo, sty = eng.ast_features(fb)
    for name, got in (("con", con), ("emo", emo), ("sty", sty)):
        ref64, ref32 = g[f"{name}_f64"], g[f"{name}_f32"]
        e64 = np.abs(got.cpu().double().numpy() - ref64).max()
        r = np.abs(ref32.astype(np.float64) - ref64).max()
        print(f"[parity] ast depth=12 {name}: |cuda-f64|={e64:.3e} |f32ref-f64|={r:.3e} |feat|max={np.abs(ref64).max():.2f}")
        assert e64 < 3e-4
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


def test_process_loader_style_xemo_transfer():
    """Config 5 (edit_gesture, style_Xemo_transfer): the dataset-driven edit preparation of the mirror class
    (reference infer_ldm.py:225-323, 416-502) on a synthetic data_dict: dict contract, the cross-links, and the
    latents against the oracle (AST features of the reference's chunk slicing; VAE latent with the same draw)."""
    from amuse_b200.infer_ldm import PretrainedLPDM_v1
    from oracle import lpdm_ref as R
    from oracle.make_golden import synthetic_motion
    den, vae, ast = W.denoiser_state_dict(), W.motionprior_state_dict(), W.ast_state_dict(depth=1)
    m = PretrainedLPDM_v1.from_state_dicts(den, vae, ast, device="cuda:0")
    m.style_Xemo_transfer = True
    poses, trans = synthetic_motion(4)                                   # 4 recordings x 300 frames
    g = torch.Generator().manual_seed(5)

    def rec(i, label):
        motion = torch.cat((poses[i].reshape(300, 165), trans[i]), dim=1)
        motion = torch.cat((motion, motion.flip(0)), dim=0)               # 600 frames = 2 takes of 300
        return {"ld_motion": motion.numpy(), "ld_waveform": 0.1 * torch.randn(1, 320000 + 123, generator=g),
                "ld_emo_label": label}

    t_ang, t_hap = "0_73_73", "0_65_65"
    data = {"lu": {t_ang: rec(0, "angry"), t_hap: rec(1, "happy")},
            "lawrence": {t_ang: rec(2, "angry"), t_hap: rec(3, "happy")}}
    info = f"[lu-lawrence]_[angry-happy]_*lu_angry_{t_ang}*lu_happy_{t_hap}*lawrence_angry_{t_ang}*lawrence_happy_{t_hap}*"
    torch.manual_seed(123)
    out = m.process_loader({"style_Xemo_transfer_info": info, "style_Xemo_transfer": data})["style_Xemo_transfer"]
    assert out["takes"] == f"{t_ang}*{t_hap}*{t_ang}*{t_hap}"
    e = out["lu"][t_ang]
    assert e["ld_z"].shape == (2, 128) and e["ld_z_con"].shape == (2, 256) and e["ld_z"].is_cuda
    # cross-links (infer_ldm.py:308-318): lu/angry gets lawrence/happy's emotion + style, and so on
    assert e[f"ld_z_emo_lawrence_{t_hap}"] is out["lawrence"][t_hap]["ld_z_emo"]
    assert e[f"ld_z_sty_lawrence_{t_hap}"] is out["lawrence"][t_hap]["ld_z_sty"]
    assert out["lawrence"][t_hap][f"ld_z_emo_lu_{t_ang}"] is e["ld_z_emo"]
    assert out["lu"][t_hap][f"ld_z_sty_lawrence_{t_ang}"] is out["lawrence"][t_ang]["ld_z_sty"]
    # AST features: chunk k of the reference is audio[:, k:k+160000]
    wav = data["lu"][t_ang]["ld_waveform"]
    fb = torch.stack([A.fbank_features(wav[:, k:k + 160000]) for k in range(2)])
    rc, re, rs = A.ast_features(ast, fb)
    assert (e["ld_z_con"].cpu() - rc).abs().max().item() < 3e-4
    assert (e["ld_z_sty"].cpu() - rs).abs().max().item() < 3e-4
    # VAE latent: first recording encoded first, so its draw is the first one after the seed
    motion = torch.from_numpy(data["lu"][t_ang]["ld_motion"]).view(2, 300, 168)
    mu, lv = R.vae_encode(vae, R.motion_to_feats(motion[:, :, :165].reshape(2, 300, 55, 3), motion[:, :, 165:]))
    torch.manual_seed(123)
    eps = torch.empty((1, 2, 128), device="cuda:0").normal_().cpu()[0]
    assert (e["ld_z"].cpu() - (mu + eps * lv.exp().pow(0.5))).abs().max().item() < 2e-4
    # the swapped conditions drive the sampler like any others (trainer.py:552-605)
    res = m.diffusion_backward(2, e["ld_z_con"], e[f"ld_z_emo_lawrence_{t_hap}"], e[f"ld_z_sty_lawrence_{t_hap}"])
    assert res["poses"].shape == (2, 300, 55, 3) and torch.isfinite(res["poses"]).all()
    m.engine.close()


def test_ast_batch_invariance_two_passes():
    """33 clips = two passes (17 + 16) of the AST stack: every clip's features equal its own single-clip run
    (per-row GEMM / per-(clip, head) attention: no cross-clip arithmetic), and the zero pad rows / columns of
    the q / k / v^T planes stay clean across passes."""
    eng, sd = _engine_with_ast(1)
    fb = torch.randn(33, 1024, 128, generator=torch.Generator().manual_seed(3)) * 0.5
    fb[7] *= 40.0                                              # a loud clip must not leak into its neighbours
    con, emo, sty = eng.ast_features(fb)
    for i in (0, 6, 8, 16, 17, 32):
        c1, e1, s1 = eng.ast_features(fb[i:i + 1])
        assert torch.equal(c1[0], con[i]) and torch.equal(e1[0], emo[i]) and torch.equal(s1[0], sty[i])
    assert torch.isfinite(con).all() and torch.isfinite(sty).all()
    eng.close()
