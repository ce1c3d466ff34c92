"""PretrainedLPDM_v1.setup() end to end (a14; reference infer_ldm.py:30-128, infer_pretrained_vae.py:13-49,
infer_pretrained_ast_evp.py:12-41): a temporary tree laid out like the reference's --

    <root>/configs/diff_latent_v2.json
    <root>/saved-models/<lpdm run>/latdiff_*_total<loss>_e<epoch>.pt, prior_*_total<loss>_e<epoch>.pt
    <root>/saved-models/<ast run>/wav_<epoch>_x_tEA<acc>_tPA<acc>.pkl
    <root>/data/BEAT-processed            (`processed`; the configs are found through processed.parents[1])

-- with checkpoints written by torch.save in the reference's formats ({"epoch", "model_state_dict"} for the LDM and the
prior, a bare state-dict for AST_EVP), then config read -> file choice -> torch.load -> engine -> diffusion_backward,
checked against the oracle.  The weights are the seeded synthetic state-dicts (the reference ships none)."""
import json
from pathlib import Path

import pytest
import torch

from oracle import lpdm_ref as R
from oracle import weights as W

pytestmark = pytest.mark.gpu

# the keys setup() reads, with the values of the released configs/diff_latent_v2.json
LDM_CFG = {
    "arch_denoiser": {"nfeats": 201, "latent_dim": [1, 128], "ff_size": 512, "num_layers": 9, "num_heads": 4, "dropout": 0.1,
                      "guidance_scale": 7.5, "guidance_uncondp": 0.1, "arch": "trans_enc", "normalize_before": False,
                      "activation": "gelu", "position_embedding": "learned", "cond_dim": 256, "nclasses": 7, "freq_shift": 0,
                      "ablation_skip_connection": True, "pe_type": "mld", "flip_sin_to_cos": True,
                      "return_intermediate_dec": False, "diffusion_only": False},
    "noisy_scheduler": {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
                        "variance_type": "fixed_small", "clip_sample": False, "prediction_type": "epsilon"},
    "scheduler": {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
                  "set_alpha_to_one": False, "steps_offset": 1, "num_inference_timesteps": 50, "eta": 0.0},
}


def _config(lpdm_run, ast_run):
    off = {"use": False}
    return {"TRAIN_PARAM": {
        "tag": "latent_diffusion",
        "latent_diffusion": {"smplx_data": True, "smplx_rep": "6D", "skip_trans": False, "train_upper_body": False,
                             "arch": "diff_latent_v2", "pretrained_lpdm": lpdm_run, "pretrained_ast": ast_run,
                             "pretrained_prior_lpdm_e": "best", "pretrained_ldm_lpdm_e": "best"},
        "test": {"style_transfer": off, "emotion_control": off, "content_control": off, "style_Xemo_transfer": off},
        "wav_dtw_mfcc": {"ablation": "full", "frame_based_feats": True, "target_length": 1024, "num_mel_bins": 128,
                         "dataset_mean": -9.173025, "dataset_std": 5.062332}},
        "DATA_PARAM": {"Bvh": {"train_pose_framelen": 300}}}


def _tree(root: Path, den, vae, ast):
    (root / "configs").mkdir(parents=True)
    (root / "configs" / "diff_latent_v2.json").write_text(json.dumps(LDM_CFG))
    run = root / "saved-models" / "LPDM_run"
    run.mkdir(parents=True)
    ldm_sd = {f"denoiser.{k}": v for k, v in den.items()}
    decoy = {k: torch.zeros_like(v) for k, v in ldm_sd.items()}           # a worse checkpoint that must NOT be chosen
    torch.save({"epoch": 6000, "model_state_dict": ldm_sd}, run / "latdiff_LPDM_total0.4_e6000.pt")
    torch.save({"epoch": 5800, "model_state_dict": decoy}, run / "latdiff_LPDM_total0.7_e5800.pt")
    torch.save({"epoch": 6000, "model_state_dict": vae}, run / "prior_LPDM_total1.3_e6000.pt")
    torch.save({"epoch": 5800, "model_state_dict": {k: torch.zeros_like(v) for k, v in vae.items()}},
               run / "prior_LPDM_total0.9_e5800.pt")                     # lower loss, wrong epoch: prior follows the LDM's epoch
    (run / "experiment_args.json").write_text("{}")
    adir = root / "saved-models" / "AST_run"
    adir.mkdir(parents=True)
    torch.save(ast, adir / "wav_12_x_tEA0.91_tPA0.99.pkl")
    torch.save({k: torch.zeros_like(v) for k, v in ast.items()}, adir / "wav_3_x_tEA0.55_tPA0.60.pkl")
    processed = root / "data" / "BEAT-processed"
    processed.mkdir(parents=True)
    return processed


def test_setup_end_to_end(tmp_path):
    from amuse_b200.infer_ldm import PretrainedLPDM_v1
    den, vae, ast = W.denoiser_state_dict(), W.motionprior_state_dict(), W.ast_state_dict(depth=1)
    processed = _tree(tmp_path, den, vae, ast)
    m = PretrainedLPDM_v1(base_prior=None)
    epoch = m.setup(_config("LPDM_run", "AST_run"), "cuda:0", processed, None, False)
    assert epoch == 6000                                                   # the lowest-loss LDM checkpoint's epoch
    assert m.num_inference_timesteps == 50 and m.eta == 0.0 and m.has_ast
    B = 3
    g = torch.Generator().manual_seed(11)
    con, emo, sty = (torch.randn(B, 256, generator=g).cuda() for _ in range(3))
    torch.manual_seed(123)
    out = m.diffusion_backward(B, con, emo, sty)
    torch.manual_seed(123)
    lat = torch.randn((B, 1, 128), device="cuda", dtype=torch.float)      # the draw diffusion_backward makes (infer_ldm.py:137)
    ref = R.diffusion_backward(den, vae, lat.view(B, 128).cpu(), con.cpu(), emo.cpu(), sty.cpu(), n_steps=50, sampler="ddim")
    assert out["poses"].shape == (B, 300, 55, 3) and out["trans"].shape == (B, 300, 3)
    # conditioning-aware pose check (oracle/lpdm_ref.py::check_poses): the 6D feature error is the trans columns' (the
    # API returns no feats), doubled for margin
    e_feat = 2 * (out["trans"].cpu() - ref["trans"]).abs().max().item() + 2e-5
    ok, geo, well = R.check_poses(out["poses"].cpu(), ref["poses"], ref["feats"], e_feat)
    print(f"[setup] pose geodesic vs oracle: max {geo:.4f} deg, well-conditioned max {well:.4f} deg")
    assert ok
    assert (out["trans"].cpu() - ref["trans"]).abs().max().item() < 2e-4
    # the audio side loaded by setup(): features of one synthetic chunk against the restatement
    from oracle import ast_ref as A
    wav = 0.1 * torch.randn(1, 160000, generator=torch.Generator().manual_seed(0))
    c1, e1, s1 = m.process_single_seq(wav - wav.mean(), framerate=16000)
    rc, re, rs = A.ast_features(ast, A.fbank_features(wav - wav.mean())[None])
    assert (c1.cpu() - rc).abs().max().item() < 3e-4 and (s1.cpu() - rs).abs().max().item() < 3e-4
    m.engine.close()


def test_setup_error_behaviour(tmp_path):
    from amuse_b200.infer_ldm import PretrainedLPDM_v1, mapinfo2takes
    den, vae, ast = W.denoiser_state_dict(), W.motionprior_state_dict(), W.ast_state_dict(depth=1)
    processed = _tree(tmp_path, den, vae, ast)
    cfg = _config("LPDM_run", "missing_run")
    with pytest.raises(FileNotFoundError):                                 # the reference iterates the AST directory and raises
        PretrainedLPDM_v1(None).setup(cfg, "cuda:0", processed, None, False)
    cfg = _config("LPDM_run", "AST_run")
    cfg["TRAIN_PARAM"]["wav_dtw_mfcc"]["frame_based_feats"] = False
    with pytest.raises(NotImplementedError):
        PretrainedLPDM_v1(None).setup(cfg, "cuda:0", processed, None, False)
    cfg = _config("LPDM_run", "AST_run")
    cfg["TRAIN_PARAM"]["latent_diffusion"]["pretrained_prior_lpdm_e"] = 100
    with pytest.raises(AssertionError):                                    # infer_ldm.py:64
        PretrainedLPDM_v1(None).setup(cfg, "cuda:0", processed, None, False)
    with pytest.raises(Exception, match="Unknown emotion"):
        mapinfo2takes("[scott]_[bored]")
    assert mapinfo2takes("[scott]_[happy]") == mapinfo2takes("happy", trainer=True)


def test_smplx_3d_tail(tmp_path):
    """smplx_rep != "6D" (infer_ldm.py:175-177): the decoder's features are returned as [b, t, j, 3] without a rotation
    conversion."""
    from amuse_b200.infer_ldm import PretrainedLPDM_v1
    den, vae = W.denoiser_state_dict(), W.motionprior_state_dict()
    m = PretrainedLPDM_v1.from_state_dicts(den, vae, None, device="cuda:0")
    m.smplx_rep = "3D"
    B = 2
    con = torch.randn(B, 256, generator=torch.Generator().manual_seed(1)).cuda()
    torch.manual_seed(7)
    out = m.diffusion_backward(B, con, None, None)
    torch.manual_seed(7)
    lat = torch.randn((B, 1, 128), device="cuda", dtype=torch.float)
    ref = R.diffusion_backward(den, vae, lat.view(B, 128).cpu(), con.cpu(), None, None, n_steps=50, sampler="ddim")
    assert out["poses"].shape == (B, 300, 110, 3) and out["trans"].shape == (B, 300, 3)
    assert (out["poses"].reshape(B, 300, 330).cpu() - ref["feats"][:, :, :330]).abs().max().item() < 2e-4
    m.engine.close()
