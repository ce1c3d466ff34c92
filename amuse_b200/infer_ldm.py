"""Drop-in replacement for the reference module ``models/latent_diffusion/infer_ldm.py``.

``scripts/main.py:54`` imports ``PretrainedLPDM_v1`` from that module and ``scripts/trainer.py:39``
imports ``mapinfo2takes``; the trainer then treats the object as a duck-typed model
(``trainer.py:522-523, 552-553, 605, 746, 865, 1049-1066``).  This file keeps that surface --
same constructor, same method names, argument meaning, return types and error behaviour -- and
routes every tensor operation to the CUDA engine behind the C ABI (``include/amuse_b200.h``).
Nothing here imports the reference, diffusers, timm or pytorch3d, and there is no PyTorch
fallback for the math: without the built library / a B200 the constructor of the engine raises.

Shadowing recipe: put a ``models/latent_diffusion/infer_ldm.py`` stub that does
``from amuse_b200.infer_ldm import *`` ahead of the reference on ``sys.path``.
"""
from __future__ import annotations

import json
import re
from pathlib import Path
from typing import Dict, Optional

import torch

from .engine import Engine

__all__ = ["PretrainedLPDM_v1", "mapinfo2takes"]

# take ids of the BEAT emotion takes (reference dm/utils/ldm_evals.py:79-86) -- data constants
_TAKES = {
    "happy": ["0_65_65", "0_66_66"], "sad": ["0_81_81", "0_82_82"], "angry": ["0_73_73", "0_74_74"],
    "contempt": ["0_87_87", "0_88_88"], "disgust": ["0_111_111", "0_112_112"],
    "surprise": ["0_95_95", "0_96_96"], "fear": ["0_103_103", "0_104_104"],
}


def mapinfo2takes(info, trainer=False):
    """``[ayana-scott]_[fear]`` -> the two take ids of that emotion (reference infer_ldm.py:519-527).
    Returns None when no emotion name occurs, like the reference's fall-through."""
    if not trainer:
        info = info.split("_")[1]
    for emotion in ("happy", "sad", "angry", "contempt", "disgust", "surprise", "fear"):   # reference test order
        if emotion in info:
            return _TAKES[emotion]
    return None


def _first_int(s: str) -> int:
    return int(re.search(r"\d+", s).group())


def _pick_by_loss_or_epoch(files, which):
    """Checkpoint choice of the reference (infer_ldm.py:78-86, infer_pretrained_vae.py:36-42):
    "best" = smallest total loss encoded in the second-to-last ``_`` field of the file stem,
    otherwise the file whose last field carries that epoch number."""
    if which == "best":
        best, best_loss = None, float("inf")
        for f in files:
            loss = float(re.findall(r"\d+\.\d+", f.stem.split("_")[-2])[0])
            if loss < best_loss:
                best, best_loss = f, loss
        if best is None:
            raise FileNotFoundError("no checkpoint candidates")
        return best
    return [f for f in files if _first_int(f.stem.split("_")[-1]) == int(which)][0]


def _ast_num(x: str):
    chars = "".join(c if (c.isdigit() or c == ".") else " " for c in x).split()
    return float(chars[0]) if chars else None


class PretrainedLPDM_v1:
    """Inference facade: weights -> engine, audio -> (con, emo, sty), features -> SMPL-X poses."""

    # engine-side knobs that the reference hard-wires; defaults reproduce the reference
    sampler = "ddim"               # diffusers.DDIMScheduler (infer_ldm.py:116); "ddpm" = ancestral sampler
    device_fbank = True            # filterbank on the GPU; False = torchaudio on the host, as the reference

    def __init__(self, base_prior, base_con_ae=None, base_emo_ae=None, base_audio_ae=None):
        self.base_vae = base_prior
        self.con_ae = base_con_ae
        self.emo_ae = base_emo_ae
        self.base_ae = base_audio_ae
        self.engine: Optional[Engine] = None

    # ------------------------------------------------------------------ setup (infer_ldm.py:30-128)
    def _read_config(self, config, device, processed, backup_cfg, baseline, diffonly):
        self.config, self.device, self.processed = config, torch.device(device), processed
        self.backup_cfg, self.baseline, self.diffonly = backup_cfg, baseline, diffonly
        tp = config["TRAIN_PARAM"]
        ld = tp["latent_diffusion"]
        self.tag = tp["tag"]
        self.smplx_data, self.smplx_rep = ld["smplx_data"], ld["smplx_rep"]
        self.skip_trans, self.train_upper_body = ld["skip_trans"], ld["train_upper_body"]
        if self.train_upper_body:
            self.lower_body_jts = [1, 2, 4, 5, 7, 8, 10, 11]
        test = tp["test"]
        self.style_transfer = test["style_transfer"]["use"]
        self.emotion_control = test["emotion_control"]["use"]
        self.content_control = test["content_control"]["use"]
        self.style_Xemo_transfer = test["style_Xemo_transfer"]["use"]
        self.train_pose_framelen = config["DATA_PARAM"]["Bvh"]["train_pose_framelen"]
        wav = tp["wav_dtw_mfcc"]
        self.target_length, self.norm_mean, self.norm_std = wav["target_length"], wav["dataset_mean"], wav["dataset_std"]
        self.num_mel_bins = wav["num_mel_bins"]
        self.seq_len = self.train_pose_framelen
        if self.smplx_rep != "6D" or self.skip_trans or self.train_upper_body or not self.smplx_data:
            raise NotImplementedError("amuse_b200 implements the released configuration: SMPL-X 6D, full body, with trans")
        if self.seq_len != 300:
            raise NotImplementedError("amuse_b200 is built for train_pose_framelen = 300")

    def _apply_ldm_cfg(self, ldm_cfg):
        self.ldm_cfg = ldm_cfg
        sch = ldm_cfg["scheduler"]
        want = {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
                "beta_schedule": "scaled_linear", "set_alpha_to_one": False, "steps_offset": 1}
        for k, v in want.items():
            if sch[k] != v:
                raise NotImplementedError(f"scheduler.{k}={sch[k]!r}: the engine's tables are built for {v!r}")
        self.num_inference_timesteps = sch["num_inference_timesteps"]
        self.eta = sch["eta"]
        self.latent_dim = ldm_cfg["arch_denoiser"]["latent_dim"]
        arch = ldm_cfg["arch_denoiser"]
        fixed = {"ff_size": 512, "num_layers": 9, "num_heads": 4, "arch": "trans_enc", "normalize_before": False,
                 "activation": "gelu", "position_embedding": "learned", "cond_dim": 256, "freq_shift": 0,
                 "ablation_skip_connection": True, "pe_type": "mld", "flip_sin_to_cos": True, "diffusion_only": False}
        for k, v in fixed.items():
            if arch[k] != v:
                raise NotImplementedError(f"arch_denoiser.{k}={arch[k]!r}: kernels are specialised for {v!r}")
        if list(self.latent_dim) != [1, 128]:
            raise NotImplementedError("latent_dim must be [1, 128]")

    def setup(self, config, device, processed, backup_cfg, EXEC_ON_CLUSTER, baseline=False, verbose=False,
              diffonly=False):
        """Same contract as the reference: reads ``configs/<arch>.json`` next to ``processed``, picks the
        LDM / prior / AST checkpoints by the reference's file-name rules, loads them and returns the LDM epoch."""
        self._read_config(config, device, processed, backup_cfg, baseline, diffonly)
        ld = config["TRAIN_PARAM"]["latent_diffusion"]
        if ld["pretrained_prior_lpdm_e"] != ld["pretrained_ldm_lpdm_e"]:
            raise AssertionError("Epochs for prior and ldm should be same")
        if backup_cfg is not None:
            raise NotImplementedError("Backup for LPDM not implemented yet!")
        root = Path(processed).parents[1]
        with open(root / "configs" / f"{ld['arch']}.json", "r") as f:
            self._apply_ldm_cfg(json.load(f))
        model_dir = root / ("saved-models-new" if EXEC_ON_CLUSTER else "saved-models") / ld["pretrained_lpdm"]
        files = [f for f in model_dir.iterdir() if f.is_file() and "experiment_args.json" not in str(f)]
        ldm_file = _pick_by_loss_or_epoch([f for f in files if f.stem.split("_")[0] == "latdiff"],
                                          ld["pretrained_ldm_lpdm_e"])
        ldm_epoch = _first_int(ldm_file.stem.split("_")[-1])
        prior_epoch = ldm_epoch if ld["pretrained_prior_lpdm_e"] == "best" else ld["pretrained_prior_lpdm_e"]
        prior_file = _pick_by_loss_or_epoch([f for f in files if f.stem.split("_")[0] == "prior"], prior_epoch)
        print("[LDM] <===== Chosen LDM model based on total loss: ", ldm_file, " =====>")
        print("[LATDIFF] <===== Chosen VAE model based on total loss: ", prior_file, " =====>")
        ldm_sd = torch.load(ldm_file, map_location="cpu")["model_state_dict"]
        den_sd = {k[len("denoiser."):]: v for k, v in ldm_sd.items() if k.startswith("denoiser.")}
        vae_sd = torch.load(prior_file, map_location="cpu")["model_state_dict"]

        ast_sd = None
        audio_ablation = config["TRAIN_PARAM"]["wav_dtw_mfcc"].get("ablation")
        assert audio_ablation is not None, f"[LPDM EVAL] Audio ablation flag: {audio_ablation}"
        ast_dir = root / "saved-models" / config["TRAIN_PARAM"][self.tag]["pretrained_ast"]
        if ast_dir.is_dir():
            ast_sd = torch.load(self._pick_ast(ast_dir, audio_ablation), map_location="cpu")
        self.frame_based_feats = config["TRAIN_PARAM"]["wav_dtw_mfcc"]["frame_based_feats"]
        self._load_engine(den_sd, vae_sd, ast_sd)
        return ldm_epoch

    @staticmethod
    def _pick_ast(ast_dir: Path, audio_ablation):
        """infer_pretrained_ast_evp.py:16-33: highest emotion (or identity) accuracy in the file name;
        an "epoch 0" winner is replaced by the ``_1_`` file."""
        files = [f for f in ast_dir.iterdir() if f.is_file() and "experiment_args.json" not in str(f)]
        best, best_acc = None, -float("inf")
        for f in files:
            parts = f.stem.split("_")
            acc = _ast_num(parts[4]) if audio_ablation == "identity" else _ast_num(parts[3])
            if acc is not None and acc > best_acc:
                best, best_acc = f, acc
        if int(_ast_num(best.stem.split("_")[1])) == 0:
            best = [f for f in files if "_1_" in str(f)][0]
        return best

    @classmethod
    def from_state_dicts(cls, denoiser_sd, vae_sd, ast_sd=None, device="cuda:0", num_inference_timesteps=50,
                         eta=0.0, sampler="ddim"):
        """Construct directly from in-memory state-dicts (reference key names) -- what tests and
        bench.py use in place of the checkpoint files the reference does not ship."""
        self = cls(base_prior=None)
        self.device = torch.device(device)
        self.diffonly, self.baseline = False, False
        self.smplx_rep, self.seq_len, self.train_pose_framelen = "6D", 300, 300
        self.latent_dim = [1, 128]
        self.num_inference_timesteps, self.eta, self.sampler = num_inference_timesteps, eta, sampler
        self.target_length, self.norm_mean, self.norm_std, self.num_mel_bins = 1024, -9.173025, 5.062332, 128
        self.style_transfer = self.emotion_control = self.content_control = self.style_Xemo_transfer = False
        self.frame_based_feats = True
        self._load_engine(denoiser_sd, vae_sd, ast_sd)
        return self

    def _load_engine(self, den_sd, vae_sd, ast_sd):
        eng = Engine(self.device)
        eng.load_state_dict("denoiser", den_sd)
        eng.load_state_dict("vae", vae_sd)
        if ast_sd is not None:
            enc = {k: v for k, v in ast_sd.items() if k.split(".")[0] in ("emo_enc", "sty_enc", "con_enc")}
            eng.load_state_dict("ast", enc)
        eng.finalize()
        self.engine = eng
        self.has_ast = ast_sd is not None

    # ------------------------------------------------------------------ sampling (infer_ldm.py:130-178)
    def diffusion_backward(self, bsz, z_con, z_emo, z_sty):
        """noise -> N-step denoising loop -> VAE decode -> axis-angle.
        ``z_*`` are [bsz, 256] tensors on ``self.device`` (``z_emo`` / ``z_sty`` may be None);
        returns {"poses": [bsz,300,55,3], "trans": [bsz,300,3]} fp32 on ``self.device``."""
        if z_con.shape[0] != bsz:
            raise ValueError(f"bsz={bsz} but z_con has {z_con.shape[0]} rows")
        if self.diffonly:
            raise RuntimeError("diffonly path is not available (the reference raises here too)")
        # initial noise from the global torch generator of the device, exactly like the reference
        latents = torch.randn((bsz, self.latent_dim[0], self.latent_dim[-1]), device=self.device, dtype=torch.float)
        seed = 0
        if self.sampler == "ddpm" or self.eta > 0:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # Philox stream keyed off torch's CPU generator
        with torch.no_grad():
            out = self.engine.diffusion_backward(latents.view(bsz, -1), z_con, z_emo, z_sty,
                                                 n_steps=self.num_inference_timesteps, sampler=self.sampler,
                                                 eta=float(self.eta), seed=seed)
        return {"poses": out["poses"], "trans": out["trans"]}

    # ------------------------------------------------------------------ audio (infer_ldm.py:180-193)
    def _fbank(self, sliced_chunk):
        import torchaudio
        fbank = torchaudio.compliance.kaldi.fbank(sliced_chunk, htk_compat=True, sample_frequency=16000,
                                                  use_energy=False, window_type="hanning",
                                                  num_mel_bins=self.num_mel_bins, dither=0.0, frame_shift=10)
        pad = self.target_length - fbank.shape[0]
        if pad > 0:
            fbank = torch.nn.functional.pad(fbank, (0, 0, 0, pad))
        elif pad < 0:
            fbank = fbank[: self.target_length, :]
        return (fbank - self.norm_mean) / (self.norm_std * 2)      # normalised AFTER zero padding, as the reference

    def process_single_seq(self, sliced_chunk, framerate=16000 // 2, baseline=False):
        """[C, N] waveform (assumed 16 kHz) -> (con, emo, sty), each [1, 256] on ``self.device``."""
        if not getattr(self, "has_ast", False):
            raise RuntimeError("AST encoder weights were not loaded")
        if self.device_fbank:     # Kaldi fbank + pad + normalise on the GPU (amuse_fbank); channel 0 like kaldi's channel=-1
            fbank = self.engine.fbank(sliced_chunk[:1].to(self.device), self.norm_mean, self.norm_std)[0]
        else:                     # the reference's host path (torchaudio on the CPU)
            fbank = self._fbank(sliced_chunk)
        # (the reference raises here when more than one GPU is visible, infer_pretrained_ast_evp.py:45;
        #  this engine is one-process-per-GPU and has no such restriction)
        con, emo, sty = self.engine.ast_features(fbank.unsqueeze(0))
        return con, emo, sty

    def collect_audio_metrics(self, sliced_chunk, framerate=16000 // 2, baseline=False, tgtpath=None):
        raise NotImplementedError("fbank reconstruction metrics (AST_EVP fusion/decoder heads) are outside the sampling path")

    # ------------------------------------------------------------------ edits (infer_ldm.py:225-414)
    def process_loader(self, data_dict) -> Dict:
        """Dataset-driven edit preparation.  With no edit flag set (the shipped ``edit_gesture`` default,
        scripts/overrides/edit_gesture.yaml) the reference returns an empty dict; the dataset-driven
        branches need ``MotionPrior.encode`` and BEAT data."""
        if self.style_Xemo_transfer or self.style_transfer or self.emotion_control:
            raise NotImplementedError("dataset-driven edits need MotionPrior.encode (not on the sampling path yet)")
        return dict()
