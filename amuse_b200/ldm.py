"""Training-time reuse of the sampler.

During ``train_gesture`` the reference runs the whole reverse process once per training iteration,
under ``torch.no_grad()``, with the denoiser weights of that iteration
(``scripts/trainer.py:413-415`` -> ``LatentDiffusionModel.diffusion_backward``,
``models/latent_diffusion/ldm.py:118-153``) and hands the latents to the frozen
``MotionPrior.decode``.  Training itself (``diffusion_forward``, the losses, the optimiser) is out of
scope; this module only takes that one no-grad call off the PyTorch eager path:

    sampler = LatentDiffusionSampler(ldm.denoiser, ldm.ldm_cfg, device)
    ldm.diffusion_backward = sampler.diffusion_backward          # same signature, same return shape

Every call first re-packs the engine's denoiser weights from the live module when any parameter
changed since the previous call (``amuse_load_weights`` + ``amuse_finalize_weights``, which repacks
only the groups that changed), then runs the persistent denoise-loop kernel.  There is no PyTorch
fallback: without the CUDA library the constructor of the engine raises.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Union

import torch

from .engine import Engine

__all__ = ["LatentDiffusionSampler"]

StateSource = Union[torch.nn.Module, Callable[[], Dict[str, torch.Tensor]]]


class LatentDiffusionSampler:
    """``LatentDiffusionModel.diffusion_backward`` (ldm.py:118-153) on the CUDA engine.

    ``denoiser``: the live ``Denoiser`` module (its ``state_dict()`` keys are the reference's,
    denoiser.py:66-105) or a callable returning such a state-dict.
    ``ldm_cfg``: the dict of ``configs/<arch>.json`` -- only ``scheduler.num_inference_timesteps`` and
    ``scheduler.eta`` are read here (ldm.py:36-37); the other scheduler constants are the ones the
    engine's tables are built for and are checked.
    """

    def __init__(self, denoiser: StateSource, ldm_cfg: dict, device, engine: Optional[Engine] = None,
                 seq_len: int = 300):
        sch = ldm_cfg["scheduler"]
        want = {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
                "beta_schedule": "scaled_linear", "set_alpha_to_one": False, "steps_offset": 1}
        for k, v in want.items():
            if sch[k] != v:
                raise NotImplementedError(f"scheduler.{k}={sch[k]!r}: the engine's tables are built for {v!r}")
        self.num_inference_timesteps = int(sch["num_inference_timesteps"])
        self.eta = float(sch["eta"])
        self.latent_dim = list(ldm_cfg["arch_denoiser"]["latent_dim"])
        if self.latent_dim != [1, 128]:
            raise NotImplementedError("latent_dim must be [1, 128]")
        self.seq_len = seq_len
        self.device = torch.device(device)
        self._source = denoiser
        self._owns_engine = engine is None
        self.engine = engine if engine is not None else Engine(self.device)
        self._stamp = None
        self.refreshes = 0

    # ------------------------------------------------------------------ weights
    def _state_dict(self) -> Dict[str, torch.Tensor]:
        src = self._source
        sd = src.state_dict() if isinstance(src, torch.nn.Module) else src()
        return {k: v for k, v in sd.items() if k != "mem_pos.pe"}   # never read by forward (denoiser.py:174-188)

    @staticmethod
    def _version_stamp(sd: Dict[str, torch.Tensor]):
        # in-place optimiser updates bump ``_version``; a re-assigned tensor changes ``data_ptr``
        return tuple((k, v.data_ptr(), v._version) for k, v in sd.items())

    def refresh(self, force: bool = False) -> bool:
        """Push the current denoiser weights into the engine.  Returns True when a repack happened."""
        sd = self._state_dict()
        stamp = self._version_stamp(sd)
        if not force and stamp == self._stamp:
            return False
        # one device->host copy for all ~230 tensors instead of one synchronous copy each
        flat = torch.cat([v.detach().reshape(-1).to(torch.float32) for v in sd.values()]).cpu()
        off = 0
        for k, v in sd.items():
            n = v.numel()
            self.engine.load_tensor(f"denoiser.{k}", flat[off:off + n].view(v.shape))
            off += n
        self.engine.finalize()
        self._stamp = stamp
        self.refreshes += 1
        return True

    # ------------------------------------------------------------------ the call (ldm.py:118-153)
    def diffusion_backward(self, ld_audio_con, ld_audio_emo, ld_audio_sty, ld_audio_mfcc, bsz):
        if ld_audio_mfcc is not None:
            raise NotImplementedError("LPDM: Baseline audio AE not implemented yet")   # ldm.py:124
        if ld_audio_con.shape[0] != bsz:
            raise ValueError(f"bsz={bsz} but ld_audio_con has {ld_audio_con.shape[0]} rows")
        with torch.no_grad():
            self.refresh()
            # same draw as the reference: global generator of `device`, shape (bsz, 1, 128); init_noise_sigma = 1
            latents = torch.randn((bsz, self.latent_dim[0], self.latent_dim[-1]), device=self.device, dtype=torch.float)
            step_noise = None
            if self.eta > 0:   # DDIM variance noise (diffusers draws it per step from the global generator)
                step_noise = torch.randn((self.num_inference_timesteps, bsz, 128), device=self.device, dtype=torch.float)
            z = self.engine.denoise(latents, ld_audio_con, ld_audio_emo, ld_audio_sty,
                                    n_steps=self.num_inference_timesteps, sampler="ddim", eta=self.eta,
                                    step_noise=step_noise)
        return z.view(bsz, 1, 128).permute(1, 0, 2)   # ldm.py:152

    def close(self):
        if self._owns_engine and self.engine is not None:
            self.engine.close()
        self.engine = None
