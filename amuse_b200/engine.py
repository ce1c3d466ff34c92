"""Tensor-level wrapper around the C ABI: torch owns device memory and streams, the library
owns the math.  All tensors are fp32, contiguous and live on the engine's CUDA device."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def host_alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    """``alphas_cumprod`` exactly as diffusers builds it on the host for ``scaled_linear``
    (linspace of sqrt-betas in fp32, squared, cumprod) -- reference call site infer_ldm.py:116-123."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def host_sinusoid_freqs(dim: int = 256, freq_shift: float = 0.0) -> torch.Tensor:
    """fp32 frequency table of ``get_timestep_embedding`` (utils/embeddings.py:264-270)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    return torch.exp(exponent / (half - freq_shift))


class Engine:
    def __init__(self, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AmuseLibraryError("amuse_b200 runs on a CUDA device (B200, sm_100a) only")
        if not torch.cuda.is_available():
            raise _lib.AmuseLibraryError("no CUDA device visible: amuse_b200 has no CPU path")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        h = C.c_void_p()
        rc = self.lib.amuse_create(C.byref(h), idx)
        if rc != 0:
            raise _lib.AmuseError(rc, "amuse_create failed (is this an sm_100 device?)")
        self._h = h
        self._finalized = False

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.amuse_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _lib.AmuseError(rc, self.lib.amuse_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t: Optional[torch.Tensor], shape=None) -> Optional[torch.Tensor]:
        if t is None:
            return None
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    # ------------------------------------------------------------------ weights
    def load_tensor(self, name: str, t: torch.Tensor):
        t = t.detach().to(dtype=torch.float32).contiguous()
        shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
        self._check(self.lib.amuse_load_weights(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), 0))

    def load_state_dict(self, prefix: str, sd: Dict[str, torch.Tensor]):
        """``prefix`` in {"denoiser", "vae", "ast"}; keys are the reference's state-dict keys."""
        for k, v in sd.items():
            self.load_tensor(f"{prefix}.{k}", v)

    def finalize(self, exact_host_tables: bool = True):
        if exact_host_tables:   # hand the engine the tables torch produces on this host (bit-parity with diffusers)
            self.load_tensor("scheduler.alphas_cumprod", host_alphas_cumprod())
            self.load_tensor("denoiser.time_proj.freqs", host_sinusoid_freqs())
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_finalize_weights(self._h, self._stream()))
        self._finalized = True

    def _empty(self, *shape) -> torch.Tensor:
        return torch.empty(*shape, device=self.device, dtype=torch.float32)

    def reserve(self, max_clips: int, max_steps: int = 1000):
        self._check(self.lib.amuse_reserve(self._h, max_clips, max_steps))

    # ------------------------------------------------------------------ compute
    def schedule(self, n_steps: int, sampler: str = "ddim", eta: float = 0.0):
        ts = (C.c_int32 * n_steps)()
        cf = (C.c_float * (5 * n_steps))()
        self._check(self.lib.amuse_schedule(self._h, n_steps, _lib.SAMPLER[sampler], eta, ts, cf))
        return list(ts), torch.tensor(list(cf), dtype=torch.float32).view(n_steps, 5)

    def denoise(self, latents0, z_con, z_emo=None, z_sty=None, n_steps=50, sampler="ddim", eta=0.0,
                clip_sample=None, step_noise=None, seed=0, clip_offset=0) -> torch.Tensor:
        """``clip_offset``: global index of the first clip of this call (Philox noise is keyed by global clip index,
        so shards of a batch reproduce the un-sharded run; see include/amuse_b200.h)."""
        B = latents0.shape[0]
        if B == 0:   # an empty batch has nothing to launch; the reference returns empty tensors as well
            return self._empty(0, 128)
        l0 = self._dev(latents0.reshape(B, 128), (B, 128))
        con, emo, sty = self._dev(z_con, (B, 256)), self._dev(z_emo, (B, 256)), self._dev(z_sty, (B, 256))
        noise = self._dev(step_noise, (n_steps, B, 128)) if step_noise is not None else None
        out = torch.empty(B, 128, device=self.device, dtype=torch.float32)
        clip = -1 if clip_sample is None else int(bool(clip_sample))
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_denoise(self._h, B, n_steps, _lib.SAMPLER[sampler], eta, clip, _ptr(l0),
                                               _ptr(con), _ptr(emo), _ptr(sty), _ptr(noise), seed, clip_offset, _ptr(out),
                                               self._stream()))
        return out

    def denoiser_eps(self, sample, timestep: int, z_con, z_emo=None, z_sty=None) -> torch.Tensor:
        B = sample.shape[0]
        x = self._dev(sample.reshape(B, 128), (B, 128))
        con, emo, sty = self._dev(z_con, (B, 256)), self._dev(z_emo, (B, 256)), self._dev(z_sty, (B, 256))
        out = torch.empty(B, 128, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_denoiser_eps(self._h, B, int(timestep), _ptr(x), _ptr(con), _ptr(emo),
                                                    _ptr(sty), _ptr(out), self._stream()))
        return out

    def decode(self, latents, want_feats=False):
        B = latents.shape[0]
        if B == 0:
            e = (self._empty(0, 300, 55, 3), self._empty(0, 300, 3))
            return e + (self._empty(0, 300, 333),) if want_feats else e
        z = self._dev(latents.reshape(B, 128), (B, 128))
        poses = torch.empty(B, 300, 55, 3, device=self.device, dtype=torch.float32)
        trans = torch.empty(B, 300, 3, device=self.device, dtype=torch.float32)
        feats = torch.empty(B, 300, 333, device=self.device, dtype=torch.float32) if want_feats else None
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_decode(self._h, B, _ptr(z), _ptr(feats), _ptr(poses), _ptr(trans),
                                              self._stream()))
        return (poses, trans, feats) if want_feats else (poses, trans)

    def encode(self, feats: torch.Tensor):
        """``MotionPrior.encode`` without the rsample draw: feats [B,300,333] -> (mu [B,128], logvar [B,128])."""
        B = feats.shape[0]
        if B == 0:
            return self._empty(0, 128), self._empty(0, 128)
        x = self._dev(feats, (B, 300, 333))
        mu, logvar = (torch.empty(B, 128, device=self.device, dtype=torch.float32) for _ in range(2))
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_encode(self._h, B, _ptr(x), _ptr(mu), _ptr(logvar), self._stream()))
        return mu, logvar

    def motion_to_feats(self, poses: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
        """poses [...,55,3] axis-angle + trans [...,3] -> [...,333] (55 x 6D | trans), infer_ldm.py:454-461."""
        p, t = self._dev(poses), self._dev(trans)
        lead = tuple(p.shape[:-2])
        if p.shape[-2:] != (55, 3) or tuple(t.shape) != lead + (3,):
            raise ValueError("poses must be [...,55,3] and trans [...,3]")
        n = p.numel() // 165
        out = torch.empty(*lead, 333, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_motion_to_feats(self._h, n, _ptr(p), _ptr(t), _ptr(out), self._stream()))
        return out

    def rot6d_to_axis_angle(self, d6: torch.Tensor) -> torch.Tensor:
        x = self._dev(d6)
        n = x.numel() // 6
        out = torch.empty(*x.shape[:-1], 3, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_rot6d_to_axis_angle(self._h, n, _ptr(x), _ptr(out), self._stream()))
        return out

    def diffusion_backward(self, latents0, z_con, z_emo=None, z_sty=None, n_steps=50, sampler="ddim", eta=0.0,
                           clip_sample=None, step_noise=None, seed=0, want_latents=False, want_feats=False, clip_offset=0):
        """Device tensors in, device tensors out: {"poses": [B,300,55,3], "trans": [B,300,3]}."""
        B = latents0.shape[0]
        if B == 0:
            out = {"poses": self._empty(0, 300, 55, 3), "trans": self._empty(0, 300, 3)}
            if want_latents:
                out["latents"] = self._empty(0, 128)
            if want_feats:
                out["feats"] = self._empty(0, 300, 333)
            return out
        l0 = self._dev(latents0.reshape(B, 128), (B, 128))
        con, emo, sty = self._dev(z_con, (B, 256)), self._dev(z_emo, (B, 256)), self._dev(z_sty, (B, 256))
        noise = self._dev(step_noise, (n_steps, B, 128)) if step_noise is not None else None
        poses = torch.empty(B, 300, 55, 3, device=self.device, dtype=torch.float32)
        trans = torch.empty(B, 300, 3, device=self.device, dtype=torch.float32)
        lat = torch.empty(B, 128, device=self.device, dtype=torch.float32) if want_latents else None
        feats = torch.empty(B, 300, 333, device=self.device, dtype=torch.float32) if want_feats else None
        clip = -1 if clip_sample is None else int(bool(clip_sample))
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_diffusion_backward(
                self._h, B, n_steps, _lib.SAMPLER[sampler], eta, clip, _ptr(l0), _ptr(con), _ptr(emo), _ptr(sty),
                _ptr(noise), seed, clip_offset, _ptr(lat), _ptr(feats), _ptr(poses), _ptr(trans), self._stream()))
        out = {"poses": poses, "trans": trans}
        if want_latents:
            out["latents"] = lat
        if want_feats:
            out["feats"] = feats
        return out

    def diffusion_backward_host(self, latents0, z_con, z_emo=None, z_sty=None, n_steps=50, sampler="ddim", eta=0.0,
                                clip_sample=None, step_noise=None, seed=0, out_poses=None, out_trans=None, clip_offset=0):
        """HOST tensors in (ideally pinned), HOST tensors out; copies are inside the call."""
        B = latents0.shape[0]
        if B == 0:
            return {"poses": torch.empty(0, 300, 55, 3, dtype=torch.float32), "trans": torch.empty(0, 300, 3, dtype=torch.float32)}
        h = lambda t: None if t is None else t.detach().to(dtype=torch.float32).contiguous()
        l0, con, emo, sty, noise = h(latents0.reshape(B, 128)), h(z_con), h(z_emo), h(z_sty), h(step_noise)
        for t in (l0, con, emo, sty, noise):
            if t is not None and t.device.type != "cpu":
                raise ValueError("diffusion_backward_host takes host tensors")
        poses = out_poses if out_poses is not None else torch.empty(B, 300, 55, 3, dtype=torch.float32).pin_memory()
        trans = out_trans if out_trans is not None else torch.empty(B, 300, 3, dtype=torch.float32).pin_memory()
        clip = -1 if clip_sample is None else int(bool(clip_sample))
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_diffusion_backward_host(
                self._h, B, n_steps, _lib.SAMPLER[sampler], eta, clip, _ptr(l0), _ptr(con), _ptr(emo), _ptr(sty),
                _ptr(noise), seed, clip_offset, _ptr(poses), _ptr(trans), self._stream()))
        return {"poses": poses, "trans": trans}

    def fbank(self, wave: torch.Tensor, norm_mean: float = -9.173025, norm_std: float = 5.062332) -> torch.Tensor:
        """[B, n_samples] 16 kHz mono waveforms -> [B, 1024, 128] normalised Kaldi log-mel filterbank."""
        x = self._dev(wave)
        if x.dim() != 2:
            raise ValueError("wave must be [B, n_samples]")
        B, n = x.shape
        out = torch.empty(B, 1024, 128, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_fbank(self._h, B, n, _ptr(x), norm_mean, norm_std, _ptr(out), self._stream()))
        return out

    def ast_features(self, fbank: torch.Tensor):
        B = fbank.shape[0]
        if B == 0:
            return self._empty(0, 256), self._empty(0, 256), self._empty(0, 256)
        x = self._dev(fbank, (B, 1024, 128))
        con, emo, sty = (torch.empty(B, 256, device=self.device, dtype=torch.float32) for _ in range(3))
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_ast_features(self._h, B, _ptr(x), _ptr(con), _ptr(emo), _ptr(sty),
                                                    self._stream()))
        return con, emo, sty

    def debug_tc_gemm(self, epi: int, A, W, bias, A2=None, R=None, ln=None, cvec=None, rows_per_clip=1):
        """Unit-test hook for the tcgen05 3xTF32 GEMM: epi(A [| A2] . W^T + bias) -> [M, N] fp32."""
        A, W, bias = self._dev(A), self._dev(W), self._dev(bias)
        A2, R, ln, cvec = self._dev(A2), self._dev(R), self._dev(ln), self._dev(cvec)
        M, K1 = A.shape
        K = K1 + (A2.shape[1] if A2 is not None else 0)
        N = W.shape[0]
        out = torch.empty(M, N, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_debug_tc_gemm(self._h, epi, M, N, K, _ptr(A), _ptr(A2), K1, _ptr(W), _ptr(bias),
                                                     _ptr(R), _ptr(ln), _ptr(cvec), rows_per_clip, _ptr(out),
                                                     self._stream()))
        return out

    def philox_normals(self, seed: int, B: int, n_steps: int, clip_offset: int = 0) -> torch.Tensor:
        """The N(0,1) draws the sampler makes in-kernel for ``step_noise=None``: [n_steps, B, 128] (debug / test hook)."""
        out = torch.empty(n_steps, B, 128, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.amuse_debug_philox_normals(self._h, seed, clip_offset, B, n_steps, _ptr(out),
                                                            self._stream()))
        return out

    # ------------------------------------------------------------------ introspection
    def launch_count(self) -> int:
        return int(self.lib.amuse_launch_count(self._h))

    def profile_arm(self, step: int):
        self._check(self.lib.amuse_profile_arm(self._h, step))

    def profile_read(self, n: int = 96):
        buf = (C.c_int64 * n)()
        torch.cuda.synchronize(self.device)
        self._check(self.lib.amuse_profile_read(self._h, buf, n))
        return list(buf)
