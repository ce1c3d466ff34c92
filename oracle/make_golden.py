"""Generate tests/golden/*.npz by running the REFERENCE'S OWN modules on seeded inputs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only
(``python -m oracle.make_golden``): it imports ``Denoiser`` / ``MotionPrior`` /
``dm.utils.transforms`` from /root/reference through oracle/reference_loader.py,
loads the synthetic state-dicts of oracle/weights.py into them, drives them with the
restated scheduler of oracle/lpdm_ref.py (diffusers is not installed) and stores
inputs + outputs.  The reference cannot travel to the GPU box; these fixtures can.

Every fixture carries fp32 outputs (what the reference computes) and fp64 outputs
(same modules, ``.double()`` -- the rounding-free yardstick used to state tolerances),
plus SHA-1 checksums of the synthetic weights so a test can prove it regenerated the
same tensors.  Big tensors are stored on a frame subsample (``frame_idx``).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

from . import lpdm_ref as R
from . import reference_loader as L
from . import weights as W

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"
FRAME_IDX = np.arange(0, 300, 7)          # 43 of the 300 frames


class RefPath:
    """The reference modules wired like ``PretrainedLPDM_v1.diffusion_backward``
    (infer_ldm.py:130-178), noise injected, scheduler restated."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.dsd, self.vsd = W.denoiser_state_dict(), W.motionprior_state_dict()
        self.den = L.load_denoiser(self.dsd).to(dtype)
        self.vae = L.load_motionprior(self.vsd).to(dtype)
        self.tf = L.load_transforms()

    def c(self, x):
        return None if x is None else x.to(self.dtype)

    @torch.no_grad()
    def eps(self, x, t, con, emo, sty):
        B = x.shape[0]
        u = lambda z: None if z is None else self.c(z)[:, None, :]
        return self.den(sample=self.c(x)[:, None, :], timestep=torch.tensor(t), con_hidden=u(con),
                        emo_hidden=u(emo), sty_hidden=u(sty), lengths=[300] * B)[0][:, 0, :]

    @torch.no_grad()
    def sample(self, l0, con, emo, sty, n, sampler, noise=None):
        plan = R.ddim_coeffs(n) if sampler == "ddim" else R.ddpm_coeffs(n)
        x = self.c(l0)
        for i, t in enumerate(plan["timesteps"]):
            e = self.eps(x, t, con, emo, sty)
            x = R.scheduler_step(plan, i, x, e, None if noise is None else self.c(noise[i]))
        return x

    @torch.no_grad()
    def decode(self, z):
        B = z.shape[0]
        feats = self.vae.decode(self.c(z)[None], [300] * B)                      # vae.py:216-278
        rot6d = feats[:, :, :-3].reshape(B, 300, 55, 6)                          # infer_ldm.py:167-168
        poses = self.tf.matrix_to_axis_angle(self.tf.rotation_6d_to_matrix(rot6d))
        return feats, poses, feats[:, :, -3:]


def synthetic_motion(B):
    """Closed-form smooth SMPL-X motion (axis-angle up to ~0.9 rad, a few exactly-zero joints for the
    small-angle branch): poses [B,300,55,3], trans [B,300,3]."""
    b = torch.arange(B, dtype=torch.float64)[:, None, None, None]
    t = torch.arange(300, dtype=torch.float64)[None, :, None, None]
    j = torch.arange(55, dtype=torch.float64)[None, None, :, None]
    c = torch.arange(3, dtype=torch.float64)[None, None, None, :]
    poses = 0.5 * torch.sin(0.05 * t + 0.7 * j + 1.3 * c + b) + 0.2 * torch.cos(0.011 * t * (c + 1) - 0.3 * j)
    poses[:, :, 7] = 0.0
    trans = 0.1 * torch.cos(0.03 * t[:, :, 0] + c[:, :, 0] + b[:, :, 0])
    return poses.to(torch.float32), trans.to(torch.float32)


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def main():
    assert L.available(), "needs /root/reference"
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    r32, r64 = RefPath(torch.float32), RefPath(torch.float64)
    meta = dict(denoiser_sha1=W.checksum(r32.dsd), motionprior_sha1=W.checksum(r32.vsd),
                torch_version=torch.__version__)
    f32 = lambda t: t.detach().to(torch.float32).numpy()
    f64 = lambda t: t.detach().to(torch.float64).numpy()

    # G1 -- single denoiser evaluations (a2-a7), incl. the None-condition ablations
    B = 3
    x, con, emo, sty = randn(11, B, 128), randn(12, B, 256), randn(13, B, 256), randn(14, B, 256)
    g = dict(x=f32(x), con=f32(con), emo=f32(emo), sty=f32(sty), timesteps=np.array([981, 500, 1]))
    for t in (981, 500, 1):
        g[f"eps_t{t}_f32"] = f32(r32.eps(x, t, con, emo, sty))
        g[f"eps_t{t}_f64"] = f64(r64.eps(x, t, con, emo, sty))
    g["eps_t981_noemo_nosty_f32"] = f32(r32.eps(x, 981, con, None, None))
    g["eps_t981_noemo_nosty_f64"] = f64(r64.eps(x, 981, con, None, None))
    g["eps_t981_nosty_f32"] = f32(r32.eps(x, 981, con, emo, None))
    g["eps_t981_nosty_f64"] = f64(r64.eps(x, 981, con, emo, None))
    np.savez_compressed(OUT / "denoiser_step.npz", **g, **meta)

    # G2/G3 -- DDIM eta=0: 50 steps (shipped config), 1 step (config 1); latents only
    for n, B, seed in ((50, 4, 20), (1, 1, 30), (50, 1, 40)):
        l0, con, emo, sty = randn(seed, B, 128), randn(seed + 1, B, 256), randn(seed + 2, B, 256), randn(seed + 3, B, 256)
        g = dict(latents0=f32(l0), con=f32(con), emo=f32(emo), sty=f32(sty), n_steps=n,
                 z_f32=f32(r32.sample(l0, con, emo, sty, n, "ddim")),
                 z_f64=f64(r64.sample(l0, con, emo, sty, n, "ddim")))
        np.savez_compressed(OUT / f"ddim{n}_b{B}.npz", **g, **meta)

    # G4 -- DDPM ancestral 1000 steps with injected noise (regenerated from noise_seed)
    n, B, seed = 1000, 2, 50
    l0, con, emo, sty = randn(seed, B, 128), randn(seed + 1, B, 256), randn(seed + 2, B, 256), randn(seed + 3, B, 256)
    noise = randn(seed + 4, n, B, 128)
    g = dict(latents0=f32(l0), con=f32(con), emo=f32(emo), sty=f32(sty), n_steps=n, noise_seed=seed + 4,
             noise_probe=f32(noise[[0, 499, 999]]),
             z_f32=f32(r32.sample(l0, con, emo, sty, n, "ddpm", noise)),
             z_f64=f64(r64.sample(l0, con, emo, sty, n, "ddpm", noise)))
    np.savez_compressed(OUT / "ddpm1000_b2.npz", **g, **meta)
    # a shorter ancestral chain for quick checks
    n = 100
    noise = randn(seed + 5, n, B, 128)
    g = dict(latents0=f32(l0), con=f32(con), emo=f32(emo), sty=f32(sty), n_steps=n, noise_seed=seed + 5,
             z_f32=f32(r32.sample(l0, con, emo, sty, n, "ddpm", noise)),
             z_f64=f64(r64.sample(l0, con, emo, sty, n, "ddpm", noise)))
    np.savez_compressed(OUT / "ddpm100_b2.npz", **g, **meta)

    # G5 -- MotionPrior.decode + 6D -> axis-angle on given latents (a9-a11)
    B = 2
    z = randn(60, B, 128)
    fe32, po32, tr32 = r32.decode(z)
    fe64, po64, tr64 = r64.decode(z)
    g = dict(z=f32(z), frame_idx=FRAME_IDX,
             feats_f32=f32(fe32[:, FRAME_IDX]), feats_f64=f64(fe64[:, FRAME_IDX]),
             poses_f32=f32(po32[:, FRAME_IDX]), poses_f64=f64(po64[:, FRAME_IDX]))
    np.savez_compressed(OUT / "decode_b2.npz", **g, **meta)

    # G6 -- the whole diffusion_backward (a1), DDIM 50, B=2
    B, seed = 2, 70
    l0, con, emo, sty = randn(seed, B, 128), randn(seed + 1, B, 256), randn(seed + 2, B, 256), randn(seed + 3, B, 256)
    z32 = r32.sample(l0, con, emo, sty, 50, "ddim")
    z64 = r64.sample(l0, con, emo, sty, 50, "ddim")
    fe32, po32, _ = r32.decode(z32)
    fe64, po64, _ = r64.decode(z64)
    g = dict(latents0=f32(l0), con=f32(con), emo=f32(emo), sty=f32(sty), n_steps=50, frame_idx=FRAME_IDX,
             z_f32=f32(z32), z_f64=f64(z64),
             feats_f32=f32(fe32[:, FRAME_IDX]), feats_f64=f64(fe64[:, FRAME_IDX]),
             poses_f32=f32(po32[:, FRAME_IDX]), poses_f64=f64(po64[:, FRAME_IDX]))
    np.savez_compressed(OUT / "backward_ddim50_b2.npz", **g, **meta)

    # G7 -- rotation conversion corner cases (a11): identity, near-pi, degenerate 6D input
    d6 = randn(80, 64, 6)
    d6[0] = torch.tensor([1., 0, 0, 0, 1, 0])                  # identity -> small-angle branch
    d6[1] = torch.tensor([-1., 0, 0, 0, -1, 0])                # rotation by pi about z
    d6[2] = torch.tensor([1., 0, 0, 0, -1, 0])                 # rotation by pi about x
    d6[3] = torch.tensor([1e-3, 0, 0, 0, 2e-3, 0])             # tiny magnitudes (normalize)
    d6[4] = torch.tensor([1., 1e-4, 0, 1e-4, 1, 0])            # near identity
    aa32 = r32.tf.matrix_to_axis_angle(r32.tf.rotation_6d_to_matrix(d6))
    aa64 = r64.tf.matrix_to_axis_angle(r64.tf.rotation_6d_to_matrix(d6.double()))
    np.savez_compressed(OUT / "rot6d_cases.npz", d6=f32(d6), aa_f32=f32(aa32), aa_f64=f64(aa64), **meta)

    # G8 -- MotionPrior.encode (section 8f rank 3) on closed-form motion (regenerated by the tests from the
    # same formula: no input storage), through the reference's own axis-angle -> 6D conversion
    poses, trans = synthetic_motion(2)
    out = {}
    for tag, r, cast in (("f32", r32, f32), ("f64", r64, f64)):
        pz = r.c(poses)
        rot6 = r.tf.matrix_to_rotation_6d(r.tf.axis_angle_to_matrix(pz)).reshape(2, 300, 330)    # infer_ldm.py:457-460
        feats = torch.cat((rot6, r.c(trans)), dim=-1)
        with torch.no_grad():
            _, dist = r.vae.encode(feats, [300] * 2)                                               # vae.py:154-214
        out[f"feats_{tag}"] = cast(feats[:, FRAME_IDX])
        out[f"mu_{tag}"] = cast(dist.loc[0])
        out[f"std_{tag}"] = cast(dist.scale[0])
    np.savez_compressed(OUT / "encode_b2.npz", frame_idx=FRAME_IDX, **out, **meta)

    for p in sorted(OUT.glob("*.npz")):
        print(f"{p.name:32s} {p.stat().st_size/1024:8.1f} KiB")


if __name__ == "__main__":
    sys.exit(main())
