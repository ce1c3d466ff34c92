"""CPU restatement of ``PretrainedLPDM_v1.diffusion_backward`` and everything below it.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- plain PyTorch on CPU, dtype-generic
(fp32 to mirror the reference, fp64 as the rounding-free yardstick).  Functional
style: every function takes the state-dict (reference key names, oracle/weights.py).
All citations are relative to /root/reference.

Pinned against the reference's own modules by tests/test_oracle_vs_reference.py
(Denoiser, MotionPrior.decode, rotation conversions).  The scheduler maths restates
diffusers==0.17.1 (not installed anywhere in the container): PARITY UNPINNED for
``ddim_*`` / ``ddpm_*`` (call sites infer_ldm.py:116-125,142-161).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

NHEAD = 4
LN_EPS = 1e-5
BLOCKS_IN = ["input_blocks.%d" % i for i in range(4)]
BLOCKS_OUT = ["output_blocks.%d" % i for i in range(4)]


def cast_sd(sd: SD, dtype) -> SD:
    return {k: v.to(dtype) for k, v in sd.items()}


# --------------------------------------------------------------------------- pieces
def timestep_sinusoid(t: Tensor, dim: int = 256, flip_sin_to_cos: bool = True,
                      freq_shift: float = 0.0, dtype=torch.float32) -> Tensor:
    """``get_timestep_embedding`` (models/latent_diffusion/utils/embeddings.py:245-285).
    Frequencies, product, sin and cos are all fp32 (the reference calls ``.float()`` and only
    casts the finished embedding to the sample dtype), also when the modules run in fp64."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    freqs = torch.exp(exponent)                                   # fp32, as reference
    emb = t[:, None].float() * freqs[None, :]                     # fp32 product even for fp64 modules
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb.to(dtype)                                          # denoiser.py:148 ``.to(dtype=sample.dtype)``


def sinusoid_freqs(dim: int = 256, freq_shift: float = 0.0) -> Tensor:
    """The fp32 frequency table alone (embeddings.py:264-270) -- what the product
    hands to the engine as the hoisted ``time_proj.freqs`` table."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    return torch.exp(exponent / (half - freq_shift))


def mha(q_in: Tensor, kv_in: Tensor, sd: SD, p: str) -> Tensor:
    """``nn.MultiheadAttention(128, 4)`` forward, batch-first restatement:
    packed in_proj rows [q;k;v], heads = contiguous 32-wide slices, scale 1/sqrt(32)."""
    w, b = sd[f"{p}.in_proj_weight"], sd[f"{p}.in_proj_bias"]
    d = w.shape[1]
    hd = d // NHEAD
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(kv_in, w[d:2 * d], b[d:2 * d])
    v = F.linear(kv_in, w[2 * d:], b[2 * d:])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    q = q.view(B, Tq, NHEAD, hd).transpose(1, 2)
    k = k.view(B, Tk, NHEAD, hd).transpose(1, 2)
    v = v.view(B, Tk, NHEAD, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Tq, d)
    return F.linear(o, sd[f"{p}.out_proj.weight"], sd[f"{p}.out_proj.bias"])


def _ln(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"], sd[f"{p}.bias"], LN_EPS)


def _ffn(x: Tensor, sd: SD, p: str) -> Tensor:
    h = F.gelu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"]))   # erf GELU
    return F.linear(h, sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])


def encoder_layer(x: Tensor, sd: SD, p: str) -> Tensor:
    """``TransformerEncoderLayer.forward_post`` (utils/cross_attention.py:259-272)."""
    x = _ln(x + mha(x, x, sd, f"{p}.self_attn"), sd, f"{p}.norm1")
    return _ln(x + _ffn(x, sd, p), sd, f"{p}.norm2")


def decoder_layer(x: Tensor, mem: Tensor, sd: SD, p: str) -> Tensor:
    """``TransformerDecoderLayer.forward_post`` (utils/cross_attention.py:323-345)."""
    x = _ln(x + mha(x, x, sd, f"{p}.self_attn"), sd, f"{p}.norm1")
    x = _ln(x + mha(x, mem, sd, f"{p}.multihead_attn"), sd, f"{p}.norm2")
    return _ln(x + _ffn(x, sd, p), sd, f"{p}.norm3")


def skip_encoder(x: Tensor, sd: SD, p: str) -> Tensor:
    """``SkipTransformerEncoder.forward`` (utils/cross_attention.py:41-64)."""
    xs: List[Tensor] = []
    for b in BLOCKS_IN:
        x = encoder_layer(x, sd, f"{p}.{b}")
        xs.append(x)
    x = encoder_layer(x, sd, f"{p}.middle_block")
    for i, b in enumerate(BLOCKS_OUT):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = F.linear(x, sd[f"{p}.linear_blocks.{i}.weight"], sd[f"{p}.linear_blocks.{i}.bias"])
        x = encoder_layer(x, sd, f"{p}.{b}")
    return _ln(x, sd, f"{p}.norm")


def skip_decoder(x: Tensor, mem: Tensor, sd: SD, p: str) -> Tensor:
    """``SkipTransformerDecoder.forward`` (utils/cross_attention.py:89-125)."""
    xs: List[Tensor] = []
    for b in BLOCKS_IN:
        x = decoder_layer(x, mem, sd, f"{p}.{b}")
        xs.append(x)
    x = decoder_layer(x, mem, sd, f"{p}.middle_block")
    for i, b in enumerate(BLOCKS_OUT):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = F.linear(x, sd[f"{p}.linear_blocks.{i}.weight"], sd[f"{p}.linear_blocks.{i}.bias"])
        x = decoder_layer(x, mem, sd, f"{p}.{b}")
    return _ln(x, sd, f"{p}.norm")


# --------------------------------------------------------------------------- denoiser
def time_token(sd: SD, t: Tensor, dtype) -> Tensor:
    """``Timesteps`` + ``TimestepEmbedding`` (embeddings.py:288-322): Linear-SiLU-Linear."""
    e = timestep_sinusoid(t, 256, True, 0.0, dtype).to(sd["time_embedding.linear_1.weight"].device)
    e = F.linear(e, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    e = F.silu(e)
    return F.linear(e, sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])


def cond_tokens(sd: SD, con: Tensor, emo: Optional[Tensor], sty: Optional[Tensor]) -> Tensor:
    """``emb_proj_{con,emo,sty}`` = ReLU -> Linear(256,128) (denoiser.py:74-79,153-171).
    Inputs [B,256]; returns [B, n_cond, 128] in the reference token order con, emo, sty."""
    toks = []
    for name, z in (("con", con), ("emo", emo), ("sty", sty)):
        if z is None:
            continue
        toks.append(F.linear(F.relu(z), sd[f"emb_proj_{name}.1.weight"], sd[f"emb_proj_{name}.1.bias"]))
    return torch.stack(toks, dim=1)


def denoiser_forward(sd: SD, sample: Tensor, t: int, con: Tensor, emo: Optional[Tensor],
                     sty: Optional[Tensor]) -> Tensor:
    """``Denoiser.forward`` (models/latent_diffusion/denoiser.py:135-204), trans_enc arch,
    diffusion_only False.  ``sample`` [B,128], conditions [B,256]; returns eps [B,128].
    Token order z, t, con, emo, sty (denoiser.py:174,180); learned PE added to every token
    (position_encoding.py:138-159)."""
    dtype = sample.dtype
    B = sample.shape[0]
    tt = time_token(sd, torch.full((B,), t, dtype=torch.int64), dtype)           # [B,128]
    x = torch.cat([sample[:, None, :], tt[:, None, :], cond_tokens(sd, con, emo, sty)], dim=1)
    T = x.shape[1]
    x = x + sd["query_pos.pe"][:T, 0, :][None]
    x = skip_encoder(x, sd, "encoder")
    return x[:, 0, :]


# --------------------------------------------------------------------------- schedulers
def alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                   dtype=torch.float32) -> Tensor:
    """``scaled_linear`` betas (diffusers 0.17.1): linspace(sqrt(b0), sqrt(b1), N, fp32)**2,
    cumprod(1-betas) in fp32 -- kept in fp32 like the library, then cast."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(dtype)


def ddim_timesteps(n: int, num_train: int = 1000, steps_offset: int = 1) -> List[int]:
    """``DDIMScheduler.set_timesteps`` ("leading" spacing): (arange(n)*ratio)[::-1] + offset.
    n=50 -> 981, 961, ..., 1 (infer_ldm.py:142-143)."""
    r = num_train // n
    ts = [i * r + steps_offset for i in range(n)][::-1]
    if ts[0] >= num_train:
        raise IndexError(f"DDIM timestep {ts[0]} indexes alphas_cumprod[{num_train}] "
                         "(the reference hits the same IndexError, SURVEY App. C)")
    return ts


def ddim_coeffs(n: int, clip_sample: bool = True) -> Dict[str, object]:
    """Per-step scalars of ``DDIMScheduler.step`` with eta = 0:
    x0 = (x - sqrt(1-a) e)/sqrt(a); x0 = clamp(x0,-1,1); x' = sqrt(a') x0 + sqrt(1-a') e.
    set_alpha_to_one False -> final_alpha = alphas_cumprod[0].  The scalars are computed
    with fp32 0-dim tensor arithmetic in the library's own order (``1 - a`` rounded in
    fp32 before the sqrt), so the table is bit-identical to what the library multiplies by.
    Columns: sqrt(a), sqrt(1-a), sqrt(a'), sqrt(1-a'), 0."""
    ac = alphas_cumprod(dtype=torch.float32)
    ts = ddim_timesteps(n)
    r = 1000 // n
    rows = []
    for t in ts:
        p = t - r
        a = ac[t]
        ap = ac[p] if p >= 0 else ac[0]
        rows.append(torch.stack([a ** 0.5, (1 - a) ** 0.5, ap ** 0.5, (1 - ap) ** 0.5,
                                 torch.zeros((), dtype=torch.float32)]))
    return {"timesteps": ts, "coef": torch.stack(rows), "clip": clip_sample, "sampler": "ddim"}


def ddpm_timesteps(n: int, num_train: int = 1000) -> List[int]:
    """``DDPMScheduler.set_timesteps``: (arange(n) * (N // n))[::-1], no offset (App. B.2)."""
    r = num_train // n
    return [i * r for i in range(n)][::-1]


def ddpm_coeffs(n: int, clip_sample: bool = False) -> Dict[str, object]:
    """Per-step scalars of ``DDPMScheduler.step`` (epsilon prediction, fixed_small variance,
    clip_sample False -- configs/diff_latent_v2.json:48-56), fp32 scalar
    arithmetic in the library's order.  x0 = (x - sqrt(1-a) e)/sqrt(a);
    x' = c0 x0 + cx x + [t>0] sigma z.   Columns: sqrt(a), sqrt(1-a), c0, cx, sigma."""
    ac = alphas_cumprod(dtype=torch.float32)
    one = torch.ones((), dtype=torch.float32)
    ts = ddpm_timesteps(n)
    r = 1000 // n
    rows = []
    for t in ts:
        p = t - r
        a = ac[t]
        ap = ac[p] if p >= 0 else one
        beta_prod, beta_prod_prev = 1 - a, 1 - ap
        alpha_t = a / ap
        beta_t = 1 - alpha_t
        c0 = (ap ** 0.5 * beta_t) / beta_prod
        cx = alpha_t ** 0.5 * beta_prod_prev / beta_prod
        var = torch.clamp((1 - ap) / (1 - a) * beta_t, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.zeros((), dtype=torch.float32)
        rows.append(torch.stack([a ** 0.5, beta_prod ** 0.5, c0, cx, sigma]))
    return {"timesteps": ts, "coef": torch.stack(rows), "clip": clip_sample, "sampler": "ddpm"}


def scheduler_step(plan: Dict[str, object], i: int, x: Tensor, eps: Tensor,
                   noise: Optional[Tensor]) -> Tensor:
    c = plan["coef"][i].to(x.dtype)
    x0 = (x - c[1] * eps) / c[0]
    if plan["clip"]:
        x0 = x0.clamp(-1.0, 1.0)
    if plan["sampler"] == "ddim":
        return c[2] * x0 + c[3] * eps
    out = c[2] * x0 + c[3] * x
    if noise is not None:
        out = out + c[4] * noise
    return out


def sample_latents(sd: SD, latents0: Tensor, con: Tensor, emo: Optional[Tensor], sty: Optional[Tensor],
                   n_steps: int = 50, sampler: str = "ddim", step_noise: Optional[Tensor] = None,
                   clip_sample: Optional[bool] = None) -> Tensor:
    """The loop of ``diffusion_backward`` (infer_ldm.py:137-161).  ``latents0`` [B,128]
    (init_noise_sigma = 1).  ``step_noise`` [n_steps,B,128] for the ancestral sampler."""
    if sampler == "ddim":
        plan = ddim_coeffs(n_steps, True if clip_sample is None else clip_sample)
    else:
        plan = ddpm_coeffs(n_steps, False if clip_sample is None else clip_sample)
    x = latents0
    for i, t in enumerate(plan["timesteps"]):
        eps = denoiser_forward(sd, x, t, con, emo, sty)
        x = scheduler_step(plan, i, x, eps, None if step_noise is None else step_noise[i])
    return x


# --------------------------------------------------------------------------- VAE decoder
def vae_decode(sd: SD, z: Tensor, nframes: int = 300) -> Tensor:
    """``MotionPrior.decode`` (models/latent_diffusion/vae.py:216-278), encoder_decoder arch,
    pe_type mld, all lengths = nframes (mask all-true => the zeroing at :274 is a no-op).
    ``z`` [B,128] (one latent token per clip) -> feats [B,nframes,333]."""
    B = z.shape[0]
    q = torch.zeros(B, nframes, z.shape[-1], dtype=z.dtype, device=z.device) + sd["query_pos_decoder.pe"][:nframes, 0, :][None]
    x = skip_decoder(q, z[:, None, :], sd, "decoder")
    return F.linear(x, sd["final_layer.weight"], sd["final_layer.bias"])


# --------------------------------------------------------------------------- VAE encoder
def vae_encode(sd: SD, feats: Tensor):
    """Deterministic part of ``MotionPrior.encode`` (models/latent_diffusion/vae.py:154-214), arch
    encoder_decoder, pe_type mld, MLP_DIST false, latent_size 1, all lengths = nframes (mask all-true):
    feats [B,T,333] -> (mu [B,128], logvar [B,128]).  The reference then draws
    ``latent = Normal(mu, exp(logvar)**0.5).rsample()`` from the global torch generator (vae.py:210-213);
    that draw stays on the host side of the boundary."""
    B = feats.shape[0]
    x = F.linear(feats, sd["skel_embedding.weight"], sd["skel_embedding.bias"])        # vae.py:169
    dist = sd["global_motion_token"][None].expand(B, -1, -1)                           # vae.py:176
    xseq = torch.cat((dist, x), dim=1)                                                 # vae.py:185
    xseq = xseq + sd["query_pos_encoder.pe"][: xseq.shape[1], 0, :][None]               # vae.py:191
    out = skip_encoder(xseq, sd, "encoder")[:, :2]                                      # vae.py:192-193
    return out[:, 0], out[:, 1]                                                         # vae.py:205-206


# --------------------------------------------------------------------------- rotations
def axis_angle_to_rot6d(aa: Tensor) -> Tensor:
    """``matrix_to_rotation_6d(axis_angle_to_matrix(aa))`` (dm/utils/transforms.py:228-257 -> 96-124 ->
    211-226): axis-angle [...,3] -> quaternion -> matrix -> first two rows [...,6]."""
    ang = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = 0.5 * ang
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    q = torch.cat([torch.cos(half), aa * s], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    return torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)), -1)


def motion_to_feats(poses: Tensor, trans: Tensor) -> Tensor:
    """``_loader_helper_v1`` motion branch (infer_ldm.py:454-462): poses [B,T,55,3] axis-angle + trans
    [B,T,3] -> the VAE's [B,T,333] feature layout (55 x 6D, then trans)."""
    B, T = poses.shape[:2]
    return torch.cat((axis_angle_to_rot6d(poses).reshape(B, T, 330), trans), dim=-1)



def rotation_6d_to_matrix(d6: Tensor) -> Tensor:
    """Gram-Schmidt, rows (b1,b2,b3) (dm/utils/transforms.py:187-208; F.normalize eps 1e-12)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def matrix_to_quaternion(m: Tensor) -> Tensor:
    """sqrt-positive-part + copysign variant (dm/utils/transforms.py:259-309)."""
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]

    def sp(x):
        return torch.sqrt(torch.clamp(x, min=0.0))

    def cs(a, b):
        return torch.where((a < 0) != (b < 0), -a, a)

    o0 = 0.5 * sp(1 + m00 + m11 + m22)
    x = 0.5 * sp(1 + m00 - m11 - m22)
    y = 0.5 * sp(1 - m00 + m11 - m22)
    z = 0.5 * sp(1 - m00 - m11 + m22)
    return torch.stack((o0, cs(x, m[..., 2, 1] - m[..., 1, 2]), cs(y, m[..., 0, 2] - m[..., 2, 0]),
                        cs(z, m[..., 1, 0] - m[..., 0, 1])), -1)


def quaternion_to_axis_angle(q: Tensor) -> Tensor:
    """dm/utils/transforms.py:156-184."""
    norms = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return q[..., 1:] / s


def rot6d_to_axis_angle(d6: Tensor) -> Tensor:
    return quaternion_to_axis_angle(matrix_to_quaternion(rotation_6d_to_matrix(d6)))


def feats_to_motion(feats: Tensor) -> Dict[str, Tensor]:
    """Tail of ``diffusion_backward`` (infer_ldm.py:165-174): [B,T,333] -> poses [B,T,55,3], trans [B,T,3]."""
    B, T, _ = feats.shape
    rot6d = feats[:, :, :-3].reshape(B, T, 55, 6)
    return {"poses": rot6d_to_axis_angle(rot6d), "trans": feats[:, :, -3:]}


def geodesic_deg(aa_a: Tensor, aa_b: Tensor) -> Tensor:
    """Rotation angle (degrees) between two axis-angle fields -- the 2*pi-wrap-safe pose metric."""
    def to_R(aa):
        ang = aa.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        ax = aa / ang
        K = torch.zeros(*aa.shape[:-1], 3, 3, dtype=aa.dtype)
        K[..., 0, 1], K[..., 0, 2] = -ax[..., 2], ax[..., 1]
        K[..., 1, 0], K[..., 1, 2] = ax[..., 2], -ax[..., 0]
        K[..., 2, 0], K[..., 2, 1] = -ax[..., 1], ax[..., 0]
        s, c = torch.sin(ang)[..., None], torch.cos(ang)[..., None]
        return torch.eye(3, dtype=aa.dtype) + s * K + (1 - c) * (K @ K)
    R = to_R(aa_a.double()).transpose(-1, -2) @ to_R(aa_b.double())
    tr = (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1) / 2
    return torch.rad2deg(torch.acos(tr.clamp(-1, 1)))


def gram_schmidt_cond(feats: Tensor) -> Tensor:
    """1 / min(|a1|, |a2 - (b1.a2) b1|) per rotation: how much Gram-Schmidt (dm/utils/transforms.py:141-160) amplifies an
    error of the 6D features into an error of the rotation.  [..., 333] -> [..., 55]."""
    d6 = feats[..., :330].double().reshape(*feats.shape[:-1], 55, 6)
    a1, a2 = d6[..., :3], d6[..., 3:]
    n1 = a1.norm(dim=-1)
    b1 = a1 / n1[..., None]
    n2 = (a2 - (b1 * a2).sum(-1, keepdim=True) * b1).norm(dim=-1)
    return 1.0 / torch.minimum(n1, n2)


def check_poses(poses: Tensor, ref_poses: Tensor, ref_feats: Tensor, feats_err: float, tol_deg: float = 0.05):
    """Pose parity through the rotation geodesic, conditioning-aware: a rotation whose 6D vectors are short is
    ill-conditioned (synthetic weights give |a| down to ~0.01), so its bound is the measured 6D feature error carried
    through the conditioning of Gram-Schmidt, on top of `tol_deg` (the fp32 matrix -> quaternion -> axis-angle chain alone
    moves well-conditioned rotations by up to 0.02 deg between fp32 and fp64 evaluations of this very restatement).
    The strict form -- well-conditioned rotations against fp64 goldens -- is tests/test_gpu_parity.py::_assert_poses.
    Returns (ok, geodesic max, geodesic max over the well-conditioned rotations, cond < 4)."""
    geo = geodesic_deg(poses, ref_poses)
    cond = gram_schmidt_cond(ref_feats)
    bound = torch.rad2deg(3.0 * feats_err * 6 ** 0.5 * cond) + tol_deg
    well = geo[cond < 4].max().item() if bool((cond < 4).any()) else 0.0
    return bool((geo <= bound).all()), geo.max().item(), well


# --------------------------------------------------------------------------- whole path
def diffusion_backward(den_sd: SD, vae_sd: SD, latents0: Tensor, con: Tensor, emo: Optional[Tensor],
                       sty: Optional[Tensor], n_steps: int = 50, sampler: str = "ddim",
                       step_noise: Optional[Tensor] = None, nframes: int = 300) -> Dict[str, Tensor]:
    """``PretrainedLPDM_v1.diffusion_backward`` (infer_ldm.py:130-178) with the initial noise
    (and the ancestral noise) passed in so that the run is reproducible."""
    with torch.no_grad():
        z = sample_latents(den_sd, latents0, con, emo, sty, n_steps, sampler, step_noise)
        feats = vae_decode(vae_sd, z, nframes)
        out = feats_to_motion(feats)
    out["latents"] = z
    out["feats"] = feats
    return out
