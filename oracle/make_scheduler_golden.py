"""Golden vectors for the scheduler math, produced by the REFERENCE'S OWN GaussianDiffusion class
(models/diffusion/utils/mdm_gaussian_diffusion.py) -- run in the build container, where /root/reference exists:

    python -m oracle.make_scheduler_golden        ->  tests/golden/scheduler_gd.npz

TEST INFRASTRUCTURE (see oracle/__init__.py).  What is stored (float64, the class computes its tables in float64):
  * the posterior tables of the 1000-step scaled-linear schedule (configs/diff_latent_v2.json:48-56);
  * one ancestral DDPM step at several timesteps:  x0 = _predict_xstart_from_eps (:528), mean / variance =
    q_posterior_mean_variance (:343-366), sample = mean + [t > 0] exp(0.5 log_var) noise  (p_sample :634-700,
    ModelVarType.FIXED_SMALL);
  * one deterministic DDIM step (eta = 0) of the 50-step "leading" schedule with steps_offset 1 and
    set_alpha_to_one False, through a GaussianDiffusion re-spaced on the timesteps [0, 1, 21, ..., 981] so that its
    (alphas_cumprod, alphas_cumprod_prev) pairs are exactly diffusers' (alpha_t, alpha_prev): Equation 12 of
    ddim_sample (:927-941).  Inputs are scaled so that clamp(x0, -1, 1) is inactive: the reference class re-derives
    eps from the CLIPPED x0 (:925) while diffusers 0.17.1 (use_clipped_model_output False) keeps the model's eps, so
    the two only agree where the clamp does nothing.
"""
from pathlib import Path

import numpy as np
import torch

from oracle import reference_loader as RL


def scaled_linear_betas(n=1000, b0=0.00085, b1=0.012):
    # diffusers 0.17.1 "scaled_linear": linspace(sqrt(b0), sqrt(b1), n, float32) ** 2
    return (torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2).double().numpy()


def build(gd_mod, betas):
    return gd_mod.GaussianDiffusion(betas=betas, model_mean_type=gd_mod.ModelMeanType.EPSILON,
                                    model_var_type=gd_mod.ModelVarType.FIXED_SMALL, loss_type=gd_mod.LossType.MSE)


def main():
    gd_mod = RL.load_gaussian_diffusion()
    betas = scaled_linear_betas()
    gd = build(gd_mod, betas)
    out = {"betas": betas, "alphas_cumprod": gd.alphas_cumprod, "posterior_mean_coef1": gd.posterior_mean_coef1,
           "posterior_mean_coef2": gd.posterior_mean_coef2, "posterior_variance": gd.posterior_variance,
           "sqrt_recip_alphas_cumprod": gd.sqrt_recip_alphas_cumprod,
           "sqrt_recipm1_alphas_cumprod": gd.sqrt_recipm1_alphas_cumprod}
    g = torch.Generator().manual_seed(17)
    x = torch.randn(3, 128, generator=g, dtype=torch.float64)
    eps = torch.randn(3, 128, generator=g, dtype=torch.float64)
    noise = torch.randn(3, 128, generator=g, dtype=torch.float64)
    out.update(x=x.numpy(), eps=eps.numpy(), noise=noise.numpy())
    # the class casts its tables to float32 in _extract_into_tensor (:1793); evaluate the same methods in float64 by
    # patching that one helper, so that the golden is the exact math and not one particular rounding of it
    def extract64(arr, timesteps, broadcast_shape):
        res = torch.from_numpy(arr)[timesteps].double()
        while len(res.shape) < len(broadcast_shape):
            res = res[..., None]
        return res.expand(broadcast_shape)
    gd_mod._extract_into_tensor = extract64
    ddpm_t = [999, 500, 37, 1, 0]
    steps = []
    for t in ddpm_t:
        tt = torch.full((3,), t, dtype=torch.long)
        x0 = gd._predict_xstart_from_eps(x, tt, eps)
        mean, var, logvar = gd.q_posterior_mean_variance(x0, x, tt)
        sample = mean + (0.0 if t == 0 else 1.0) * torch.exp(0.5 * logvar) * noise
        steps.append(sample.numpy())
    out["ddpm_t"] = np.array(ddpm_t)
    out["ddpm_sample"] = np.stack(steps)
    # DDIM, 50 steps, eta = 0: re-spaced class
    seq = [0] + [i * 20 + 1 for i in range(50)]
    ac = gd.alphas_cumprod[seq]
    sp_betas = np.concatenate([[1.0 - ac[0]], 1.0 - ac[1:] / ac[:-1]])
    sp = build(gd_mod, sp_betas)
    assert np.allclose(sp.alphas_cumprod, ac, rtol=1e-13) and np.allclose(sp.alphas_cumprod_prev[1:], ac[:-1], rtol=1e-13)
    xs, es = 0.004 * x, 0.004 * eps        # |x0| < 1: the clamp is inactive
    ddim_idx = [50, 25, 2, 1]              # positions in `seq`: timesteps 981, 481, 21, 1
    outs = []
    for j in ddim_idx:
        tt = torch.full((3,), j, dtype=torch.long)
        x0 = gd_mod.GaussianDiffusion._predict_xstart_from_eps(sp, xs, tt, es).clamp(-1, 1)
        assert x0.abs().max() < 1.0
        e2 = sp._predict_eps_from_xstart(xs, tt, x0)                       # ddim_sample :925
        abar = extract64(sp.alphas_cumprod, tt, xs.shape)                  # :927
        abar_prev = extract64(sp.alphas_cumprod_prev, tt, xs.shape)        # :928
        sigma = 0.0 * torch.sqrt((1 - abar_prev) / (1 - abar)) * torch.sqrt(1 - abar / abar_prev)   # :929-933, eta = 0
        outs.append((x0 * torch.sqrt(abar_prev) + torch.sqrt(1 - abar_prev - sigma ** 2) * e2).numpy())   # :936-939
    out["ddim_t"] = np.array([seq[j] for j in ddim_idx])
    out["ddim_scale"] = np.array(0.004)
    out["ddim_sample"] = np.stack(outs)
    dst = Path(__file__).resolve().parents[1] / "tests" / "golden" / "scheduler_gd.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
