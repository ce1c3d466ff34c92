"""the training-time caller of the sampler
(``LatentDiffusionModel.diffusion_backward``, reference ldm.py:118-153; trainer.py:413-415)."""
import pytest
import torch

from amuse_b200.ldm import LatentDiffusionSampler
from oracle import lpdm_ref as R

LDM_CFG = {"scheduler": {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
                         "beta_schedule": "scaled_linear", "set_alpha_to_one": False, "steps_offset": 1,
                         "num_inference_timesteps": 50, "eta": 0.0},
           "arch_denoiser": {"latent_dim": [1, 128]}}


class _LiveDenoiser(torch.nn.Module):
    """Stand-in for the reference ``Denoiser`` being trained: parameters under the reference's
    state-dict keys, updated in place like an optimiser does."""

    def __init__(self, sd, device):
        super().__init__()
        self._keys = list(sd.keys())
        self.params = torch.nn.ParameterList([torch.nn.Parameter(v.clone().to(device)) for v in sd.values()])

    def state_dict(self, *a, **k):
        return {key: p.detach() for key, p in zip(self._keys, self.params)}

    @torch.no_grad()
    def optimiser_step(self, seed):
        g = torch.Generator().manual_seed(seed)
        for p in self.params:
            p.add_((0.02 * torch.randn(p.shape, generator=g)).to(p.device))


class _FakeEngine:
    def __init__(self):
        self.loaded, self.finalized = {}, 0

    def load_tensor(self, name, t):
        assert t.device.type == "cpu" and t.dtype == torch.float32
        self.loaded[name] = t.clone()

    def finalize(self):
        self.finalized += 1

    def close(self):
        pass


def test_refresh_only_when_weights_change(synthetic_weights):
    live = _LiveDenoiser(synthetic_weights["denoiser"], "cpu")
    eng = _FakeEngine()
    s = LatentDiffusionSampler(live, LDM_CFG, "cpu", engine=eng)
    assert s.refresh() and eng.finalized == 1
    assert "denoiser.mem_pos.pe" not in eng.loaded                       # never read by forward
    for k, v in live.state_dict().items():
        if k != "mem_pos.pe":
            assert torch.equal(eng.loaded[f"denoiser.{k}"], v), k         # flat copy sliced back correctly
    assert not s.refresh() and eng.finalized == 1                         # unchanged -> no repack
    live.optimiser_step(1)
    assert s.refresh() and eng.finalized == 2
    assert torch.equal(eng.loaded["denoiser.encoder.norm.weight"], live.state_dict()["encoder.norm.weight"])


def test_rejects_what_the_reference_rejects(synthetic_weights):
    s = LatentDiffusionSampler(lambda: synthetic_weights["denoiser"], LDM_CFG, "cpu", engine=_FakeEngine())
    with pytest.raises(NotImplementedError):                              # ldm.py:124
        s.diffusion_backward(torch.zeros(2, 256), None, None, torch.zeros(2, 13), 2)
    bad = {"scheduler": dict(LDM_CFG["scheduler"], beta_schedule="linear"), "arch_denoiser": LDM_CFG["arch_denoiser"]}
    with pytest.raises(NotImplementedError):
        LatentDiffusionSampler(lambda: {}, bad, "cpu", engine=_FakeEngine())


@pytest.mark.gpu
def test_training_loop_reuse_matches_oracle(synthetic_weights):
    """Three 'training iterations': sample, optimiser step (in-place update), sample again with the
    refreshed weights, sample a third time without a change.  Each result is compared with the CPU
    restatement run on the weights of that iteration and the same initial noise."""
    dev = "cuda:0"
    live = _LiveDenoiser(synthetic_weights["denoiser"], dev)
    s = LatentDiffusionSampler(live, LDM_CFG, dev)
    try:
        B = 6
        g = torch.Generator().manual_seed(11)
        con, emo, sty = (torch.randn(B, 256, generator=g) for _ in range(3))
        outs = []
        for it, (e, st) in enumerate([(emo, sty), (emo, sty), (emo, None)]):
            if it == 1:
                live.optimiser_step(seed=3)
            torch.manual_seed(100 + it)
            z = s.diffusion_backward(con.to(dev), None if e is None else e.to(dev), None if st is None else st.to(dev),
                                     None, B)
            assert z.shape == (1, B, 128) and z.device.type == "cuda"     # ldm.py:152 layout
            torch.manual_seed(100 + it)
            l0 = torch.randn((B, 1, 128), device=dev, dtype=torch.float).cpu().view(B, 128)
            sd = {k: v.cpu() for k, v in live.state_dict().items()}
            ref = R.sample_latents(sd, l0, con, e, st, 50, "ddim")
            err = (z[0].cpu() - ref).abs().max().item()
            print(f"[parity] train-time sampler iteration {it}: max|d|={err:.3e}")
            assert err < 2e-4
            outs.append(z[0].cpu())
        assert s.refreshes == 2                                            # iteration 2 reused the packed weights
        assert (outs[0] - outs[1]).abs().max().item() > 1e-3               # the update really reached the kernel
    finally:
        s.close()
