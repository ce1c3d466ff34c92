"""Import the reference's OWN hot-path modules from /root/reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` does not exist on
the GPU box, so nothing that runs there may call into this file; it is used by
``oracle/make_golden.py`` (fixture generation) and by the ``not gpu`` test that pins the
restatement in ``lpdm_ref.py`` to the reference code.

Recipe: ``models/__init__.py`` instantiates timm models and
``models/latent_diffusion/__init__.py`` imports diffusers, neither of which is
installed, so the two package ``__init__`` files are bypassed by pre-registering
stub packages whose ``__path__`` points at the reference directories; the real
``denoiser.py`` / ``vae.py`` / ``dm/utils/transforms.py`` then import unmodified.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types
from pathlib import Path

REF = Path(os.environ.get("AMUSE_REFERENCE", "/root/reference"))


def available() -> bool:
    return (REF / "models" / "latent_diffusion" / "denoiser.py").is_file()


def _stub_packages():
    for name, path in [("models", REF / "models"),
                       ("models.latent_diffusion", REF / "models" / "latent_diffusion"),
                       ("dm", REF / "dm"), ("dm.utils", REF / "dm" / "utils")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [str(path)]
            sys.modules[name] = m


def load_denoiser(state_dict=None):
    """The reference ``Denoiser`` built exactly as ``PretrainedLPDM_v1.setup`` does
    (infer_ldm.py:66-73) from configs/diff_latent_v2.json, 6D SMPL-X data."""
    _stub_packages()
    from models.latent_diffusion.denoiser import Denoiser        # noqa: the reference's file
    cfg = json.load(open(REF / "configs" / "diff_latent_v2.json"))
    dc = dict(cfg["arch_denoiser"])
    dc["smplx_data"] = True
    dc["smplx_rep"] = "6D"
    with contextlib.redirect_stdout(io.StringIO()):
        m = Denoiser(dc)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.eval()


def load_motionprior(state_dict=None):
    """The reference ``MotionPrior`` after ``setup`` (vae.py:25-146; main.py:140-142)."""
    _stub_packages()
    from models.latent_diffusion.vae import MotionPrior          # noqa
    base = json.load(open(REF / "configs" / "base_new.json"))
    m = MotionPrior()
    with contextlib.redirect_stdout(io.StringIO()):
        m.setup(REF / "data" / "BEAT-processed", base)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.eval()


def load_transforms():
    """The vendored pytorch3d rotation conversions (dm/utils/transforms.py:141-309)."""
    _stub_packages()
    import importlib
    return importlib.import_module("dm.utils.transforms")


def load_gaussian_diffusion():
    """The reference's vendored ``GaussianDiffusion`` module (models/diffusion/utils/mdm_gaussian_diffusion.py:181):
    the one piece of reference-held code that contains the DDPM posterior (q_posterior_mean_variance :343-366,
    _predict_xstart_from_eps :528) and the DDIM update (ddim_sample :895-945).  It is not on the reference's sampling
    path (that path calls diffusers, which is not installable here) but it is the same published math, so it pins the
    scheduler restatement in ``lpdm_ref.py``.  Imports with torch / numpy / einops only."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_amuse_ref_gaussian_diffusion", REF / "models" / "diffusion" / "utils" / "mdm_gaussian_diffusion.py")
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod
