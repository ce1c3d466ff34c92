import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box without a CUDA device: plain `pytest tests` stays green in the
    build container, `-m gpu` on the GPU box runs them.  A box WITH a GPU but without the built library still fails
    loudly inside the tests (AmuseLibraryError) -- there is no fallback to skip to."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def synthetic_weights():
    """Seeded synthetic state-dicts with the reference's key names (oracle/weights.py)."""
    from oracle import weights as W
    return {"denoiser": W.denoiser_state_dict(), "vae": W.motionprior_state_dict()}


@pytest.fixture(scope="session")
def engine(synthetic_weights):
    """The product: the CUDA engine behind the C ABI, loaded with the synthetic weights."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from amuse_b200.engine import Engine
    eng = Engine("cuda:0")
    eng.load_state_dict("denoiser", synthetic_weights["denoiser"])
    eng.load_state_dict("vae", synthetic_weights["vae"])
    eng.finalize()
    yield eng
    eng.close()
