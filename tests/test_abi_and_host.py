"""CPU tests: the C-ABI library loads and exports every symbol of include/amuse_b200.h; the host-side
mirror of the reference class keeps the reference's surface and file-selection rules."""
import ctypes as C
import re
import sys
import types
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_exported():
    from amuse_b200 import _lib
    hdr = (ROOT / "include" / "amuse_b200.h").read_text()
    declared = set(re.findall(r"\b(amuse_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    lib = _lib.load()
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.amuse_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    from amuse_b200 import _lib
    from amuse_b200.engine import Engine
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.amuse_create(C.byref(h), 0) != 0 and not h.value      # error code, no context, no crash
    with pytest.raises(_lib.AmuseLibraryError):
        Engine("cuda:0")                                             # no CPU / PyTorch fallback
    with pytest.raises(_lib.AmuseLibraryError):
        Engine("cpu")


def test_product_does_not_import_oracle():
    for p in (ROOT / "amuse_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_mapinfo2takes():
    from amuse_b200.infer_ldm import mapinfo2takes
    assert mapinfo2takes("[ayana-scott]_[fear]") == ["0_103_103", "0_104_104"]
    assert mapinfo2takes("[yingqing]_[happy]_first") == ["0_65_65", "0_66_66"]
    assert mapinfo2takes("angry", trainer=True) == ["0_73_73", "0_74_74"]
    with pytest.raises(Exception, match="Unknown emotion"):      # the reference's fall-through (infer_ldm.py:528)
        mapinfo2takes("[a-b]_[neutral]")


def test_checkpoint_selection_rules(tmp_path):
    from amuse_b200 import infer_ldm as M
    names = ["latdiff_x_total0.412_e200.pt", "latdiff_x_total0.398_e400.pt", "latdiff_x_total0.405_e600.pt",
             "prior_x_total0.9_e200.pt", "prior_x_total0.8_e400.pt", "experiment_args.json"]
    for n in names:
        (tmp_path / n).write_bytes(b"")
    files = [f for f in tmp_path.iterdir() if f.is_file() and "experiment_args.json" not in str(f)]
    ldm = [f for f in files if f.stem.split("_")[0] == "latdiff"]
    assert M._pick_by_loss_or_epoch(ldm, "best").name == "latdiff_x_total0.398_e400.pt"
    assert M._pick_by_loss_or_epoch(ldm, "600").name == "latdiff_x_total0.405_e600.pt"
    pri = [f for f in files if f.stem.split("_")[0] == "prior"]
    assert M._pick_by_loss_or_epoch(pri, 400).name == "prior_x_total0.8_e400.pt"
    ast = tmp_path / "ast"
    ast.mkdir()
    for n in ["wav_7_x_tEA0.81_tPA0.99.pkl", "wav_0_x_tEA0.95_tPA0.10.pkl", "wav_1_x_tEA0.60_tPA0.70.pkl"]:
        (ast / n).write_bytes(b"")
    assert M.PretrainedLPDM_v1._pick_ast(ast, "full").name == "wav_1_x_tEA0.60_tPA0.70.pkl"   # epoch-0 winner -> "_1_"
    assert M.PretrainedLPDM_v1._pick_ast(ast, "identity").name == "wav_7_x_tEA0.81_tPA0.99.pkl"


def test_reference_surface_and_shadowing(tmp_path, monkeypatch):
    """The recipe of INTEGRATION.md: a stub package ahead of the reference on sys.path makes
    `from models.latent_diffusion.infer_ldm import PretrainedLPDM_v1, mapinfo2takes` resolve to us."""
    import inspect
    pkg = tmp_path / "models" / "latent_diffusion"
    pkg.mkdir(parents=True)
    (tmp_path / "models" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("")
    (pkg / "infer_ldm.py").write_text("from amuse_b200.infer_ldm import *  # noqa\n")
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(str(tmp_path))
    mod = __import__("models.latent_diffusion.infer_ldm", fromlist=["x"])
    from amuse_b200 import infer_ldm as ours
    assert mod.PretrainedLPDM_v1 is ours.PretrainedLPDM_v1 and mod.mapinfo2takes is ours.mapinfo2takes
    sig = inspect.signature(ours.PretrainedLPDM_v1.setup)
    assert list(sig.parameters) == ["self", "config", "device", "processed", "backup_cfg", "EXEC_ON_CLUSTER",
                                    "baseline", "verbose", "diffonly"]
    assert list(inspect.signature(ours.PretrainedLPDM_v1.diffusion_backward).parameters) == \
        ["self", "bsz", "z_con", "z_emo", "z_sty"]
    assert list(inspect.signature(ours.PretrainedLPDM_v1.__init__).parameters) == \
        ["self", "base_prior", "base_con_ae", "base_emo_ae", "base_audio_ae"]
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        sys.modules.pop(k, None)


def test_host_tables_match_oracle():
    from amuse_b200.engine import host_alphas_cumprod, host_sinusoid_freqs
    from oracle import lpdm_ref as R
    assert torch.equal(host_alphas_cumprod(), R.alphas_cumprod())
    assert torch.equal(host_sinusoid_freqs(), R.sinusoid_freqs())


def test_shard_range():
    from amuse_b200.shard import shard_range
    for n, w in ((512, 8), (256, 8), (70, 8), (5, 8), (64, 1)):
        got = [shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(got[i][1] == got[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1


def test_setup_fixture_config_matches_the_reference_file():
    """The configs/diff_latent_v2.json keys that tests/test_gpu_setup.py writes into its temporary tree are the released
    file's (checked here, where /root/reference exists; the GPU box has no reference tree)."""
    import importlib.util
    import json
    from oracle import reference_loader as RL
    if not RL.available():
        pytest.skip("/root/reference is only present in the build container")
    spec = importlib.util.spec_from_file_location("_setup_fixture", Path(__file__).parent / "test_gpu_setup.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = json.loads((RL.REF / "configs" / "diff_latent_v2.json").read_text())
    for sec in ("arch_denoiser", "noisy_scheduler", "scheduler"):
        assert mod.LDM_CFG[sec] == ref[sec], sec
