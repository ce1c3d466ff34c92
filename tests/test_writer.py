"""Output writer: format pinned on the reference's shipped fixture (keys, dtypes, shapes), plus the
v1 / v2 pose manipulations of visualizer.py:192-225,344-364."""
from pathlib import Path

import numpy as np
import pytest
import torch

from amuse_b200 import writer

FIXTURES = sorted(Path("/root/reference/viz_dump").rglob("*_motion_smplx.npz")) if Path("/root/reference").is_dir() else []
# what the shipped fixtures look like (recorded here so the test also runs where /root/reference is absent)
SPEC = {"poses": ((300, 55, 3), "float32"), "trans": ((300, 3), "float64"), "gender": ((), "<U"),
        "betas": ((300,), "float64"), "mocap_frame_rate": ((), "float64")}


def _feats(B=2, T=300, seed=0):
    g = torch.Generator().manual_seed(seed)
    return writer.feats_from_motion(torch.randn(B, T, 55, 3, generator=g), torch.randn(B, T, 3, generator=g))


def test_v1_file_matches_fixture_format(tmp_path):
    files = writer.write_batch(tmp_path, _feats(), subject="scott", version="v1")
    d = np.load(files[0], allow_pickle=True)
    assert set(d.files) == set(SPEC)
    for k, (shape, dt) in SPEC.items():
        assert d[k].shape == shape and str(d[k].dtype).startswith(dt), k
    assert float(d["mocap_frame_rate"]) == 30.0 and np.abs(d["trans"]).max() == 0.0
    p = d["poses"]
    assert np.array_equal(p[:, writer.LOWER_BODY_JOINTS], np.broadcast_to(p[0:1, writer.LOWER_BODY_JOINTS], (300, 8, 3)))
    assert np.abs(p[:, 3] - p[0:1, 3]).max() > 0        # upper body untouched


@pytest.mark.skipif(not FIXTURES, reason="/root/reference not present")
def test_spec_equals_shipped_fixture():
    d = np.load(FIXTURES[0], allow_pickle=True)
    assert set(d.files) == set(SPEC)
    for k, (shape, dt) in SPEC.items():
        assert d[k].shape == shape and str(d[k].dtype).startswith(dt), k
    p = d["poses"]                                       # the shipped sample was written by v1: lower body locked, trans 0
    assert np.abs(p[:, writer.LOWER_BODY_JOINTS] - p[0:1, writer.LOWER_BODY_JOINTS]).max() == 0.0
    assert np.abs(d["trans"]).max() == 0.0


def test_v2_variants():
    f = _feats(1)[0].numpy()
    poses, trans = writer.prepare_v2(f)
    assert np.abs(poses[:, writer.JAW_JOINT]).max() == 0 and np.allclose(trans, f.reshape(300, 56, 3)[:, -1])
    # dtypes as the reference's v2 writer produces them (visualizer.py:192-222): the jaw-zeroing concatenate with np.zeros
    # promotes poses to float64; a kept trans is the float32 slice, a zeroed one is np.zeros (float64)
    assert poses.dtype == np.float64 and trans.dtype == np.float32
    assert writer.prepare_v2(f, zero_trans=True)[1].dtype == np.float64
    assert writer.prepare_v1(f)[0].dtype == np.float32 and writer.prepare_v1(f)[1].dtype == np.float64
    poses, trans = writer.prepare_v2(f, zero_trans=True, freeze_init_lobody=True)
    assert np.abs(trans).max() == 0 and np.abs(poses[:, 1] - poses[0, 1]).max() == 0
    poses, trans = writer.prepare_v2(f, zero_trans=True)
    assert np.abs(poses[:, 1] - poses[0, 1]).max() > 0
    poses, trans = writer.prepare_v2(f, half_body=True)
    assert np.abs(trans).max() == 0 and np.abs(poses[:, 11] - poses[0, 11]).max() == 0
    with pytest.raises(AssertionError):
        writer.prepare_v2(f[:, :165])
