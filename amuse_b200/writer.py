"""Output writer: SMPL-X ``*_motion_smplx.npz`` files in the format the
reference's Blender pipeline consumes (``CaMNVisualizer.animate_ldm_sample_v1`` / ``_v2``,
models/diffusion/viz/visualizer.py:344-364 and :192-225).  The caller (trainer.py:524-532) hands over
``feats [B,300,168]`` = 55 axis-angle joints (165) followed by the root translation (3).

v1 (``infer_gesture`` and the demo edit):  drop trans, freeze the 8 lower-body joints to frame 0, zero trans.
v2 (dataset edits): zero the jaw joint 22; keep trans unless ``zero_trans`` / ``half_body``; the lower body
   is frozen only for ``half_body`` or ``zero_trans and freeze_init_LoBody``.
Keys / dtypes follow the reference's writers: v1 (the shipped fixtures viz_dump/test/**/*_motion_smplx.npz) poses f32
[T,55,3], trans f64 [T,3]; v2 poses f64 (its jaw-zeroing ``np.concatenate`` with ``np.zeros`` promotes them,
visualizer.py:197) and trans f32 when it is kept, f64 when it is zeroed; gender str, betas f64 [300],
mocap_frame_rate f64 scalar.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence, Union

import numpy as np
import torch

LOWER_BODY_JOINTS = [1, 2, 4, 5, 7, 8, 10, 11]      # "lock below hips" (visualizer.py:345)
JAW_JOINT = 22


def feats_from_motion(poses: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    """``rearrange(poses, 'b t j d -> b t (j d)')`` + ``cat(trans)`` (trainer.py:524-527): [B,T,168]."""
    return torch.cat((poses.reshape(poses.shape[0], poses.shape[1], -1), trans), dim=-1)


def prepare_v1(feat: np.ndarray):
    """One clip [T,168|165] -> (poses [T,55,3] f32, trans [T,3] f64 zeros), visualizer.py:344-357."""
    f = np.array(feat, dtype=np.float32, copy=True).reshape(feat.shape[0], -1, 3)
    if f.shape[1] == 56:
        f = f[:, :-1, :]
    if f.shape[1] != 55:
        raise ValueError(f"expected 55 (+1 trans) joints, got {f.shape[1]}")
    f[:, LOWER_BODY_JOINTS, :] = f[0, LOWER_BODY_JOINTS, :]
    return f, np.zeros((f.shape[0], 3))


def prepare_v2(feat: np.ndarray, zero_trans=False, freeze_init_lobody=False, half_body=False):
    """One clip [T,168] -> (poses, trans), visualizer.py:192-211."""
    f = np.array(feat, dtype=np.float32, copy=True).reshape(feat.shape[0], -1, 3)
    if f.shape[1] != 56:
        raise AssertionError(f"SMPL-X data should have 56 joints, got {f.shape[1]}.")
    poses, trans = f[:, :-1, :].astype(np.float64), f[:, -1, :].copy()     # dtypes as the reference produces them (see header)
    poses[:, JAW_JOINT, :] = 0.0
    if zero_trans:
        trans = np.zeros((poses.shape[0], 3))
        if freeze_init_lobody:
            poses[:, LOWER_BODY_JOINTS, :] = poses[0, LOWER_BODY_JOINTS, :]
    elif half_body:
        trans = np.zeros((poses.shape[0], 3))
        poses[:, LOWER_BODY_JOINTS, :] = poses[0, LOWER_BODY_JOINTS, :]
    return poses, trans


def write_motion_npz(path: Union[str, Path], poses: np.ndarray, trans: np.ndarray, gender: str = "neutral",
                     betas: Optional[Sequence[float]] = None, fps: float = 30.0) -> Path:
    """``np.savez`` with the reference's keys (visualizer.py:216-222, 358-364)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    b = np.zeros(300, dtype=np.float64) if betas is None else np.asarray(betas, dtype=np.float64)
    np.savez(path, poses=np.asarray(poses), trans=np.asarray(trans),
             gender=np.array(gender), betas=b, mocap_frame_rate=np.array(fps, dtype="float64"))
    return path if path.suffix == ".npz" else path.with_suffix(path.suffix + ".npz")


def write_batch(out_dir: Union[str, Path], feats: torch.Tensor, subject: str = "scott", version: str = "v1",
                gender: str = "neutral", betas=None, fps: float = 30.0, stem: str = "seq", **v2_flags):
    """[B,T,168] (device or host) -> one ``<subject>_<stem>_<i>_motion_smplx.npz`` per clip.  One D2H copy."""
    host = feats.detach().to("cpu", non_blocking=False).numpy()
    out = []
    for i, f in enumerate(host):
        poses, trans = prepare_v1(f) if version == "v1" else prepare_v2(f, **v2_flags)
        out.append(write_motion_npz(Path(out_dir) / f"{subject}_{stem}_{i}_motion_smplx.npz", poses, trans, gender,
                                    betas, fps))
    return out
