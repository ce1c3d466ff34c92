"""CPU restatement of the audio side of the path: ``process_single_seq`` and the three AST encoders.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED against the reference itself: the
encoders are timm-0.4.5 ``vit_deit_base_distilled_patch16_384`` models (``timm==0.4.5`` hard-asserted
at models/audio/audio_main_new.py:52, pinned in amuse.yml:281) and timm is not installed in the
build container, nor does the reference hold tests or golden vectors for them.  What IS checked: the
same weights through the image's HuggingFace ``transformers.ASTModel`` -- an independent implementation
of the same published AST/DeiT encoder -- agree with this restatement to 7e-7 on |feat| = 3.3
(tests/test_oracle.py::test_ast_restatement_matches_hf_transformers; runs on the GPU box too).
The maths below restates the published DeiT/ViT block as timm 0.4.5 implements it, anchored on the
reference's own forward (models/audio/audio_main_new.py:174-204) and eval entry
(models/audio/AST_EVP.py:84-90):

    x [B,1024,128] -> unsqueeze(1).transpose(2,3) -> Conv2d(1,768,k=16,s=10) -> [B,768,12,101]
      -> flatten(2).transpose(1,2) [B,1212,768]; prepend cls, dist; + pos_embed [1,1214,768]
    12 x   x = x + proj(softmax((q k^T) * 64^-0.5) v),  q,k,v = qkv(LN_1e-6(x)) as (B,N,3,12,64)
           x = x + fc2(gelu_erf(fc1(LN_1e-6(x))))
    x = LN_1e-6(x);  frame_based_feats=True:  feature = Linear(LN_1e-5(mean(x[:, 2:], dim=1)))

State-dict keys are timm's (oracle/weights.py ``ast_state_dict``); ``depth`` is inferred.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
HEADS = 12


def _depth(sd: Dict[str, Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}.v.blocks.{n}.norm1.weight" in sd:
        n += 1
    return n


def ast_branch(sd: Dict[str, Tensor], prefix: str, fbank: Tensor) -> Tensor:
    """``ASTModel.forward(x, frame_based_feats=True)['feature']`` (audio_main_new.py:174-204)."""
    v = f"{prefix}.v"
    B = fbank.shape[0]
    x = fbank.unsqueeze(1).transpose(2, 3)                                     # [B,1,128,1024]
    x = F.conv2d(x, sd[f"{v}.patch_embed.proj.weight"], sd[f"{v}.patch_embed.proj.bias"], stride=10)
    x = x.flatten(2).transpose(1, 2)                                           # [B,1212,768]
    x = torch.cat([sd[f"{v}.cls_token"].expand(B, -1, -1), sd[f"{v}.dist_token"].expand(B, -1, -1), x], dim=1)
    x = x + sd[f"{v}.pos_embed"]
    D = x.shape[-1]
    hd = D // HEADS
    for i in range(_depth(sd, prefix)):
        b = f"{v}.blocks.{i}"
        h = F.layer_norm(x, (D,), sd[f"{b}.norm1.weight"], sd[f"{b}.norm1.bias"], 1e-6)
        qkv = F.linear(h, sd[f"{b}.attn.qkv.weight"], sd[f"{b}.attn.qkv.bias"])
        qkv = qkv.reshape(B, -1, 3, HEADS, hd).permute(2, 0, 3, 1, 4)          # timm 0.4.5 Attention.forward
        q, k, vv = qkv[0], qkv[1], qkv[2]
        att = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
        o = (att @ vv).transpose(1, 2).reshape(B, -1, D)
        x = x + F.linear(o, sd[f"{b}.attn.proj.weight"], sd[f"{b}.attn.proj.bias"])
        h = F.layer_norm(x, (D,), sd[f"{b}.norm2.weight"], sd[f"{b}.norm2.bias"], 1e-6)
        h = F.gelu(F.linear(h, sd[f"{b}.mlp.fc1.weight"], sd[f"{b}.mlp.fc1.bias"]))
        x = x + F.linear(h, sd[f"{b}.mlp.fc2.weight"], sd[f"{b}.mlp.fc2.bias"])
    x = F.layer_norm(x, (D,), sd[f"{v}.norm.weight"], sd[f"{v}.norm.bias"], 1e-6)
    feat = x[:, 2:, :].mean(dim=1)
    feat = F.layer_norm(feat, (D,), sd[f"{prefix}.feature_head.0.weight"], sd[f"{prefix}.feature_head.0.bias"], 1e-5)
    return F.linear(feat, sd[f"{prefix}.feature_head.1.weight"], sd[f"{prefix}.feature_head.1.bias"])


def ast_features(sd: Dict[str, Tensor], fbank: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """``AST_EVP.eval_func(x, frame_based_feats=True)`` (AST_EVP.py:84-90) -> (con, emo, sty), each [B,256]."""
    with torch.no_grad():
        return (ast_branch(sd, "con_enc", fbank), ast_branch(sd, "emo_enc", fbank), ast_branch(sd, "sty_enc", fbank))


def fbank_features(waveform: Tensor, target_length: int = 1024, mean: float = -9.173025, std: float = 5.062332,
                   num_mel_bins: int = 128) -> Tensor:
    """The host part of ``process_single_seq`` (infer_ldm.py:180-190): kaldi fbank, zero-pad / truncate
    to 1024 frames, THEN normalise (pad rows become +0.906)."""
    import torchaudio
    fb = torchaudio.compliance.kaldi.fbank(waveform, htk_compat=True, sample_frequency=16000, use_energy=False,
                                           window_type="hanning", num_mel_bins=num_mel_bins, dither=0.0, frame_shift=10)
    p = target_length - fb.shape[0]
    if p > 0:
        fb = F.pad(fb, (0, 0, 0, p))
    elif p < 0:
        fb = fb[:target_length]
    return (fb - mean) / (std * 2)
