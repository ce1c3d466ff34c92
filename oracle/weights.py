"""Deterministic synthetic weights with the reference's state-dict key names.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference ships no checkpoints (``saved-models/`` is not in the repo, SURVEY.md
section 8c), so parity is defined on seeded synthetic weights.  Key names and shapes
follow SURVEY.md App. A.4 (probe-printed from the reference modules):

* Denoiser      -- models/latent_diffusion/denoiser.py:66-105 (inside an LDM checkpoint
                   the keys carry the prefix ``denoiser.``, infer_ldm.py:91-104)
* MotionPrior   -- models/latent_diffusion/vae.py:85-146
* AST_EVP       -- models/audio/AST_EVP.py:53-61 + audio_main_new.py:63-90 (timm DeiT keys)

The value distributions are ours (not the reference's initialisers): they are chosen
so that attention logits, GELU inputs and the 6D rotation features are O(1), i.e. a
wrong scale / wrong head split / wrong LN epsilon shows up in the parity tests.
Everything is drawn from ONE ``torch.Generator`` in a fixed key order, so the same
seed gives the same tensors in the build container and on the GPU box (same image);
``checksum`` lets the golden fixtures verify that.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import torch

D = 128          # latent_dim[-1]               configs/diff_latent_v2.json:25-28
FF = 512         # ff_size                      configs/diff_latent_v2.json:29
COND = 256       # cond_dim                     configs/diff_latent_v2.json:39
NFEATS = 333     # 201 + 132 (6D)               vae.py:66-69
BLOCKS = (["input_blocks.%d" % i for i in range(4)] + ["middle_block"]
          + ["output_blocks.%d" % i for i in range(4)])


class _Drawer:
    def __init__(self, seed: int):
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(seed)
        self.sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def normal(self, name, shape, std, mean=0.0):
        self.sd[name] = torch.randn(shape, generator=self.g, dtype=torch.float32) * std + mean

    def uniform(self, name, shape, lo=0.0, hi=1.0):
        self.sd[name] = torch.rand(shape, generator=self.g, dtype=torch.float32) * (hi - lo) + lo

    def linear(self, prefix, out_f, in_f, gain=1.0, wname="weight", bname="bias"):
        self.normal(f"{prefix}.{wname}" if wname else prefix, (out_f, in_f), gain / in_f ** 0.5)
        self.normal(f"{prefix}.{bname}", (out_f,), 0.05)

    def layernorm(self, prefix, dim):
        self.normal(f"{prefix}.weight", (dim,), 0.1, mean=1.0)
        self.normal(f"{prefix}.bias", (dim,), 0.05)

    def mha(self, prefix, dim, gain):
        self.normal(f"{prefix}.in_proj_weight", (3 * dim, dim), gain / dim ** 0.5)
        self.normal(f"{prefix}.in_proj_bias", (3 * dim,), 0.05)
        self.linear(f"{prefix}.out_proj", dim, dim)


def denoiser_state_dict(seed: int = 2024, attn_gain: float = 1.2) -> "OrderedDict[str, torch.Tensor]":
    """Keys of ``Denoiser.state_dict()`` (denoiser.py:66-105, cross_attention.py:18-36,236-257)."""
    d = _Drawer(seed)
    d.linear("time_embedding.linear_1", D, COND)
    d.linear("time_embedding.linear_2", D, D)
    for c in ("con", "emo", "sty"):
        d.linear(f"emb_proj_{c}.1", D, COND)
    d.uniform("query_pos.pe", (500, 1, D))
    d.uniform("mem_pos.pe", (500, 1, D))      # present in the module, unused by trans_enc
    d.layernorm("encoder.norm", D)
    for b in BLOCKS:
        p = f"encoder.{b}"
        d.mha(f"{p}.self_attn", D, attn_gain)
        d.linear(f"{p}.linear1", FF, D)
        d.linear(f"{p}.linear2", D, FF)
        d.layernorm(f"{p}.norm1", D)
        d.layernorm(f"{p}.norm2", D)
    for i in range(4):
        d.linear(f"encoder.linear_blocks.{i}", D, 2 * D)
    return d.sd


def motionprior_state_dict(seed: int = 2025, attn_gain: float = 1.2, with_encoder: bool = True):
    """Keys of ``MotionPrior.state_dict()`` after ``setup`` (vae.py:85-146)."""
    d = _Drawer(seed)
    d.normal("global_motion_token", (2, D), 1.0)
    d.uniform("query_pos_encoder.pe", (500, 1, D))
    d.uniform("query_pos_decoder.pe", (500, 1, D))
    if with_encoder:
        d.layernorm("encoder.norm", D)
        for b in BLOCKS:
            p = f"encoder.{b}"
            d.mha(f"{p}.self_attn", D, attn_gain)
            d.linear(f"{p}.linear1", FF, D)
            d.linear(f"{p}.linear2", D, FF)
            d.layernorm(f"{p}.norm1", D)
            d.layernorm(f"{p}.norm2", D)
        for i in range(4):
            d.linear(f"encoder.linear_blocks.{i}", D, 2 * D)
    d.layernorm("decoder.norm", D)
    for b in BLOCKS:
        p = f"decoder.{b}"
        d.mha(f"{p}.self_attn", D, attn_gain)
        d.mha(f"{p}.multihead_attn", D, attn_gain)
        d.linear(f"{p}.linear1", FF, D)
        d.linear(f"{p}.linear2", D, FF)
        d.layernorm(f"{p}.norm1", D)
        d.layernorm(f"{p}.norm2", D)
        d.layernorm(f"{p}.norm3", D)
    for i in range(4):
        d.linear(f"decoder.linear_blocks.{i}", D, 2 * D)
    d.linear("skel_embedding", D, NFEATS)
    d.linear("final_layer", NFEATS, D)
    return d.sd


def ast_branch_state_dict(d: _Drawer, prefix: str, depth: int = 12, dim: int = 768,
                          n_tokens: int = 1214, feat: int = 256):
    """One ``ASTModel`` branch (audio_main_new.py:63-90; timm-0.4.5 DeiT key names)."""
    v = f"{prefix}.v"
    d.normal(f"{v}.cls_token", (1, 1, dim), 0.02)
    d.normal(f"{v}.dist_token", (1, 1, dim), 0.02)
    d.normal(f"{v}.pos_embed", (1, n_tokens, dim), 0.02)
    d.normal(f"{v}.patch_embed.proj.weight", (dim, 1, 16, 16), 1.0 / 16.0)
    d.normal(f"{v}.patch_embed.proj.bias", (dim,), 0.05)
    for i in range(depth):
        b = f"{v}.blocks.{i}"
        d.layernorm(f"{b}.norm1", dim)
        d.linear(f"{b}.attn.qkv", 3 * dim, dim, gain=1.1)
        d.linear(f"{b}.attn.proj", dim, dim, gain=0.5)
        d.layernorm(f"{b}.norm2", dim)
        d.linear(f"{b}.mlp.fc1", 4 * dim, dim)
        d.linear(f"{b}.mlp.fc2", dim, 4 * dim, gain=0.5)
    d.layernorm(f"{v}.norm", dim)
    d.layernorm(f"{prefix}.feature_head.0", dim)
    d.linear(f"{prefix}.feature_head.1", feat, dim)


def ast_state_dict(seed: int = 2026, depth: int = 12, dim: int = 768, n_tokens: int = 1214):
    """Encoder-side keys of ``AST_EVP.state_dict()`` (AST_EVP.py:53-61).  ``depth``/``dim``
    can be reduced for fast CPU tests; the shipped configuration is 12 / 768 / 1214."""
    d = _Drawer(seed)
    for br in ("emo_enc", "sty_enc", "con_enc"):
        ast_branch_state_dict(d, br, depth=depth, dim=dim, n_tokens=n_tokens)
    return d.sd


def checksum(sd) -> str:
    """Order-sensitive SHA-1 over the raw fp32 bytes -- stored in the golden fixtures."""
    h = hashlib.sha1()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()
