"""CPU restatement of the ContactFormer attention layer -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Follows ``contact_former/transformer.py:72-103`` (MultiHeadAttention.forward) and ``:167-177`` (PositionwiseFeedForward.forward)
in eval mode.  Pinned against the live reference classes (pure torch, importable in the build container):
``tests/golden/make_golden_cf.py`` -> ``tests/golden/cf_layer.npz``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def mha(sd, x, mask=None, prefix=""):
    """sd: w_q/w_k/w_v/fc/layer_norm .weight/.bias; x [bs,S,V,64]; mask [bs,S,S] (entries == 0 are masked) or None."""
    bs, S, V, D = x.shape
    H = sd[prefix + "w_q.weight"].shape[0] // 64

    def proj(name):  # [bs,S,V,H,64] -> [bs,V,H,S,64]
        return F.linear(x, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"]).view(bs, S, V, H, 64).permute(0, 2, 3, 1, 4)

    q, k, v = proj("w_q"), proj("w_k"), proj("w_v")
    attn = (q @ k.transpose(-1, -2)) / math.sqrt(64.0)                       # [bs,V,H,S,S]
    if mask is not None:
        attn = attn.masked_fill((mask == 0)[:, None, None], float("-inf"))
    if mask is not None and mask.sum() == 0:
        attn = torch.zeros_like(attn)                                        # transformer.py:91-92
    else:
        attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).permute(0, 3, 1, 2, 4).reshape(bs, S, V, H * 64)         # heads concatenated per (b, s, v)
    out = F.linear(out, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])
    return F.layer_norm(out + x, (D,), sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"], 1e-5)


def ffn(sd, x, prefix=""):
    D = x.shape[-1]
    h = F.relu(F.linear(x, sd[prefix + "w_1.weight"].squeeze(-1), sd[prefix + "w_1.bias"]))
    y = F.linear(h, sd[prefix + "w_2.weight"].squeeze(-1), sd[prefix + "w_2.bias"])
    return F.layer_norm(y + x, (D,), sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"], 1e-5)


def encoder_layer(sd, x, mask=None):
    return ffn(sd, mha(sd, x, mask, "self_attn."), "pos_wise_ffnn.")
