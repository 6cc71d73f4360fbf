"""CPU restatement of the CLIP text tower -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Reference call site: ``model/sdm.py:245-259`` (``clip_model.encode_text(texts).float()``) with ``clip_model`` from
``clip.load('ViT-B/32')`` (``model/sdm.py:266-277``).  The ``clip`` package (openai/CLIP, an unpinned git dependency:
reference ``README.md:27``) is NOT under ``/root/reference`` and is absent from this image, so its published algorithm
(``clip/model.py``: ``CLIP.encode_text``, ``ResidualAttentionBlock``, ``QuickGELU``, ``build_attention_mask``) is restated here.
Pinning: no reference test or golden vector exists for it; ``tests/test_clip_text.py`` pins this restatement against an
independent implementation of the same published model that IS in the image, ``transformers.CLIPTextModelWithProjection``
(random weights mapped key by key) -> "pinned against transformers 5.5, unpinned against the clip package itself".
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def n_layers(sd):
    n = 0
    while f"transformer.resblocks.{n}.ln_1.weight" in sd:
        n += 1
    return n


def encode_text(sd, tokens, trace=None):
    """clip/model.py CLIP.encode_text.  ``sd``: openai-named fp32 tensors (no ``clip_model.`` prefix); tokens int [B,ctx]."""
    tokens = tokens.long()
    W = sd["ln_final.weight"].shape[0]
    H = W // 64
    B, L = tokens.shape
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"][:L]          # [B,L,W]
    mask = torch.full((L, L), float("-inf")).triu_(1)                                    # build_attention_mask
    for l in range(n_layers(sd)):
        p = f"transformer.resblocks.{l}."
        h = F.layer_norm(x, (W,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = h @ sd[p + "attn.in_proj_weight"].T + sd[p + "attn.in_proj_bias"]
        q, k, v = (t.view(B, L, H, 64).transpose(1, 2) for t in qkv.split(W, dim=-1))   # [B,H,L,64]
        a = torch.softmax((q * 0.125) @ k.transpose(-1, -2) + mask, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, L, W)
        x = x + a @ sd[p + "attn.out_proj.weight"].T + sd[p + "attn.out_proj.bias"]
        h = F.layer_norm(x, (W,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = h @ sd[p + "mlp.c_fc.weight"].T + sd[p + "mlp.c_fc.bias"]
        h = h * torch.sigmoid(1.702 * h)                                                  # QuickGELU
        x = x + h @ sd[p + "mlp.c_proj.weight"].T + sd[p + "mlp.c_proj.bias"]
        if trace is not None:
            trace[f"block{l}"] = x
    x = F.layer_norm(x, (W,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(B), tokens.argmax(dim=-1)] @ sd["text_projection"]


def from_hf(hf_state_dict):
    """Key-by-key map of a transformers ``CLIPTextModelWithProjection`` state dict onto the openai/CLIP names."""
    g = {k: v.detach().float() for k, v in hf_state_dict.items()}
    sd = {"token_embedding.weight": g["text_model.embeddings.token_embedding.weight"],
          "positional_embedding": g["text_model.embeddings.position_embedding.weight"],
          "ln_final.weight": g["text_model.final_layer_norm.weight"], "ln_final.bias": g["text_model.final_layer_norm.bias"],
          "text_projection": g["text_projection.weight"].T.contiguous()}
    l = 0
    while f"text_model.encoder.layers.{l}.layer_norm1.weight" in g:
        s, d = f"text_model.encoder.layers.{l}.", f"transformer.resblocks.{l}."
        sd[d + "attn.in_proj_weight"] = torch.cat([g[s + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        sd[d + "attn.in_proj_bias"] = torch.cat([g[s + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, b in (("self_attn.out_proj", "attn.out_proj"), ("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"), ("mlp.fc1", "mlp.c_fc"),
                     ("mlp.fc2", "mlp.c_proj")):
            sd[d + b + ".weight"], sd[d + b + ".bias"] = g[s + a + ".weight"], g[s + a + ".bias"]
        l += 1
    return sd
