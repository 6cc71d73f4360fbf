"""Seeded synthetic weights and inputs for the SDM denoising path.

Everything is generated with ``numpy.random.RandomState`` (bit-stable across
numpy/torch versions and machines), so the golden fixtures under
``tests/golden/`` only need to store *outputs*: the inputs and the 229-entry
``model_state_dict`` are regenerated from the seed wherever the tests run
(``/root/reference`` does not exist on the GPU box).

Shapes follow the reference's state-dict contract (SURVEY.md Appendix B;
reference ``model/sdm.py:19-129``, ``model/pcd_backbone/pointnet2.py:43-59``,
``posa/posa_models.py:292-317``, ``model/diffusion_utils.py:7-62,91-103``) and the
dataset tensor contract (reference ``posa/dataset.py:445-474``).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

N_POINTS = 1024
N_OBJ = 9
CLIP_DIM = 512
LATENT = 128
CAT_EMB = 32
TRANS = 12

# (name, npoint, radius, nsample, in_channel, mlp) -- reference pointnet2.py:46-49
SA_SPECS = (
    ("sa1", 1024, 0.1, 32, 6, (32, 32, 64)),
    ("sa2", 256, 0.2, 32, 67, (64, 64, 128)),
    ("sa3", 64, 0.4, 32, 131, (128, 128, 256)),
    ("sa4", 16, 0.8, 32, 259, (256, 256, 512)),
)
# (name, in_channel, mlp) -- reference pointnet2.py:50-55
FP_SPECS = (
    ("fp4", 768, (256, 256)),
    ("fp3", 384, (256, 256)),
    ("fp2", 320, (256, 128)),
    ("fp1", 128, (128, 128, 128)),
)


def state_dict_spec(max_cats: int = 13):
    """Ordered list of ``(key, shape, kind)`` for every non-CLIP state-dict entry.

    kind: 'w' linear/conv weight (fan_in = prod(shape[1:])), 'b' bias,
    'bn_w','bn_b','bn_rm','bn_rv','bn_nbt', 'gn_w','gn_b', 'pe' (computed buffer).
    """
    s = []

    def lin(prefix, out_f, in_f):
        s.append((prefix + ".weight", (out_f, in_f), "w"))
        s.append((prefix + ".bias", (out_f,), "b"))

    def bn(prefix, c):
        s.append((prefix + ".weight", (c,), "bn_w"))
        s.append((prefix + ".bias", (c,), "bn_b"))
        s.append((prefix + ".running_mean", (c,), "bn_rm"))
        s.append((prefix + ".running_var", (c,), "bn_rv"))
        s.append((prefix + ".num_batches_tracked", (), "bn_nbt"))

    s.append(("sequence_pos_encoder.pe", (5000, 1, LATENT), "pe"))
    s.append(("embed_timestep.sequence_pos_encoder.pe", (5000, 1, LATENT), "pe"))
    lin("embed_timestep.time_embed.0", LATENT, LATENT)
    lin("embed_timestep.time_embed.2", LATENT, LATENT)
    lin("embed_text.0", CLIP_DIM // 2, CLIP_DIM)
    lin("embed_text.2", LATENT * 2, CLIP_DIM // 2)
    lin("embed_text.4", LATENT, LATENT * 2)
    lin("embed_cat.0", CAT_EMB, max_cats)
    lin("predict_cat.0", LATENT // 2, LATENT)
    lin("predict_cat.2", LATENT // 4, LATENT // 2)
    lin("predict_cat.4", max_cats, LATENT // 4)
    # nn.MultiheadAttention(E=128, H=8, kdim=32, vdim=3072)
    s.append(("attn_layer.q_proj_weight", (LATENT, LATENT), "w"))
    s.append(("attn_layer.k_proj_weight", (LATENT, CAT_EMB), "w"))
    s.append(("attn_layer.v_proj_weight", (LATENT, N_POINTS * 3), "w"))
    s.append(("attn_layer.in_proj_bias", (3 * LATENT,), "b"))
    lin("attn_layer.out_proj", LATENT, LATENT)
    lin("translation_layer.0", LATENT, LATENT + CAT_EMB)
    lin("translation_layer.2", TRANS, LATENT)
    lin("point_wise_trans_layer.0", 3, TRANS + 3)
    # nn.MultiheadAttention(E=12, H=12, kdim=vdim=3)
    s.append(("pcd_attention.q_proj_weight", (TRANS, TRANS), "w"))
    s.append(("pcd_attention.k_proj_weight", (TRANS, 3), "w"))
    s.append(("pcd_attention.v_proj_weight", (TRANS, 3), "w"))
    s.append(("pcd_attention.in_proj_bias", (3 * TRANS,), "b"))
    lin("pcd_attention.out_proj", TRANS, TRANS)
    for name, _np, _r, _ns, cin, mlp in SA_SPECS:
        last = cin
        for i, co in enumerate(mlp):
            s.append((f"pcd_backbone.{name}.mlp_convs.{i}.weight", (co, last, 1, 1), "w"))
            s.append((f"pcd_backbone.{name}.mlp_convs.{i}.bias", (co,), "b"))
            last = co
        for i, co in enumerate(mlp):
            bn(f"pcd_backbone.{name}.mlp_bns.{i}", co)
    for name, cin, mlp in FP_SPECS:
        last = cin
        for i, co in enumerate(mlp):
            s.append((f"pcd_backbone.{name}.mlp_convs.{i}.weight", (co, last, 1), "w"))
            s.append((f"pcd_backbone.{name}.mlp_convs.{i}.bias", (co,), "b"))
            last = co
        for i, co in enumerate(mlp):
            bn(f"pcd_backbone.{name}.mlp_bns.{i}", co)
    s.append(("pcd_backbone.conv1.weight", (128, 128, 1), "w"))
    s.append(("pcd_backbone.conv1.bias", (128,), "b"))
    bn("pcd_backbone.bn1", 128)
    s.append(("pcd_backbone.conv2.weight", (3, 128, 1), "w"))
    s.append(("pcd_backbone.conv2.bias", (3,), "b"))
    for i, (co, ci) in enumerate(((64, 3), (64, 64), (64, 64))):
        lin(f"human_backbone.de_spiral.{i}.conv.layer", co, ci)
        s.append((f"human_backbone.de_spiral.{i}.norm.weight", (co,), "gn_w"))
        s.append((f"human_backbone.de_spiral.{i}.norm.bias", (co,), "gn_b"))
    lin("human_backbone.de_spiral.3.layer", 3, 64)
    lin("upsampling_layer.0", 128, 1)
    lin("upsampling_layer.2", 512, 128)
    lin("upsampling_layer.4", N_POINTS, 512)
    lin("combine_extraction.0", LATENT, 2 * LATENT)
    lin("input_process.pose_embedding.0", LATENT // 2, 3)
    lin("input_process.pose_embedding.2", LATENT, LATENT // 2)
    lin("input_process.combination_extraction.0", 192, 2 * LATENT)
    lin("input_process.combination_extraction.2", LATENT, 192)
    lin("output_process.pose_final.0", LATENT // 2, LATENT)
    lin("output_process.pose_final.2", 3, LATENT // 2)
    return s


def positional_table(max_len: int = 5000, d_model: int = LATENT) -> torch.Tensor:
    """Sinusoidal table ``pe[max_len,1,d]`` (reference model/diffusion_utils.py:29-35)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def make_state_dict(seed: int = 0, kind: str = "wellcond", max_cats: int = 13):
    """Seeded non-CLIP ``model_state_dict``.

    kind='default': U(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights and biases (the
    statistics of torch's default Linear/Conv init), identity norm layers.
    Under this init the PointNet++ output is near-constant, so
    parity on it cannot see backbone bugs.
    kind='wellcond': He-normal weights, small normal biases, randomised
    BatchNorm running stats / affine and GroupNorm affine, so that every stage
    has O(1) spread and errors anywhere are visible at the output.
    """
    assert kind in ("default", "wellcond")
    rs = np.random.RandomState(seed)
    pe = positional_table()
    sd = OrderedDict()
    for key, shape, k in state_dict_spec(max_cats):
        if k == "pe":
            sd[key] = pe.clone()
            continue
        if k == "bn_nbt":
            sd[key] = torch.tensor(0 if kind == "default" else 100, dtype=torch.int64)
            continue
        n = int(np.prod(shape)) if len(shape) else 1
        if k == "w":
            fan_in = int(np.prod(shape[1:]))
            if kind == "default":
                bound = 1.0 / math.sqrt(fan_in)
                a = rs.uniform(-bound, bound, size=n)
            else:
                a = rs.standard_normal(n) * math.sqrt(2.0 / fan_in)
        elif k == "b":
            if kind == "default":
                a = rs.uniform(-0.05, 0.05, size=n)
            else:
                a = rs.standard_normal(n) * 0.1
        elif k in ("bn_w", "gn_w"):
            a = np.ones(n) if kind == "default" else rs.uniform(0.6, 1.4, size=n)
        elif k in ("bn_b", "gn_b"):
            a = np.zeros(n) if kind == "default" else rs.standard_normal(n) * 0.1
        elif k == "bn_rm":
            a = np.zeros(n) if kind == "default" else rs.standard_normal(n) * 0.1
        elif k == "bn_rv":
            a = np.ones(n) if kind == "default" else rs.uniform(0.5, 1.5, size=n)
        else:  # pragma: no cover
            raise AssertionError(k)
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape))
    return sd


def make_inputs(seed: int, batch: int, max_cats: int = 13, training: bool = False):
    """Seeded conditions in the dataset's tensor contract.

    Returns a dict of CPU tensors: ``mask[B,9]`` (slot 0 always 0 as in the
    dataset, others Bernoulli(.5)), ``given_objs[B,9,1024,3]`` (present objects
    U(-.5,.5)^3, absent all-zero, human slot 0 always filled), ``given_cats[B,9,C]``
    one-hot (absent rows zero), ``text_emb[B,512]`` standing in for CLIP(y).float(),
    ``x_T[B,1024,3]``; with ``training=True`` also ``x_start``, ``target_cat`` and ``t``.
    """
    rs = np.random.RandomState(seed)
    B = batch
    present = rs.rand(B, N_OBJ) < 0.5
    present[:, 0] = True
    mask = present.astype(np.float32)
    mask[:, 0] = 0.0
    objs = rs.uniform(-0.5, 0.5, size=(B, N_OBJ, N_POINTS, 3)).astype(np.float32)
    objs *= present[:, :, None, None]
    cats_idx = rs.randint(1, max_cats, size=(B, N_OBJ))
    cats_idx[:, 0] = 0
    cats = np.zeros((B, N_OBJ, max_cats), np.float32)
    bi, oi = np.nonzero(present)
    cats[bi, oi, cats_idx[bi, oi]] = 1.0
    out = {
        "mask": torch.from_numpy(mask),
        "given_objs": torch.from_numpy(objs),
        "given_cats": torch.from_numpy(cats),
        "text_emb": torch.from_numpy(rs.standard_normal((B, CLIP_DIM)).astype(np.float32)),
        "x_T": torch.from_numpy(rs.standard_normal((B, N_POINTS, 3)).astype(np.float32)),
    }
    if training:
        out["x_start"] = torch.from_numpy(rs.uniform(-0.5, 0.5, size=(B, N_POINTS, 3)).astype(np.float32))
        tc = np.zeros((B, max_cats), np.float32)
        tc[np.arange(B), rs.randint(0, max_cats, size=B)] = 1.0
        out["target_cat"] = torch.from_numpy(tc)
        out["t"] = torch.from_numpy(rs.randint(0, 1000, size=B).astype(np.int64))
    return out


FPS_LEVEL_N = (1024, 1024, 256, 64)  # randint upper bounds, reference pointnet2_utils.py:72


def make_step_randoms(seed: int, batch: int, steps: int = 1):
    """Per-step randoms in the order the reference consumes them (SURVEY.md trap 4):
    four FPS start vectors ``randint(0,N,(9B,))`` with N = 1024,1024,256,64, then the
    sampling noise ``randn(B,1024,3)``.  Returns ``fps_start[steps,4,9B]`` int64 and
    ``noise[steps,B,1024,3]`` float32.
    """
    rs = np.random.RandomState(seed)
    C = batch * N_OBJ
    fps = np.stack(
        [np.stack([rs.randint(0, n, size=C) for n in FPS_LEVEL_N]) for _ in range(steps)]
    ).astype(np.int64)
    noise = rs.standard_normal((steps, batch, N_POINTS, 3)).astype(np.float32)
    return torch.from_numpy(fps), torch.from_numpy(noise)


def make_dropout_mask(seed: int, batch: int):
    """Dropout(0.5) mask of the backbone head (reference pointnet2.py:76) in the reference's layout ``[9B,128,1024]``:
    0 where dropped, 2 (= 1/(1-p)) where kept."""
    rs = np.random.RandomState(seed)
    keep = rs.rand(batch * N_OBJ, 128, N_POINTS) < 0.5
    return torch.from_numpy(keep.astype(np.float32) * 2.0)


# ----------------------------------------------------------------------------------------------------------------------
# CLIP text tower (model/sdm.py:245-277): seeded random weights under the openai/CLIP state-dict names and token batches in
# clip.tokenize's layout (SOT, word ids, EOT = largest id, zero padding).  No checkpoint or BPE vocabulary exists offline.
# ----------------------------------------------------------------------------------------------------------------------
def make_clip_state_dict(seed=0, width=512, layers=12, vocab=49408, ctx=77, embed=512, prefix=""):
    """Initialisation statistics of clip/model.py ``CLIP.initialize_parameters`` (+ randomised LayerNorm affine so that
    every tensor matters); fp32 torch tensors keyed ``<prefix>token_embedding.weight`` ..."""
    r = np.random.RandomState(seed)

    def t(shape, std, mean=0.0):
        return torch.from_numpy((r.standard_normal(shape) * std + mean).astype(np.float32))

    proj_std, attn_std, fc_std = (width ** -0.5) * ((2 * layers) ** -0.5), width ** -0.5, (2 * width) ** -0.5
    sd = {"token_embedding.weight": t((vocab, width), 0.02), "positional_embedding": t((ctx, width), 0.01),
          "ln_final.weight": t((width,), 0.1, 1.0), "ln_final.bias": t((width,), 0.05),
          "text_projection": t((width, embed), width ** -0.5)}
    for l in range(layers):
        p = f"transformer.resblocks.{l}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = t((width,), 0.1, 1.0), t((width,), 0.05)
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = t((width,), 0.1, 1.0), t((width,), 0.05)
        sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = t((3 * width, width), attn_std), t((3 * width,), 0.02)
        sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"] = t((width, width), proj_std), t((width,), 0.02)
        sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"] = t((4 * width, width), fc_std), t((4 * width,), 0.02)
        sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"] = t((width, 4 * width), proj_std), t((width,), 0.02)
    return {prefix + k: v for k, v in sd.items()}


def make_clip_tokens(seed, batch, vocab=49408, ctx=77, max_words=20):
    """int32 [batch, ctx]: SOT (vocab-2), 0..max_words word ids, EOT (vocab-1), zeros -- what ``_encode_text_clip`` builds
    (model/sdm.py:248-256: tokenised to max_words+2 positions, zero-padded to the 77-token context)."""
    r = np.random.RandomState(seed)
    tok = np.zeros((batch, ctx), dtype=np.int32)
    for b in range(batch):
        n = int(r.randint(0, max_words + 1)) if b else max_words  # sample 0 has the longest sentence
        tok[b, 0] = vocab - 2
        tok[b, 1:1 + n] = r.randint(1, vocab - 2, size=n)
        tok[b, 1 + n] = vocab - 1
    return torch.from_numpy(tok)
