"""CPU oracle for the LSDM multi-conditional denoising path.  TEST INFRASTRUCTURE ONLY.

This file is an independent, functional (state-dict in, tensors out) CPU
restatement of the reference algorithm.  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``lsdm_b200/`` imports
it, and the product path raises if the CUDA library is missing -- there is no CPU
fallback.

Parity pinning: the reference has no tests, golden vectors or fixtures of its
own, so this oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, run unmodified in the build container via ``tests/golden/ref_harness.py``;
``tests/golden/make_golden.py`` (committed) wrote the fixtures in
``tests/golden/*.npz`` and ``tests/test_oracle_vs_golden.py`` checks this file
against them on every CPU test run.

Third-party arithmetic not under /root/reference (README pins, no lock file):
``pytorch3d.loss.chamfer_distance`` (pytorch3d 0.7.3) -- restated from its published
definition in :func:`chamfer_distance` ("parity unpinned" for that one function:
the package is absent); ``clip`` -- out of scope, inputs are
``[B,512]`` embeddings.

Each function cites the reference lines it follows (paths relative to
/root/reference).  ``dtype`` may be float64 to obtain a high-precision truth.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

N_HEAD = 8
N_OBJ = 9
N_PTS = 1024
BN_EPS = 1e-5
GN_EPS = 1e-5

SA_SPECS = (("sa1", 1024, 0.1, 32), ("sa2", 256, 0.2, 32), ("sa3", 64, 0.4, 32), ("sa4", 16, 0.8, 32))
FP_SPECS = (("fp4", 2), ("fp3", 2), ("fp2", 2), ("fp1", 3))


# --------------------------------------------------------------------------------------
# schedule tables (float64 host math)
# --------------------------------------------------------------------------------------
def cosine_betas(T: int, max_beta: float = 0.999) -> np.ndarray:
    """diffusion/gaussian_diffusion.py:40-66 (cosine schedule, beta capped at 0.999)."""
    ab = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - ab((i + 1) / T) / ab(i / T), max_beta) for i in range(T)], dtype=np.float64)


def linear_betas(T: int, scale_betas: float = 1.0) -> np.ndarray:
    """diffusion/gaussian_diffusion.py:31-39."""
    scale = scale_betas * 1000 / T
    return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)


def space_timesteps(T: int, section_counts) -> list:
    """diffusion/respace.py:8-61: kept timesteps for 'ddimN' or a list of section counts."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, T):
                if len(range(0, T, stride)) == want:
                    return sorted(range(0, T, stride))
            raise ValueError("no integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    base, extra = divmod(T, len(section_counts))
    start, keep = 0, []
    for i, cnt in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < cnt:
            raise ValueError("section too small")
        stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            keep.append(start + round(cur))
            cur += stride
        start += size
    return sorted(set(keep))


def respaced_betas(betas: np.ndarray, keep) -> np.ndarray:
    """diffusion/respace.py:78-87: beta'_k = 1 - abar_i / abar_prev-kept."""
    keep = set(keep)
    abar = np.cumprod(1.0 - betas)
    last, out = 1.0, []
    for i, a in enumerate(abar):
        if i in keep:
            out.append(1 - a / last)
            last = a
    return np.array(out, dtype=np.float64)


def diffusion_tables(betas: np.ndarray) -> dict:
    """diffusion/gaussian_diffusion.py:166-202 (all float64)."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    abar = np.cumprod(alphas)
    abar_prev = np.append(1.0, abar[:-1])
    post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
    return {
        "betas": betas,
        "alphas_cumprod": abar,
        "sqrt_alphas_cumprod": np.sqrt(abar),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - abar),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(abar_prev) / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * np.sqrt(alphas) / (1.0 - abar),
    }


def _extract(arr: np.ndarray, t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """diffusion/gaussian_diffusion.py:1585-1598: gather in float64, THEN cast to fp32."""
    v = torch.from_numpy(arr).to(t.device)[t].float().to(like.dtype)
    return v.view(-1, *([1] * (like.dim() - 1)))


# --------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------
def _lin(sd, key, x):
    return x @ sd[key + ".weight"].T + sd[key + ".bias"]


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _cast(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def timestep_embedding(sd, t):
    """model/diffusion_utils.py:20-21: time_embed(pe[t]) -> [B,128]."""
    pe = sd["embed_timestep.sequence_pos_encoder.pe"][t, 0]
    h = _lin(sd, "embed_timestep.time_embed.0", pe)
    h = h * torch.sigmoid(h)
    return _lin(sd, "embed_timestep.time_embed.2", h)


def text_embedding(sd, text):
    """model/sdm.py:52-59,152: 512->256->256->128, GELU after each."""
    h = text
    for i in (0, 2, 4):
        h = _gelu(_lin(sd, f"embed_text.{i}", h))
    return h


def predict_category(sd, enc):
    """model/sdm.py:68-76,157: 128->64->32->C, GELU after each, softmax over classes."""
    h = enc
    for i in (0, 2, 4):
        h = _gelu(_lin(sd, f"predict_cat.{i}", h))
    return torch.softmax(h, dim=-1)


def object_attention_weights(sd, enc, emb_cat, mask_global, b_offset):
    """model/sdm.py:79,180-182: 8-head 1x9 attention, weights only, head-averaged.

    The additive float mask is laid out ``[h*B+b]`` by ``repeat`` but consumed as ``[b*H+h]``
    by torch MHA, so head h of (global) sample b sees ``mask[(b*H+h) mod B]`` (SURVEY trap 2).
    """
    B = enc.shape[0]
    Bg = mask_global.shape[0]
    E, H = 128, N_HEAD
    d = E // H
    bias = sd["attn_layer.in_proj_bias"]
    q = enc @ sd["attn_layer.q_proj_weight"].T + bias[:E]  # [B,128]
    k = emb_cat @ sd["attn_layer.k_proj_weight"].T + bias[E : 2 * E]  # [B,9,128]
    q = q.view(B, H, d)
    k = k.view(B, N_OBJ, H, d)
    logits = torch.einsum("bhd,bohd->bho", q, k) / math.sqrt(d)
    bg = torch.arange(B) + b_offset
    rows = (bg[:, None] * H + torch.arange(H)[None, :]) % Bg  # [B,H]
    logits = logits + mask_global[rows]  # [B,H,9]
    return torch.softmax(logits, dim=-1).mean(dim=1)  # [B,9]


def translation_params(sd, enc, emb_cat):
    """model/sdm.py:81-87,185-186: 160->128->12 with GELU after each, per (b,o)."""
    z = torch.cat([emb_cat, enc[:, None, :].expand(-1, N_OBJ, -1)], dim=-1)
    h = _gelu(_lin(sd, "translation_layer.0", z))
    return _gelu(_lin(sd, "translation_layer.2", h))  # [B,9,12]


def upsample_embedding(sd, ts, enc):
    """model/sdm.py:108-122,164-167,208: per-scalar 1->128->512->1024 GELU MLP over the 256 scalars
    of [ts || enc], transposed to [B,1024,256], then Linear 256->128 + GELU."""
    s = torch.cat([ts, enc], dim=-1)[..., None]  # [B,256,1]
    h = s
    for i in (0, 2, 4):
        h = _gelu(_lin(sd, f"upsampling_layer.{i}", h))
    u = h.permute(0, 2, 1)  # [B,1024,256]
    return _gelu(_lin(sd, "combine_extraction.0", u))  # [B,1024,128]


def human_decoder(sd, pts):
    """posa/posa_models.py:152-160,181-187,320-326 with seq_length=1 (spiral = the vertex itself):
    per-point 3->64 GN ReLU, 64->64 GN ReLU over 1024 pts; first 655 pts; 64->64 GN ReLU; 64->3;
    nearest x2 upsample along points, keep the first 1024."""

    def gn(h, i):
        w, b = sd[f"human_backbone.de_spiral.{i}.norm.weight"], sd[f"human_backbone.de_spiral.{i}.norm.bias"]
        return F.group_norm(h.permute(0, 2, 1), 8, w, b, GN_EPS).permute(0, 2, 1)

    h = torch.relu(gn(_lin(sd, "human_backbone.de_spiral.0.conv.layer", pts), 0))
    h = torch.relu(gn(_lin(sd, "human_backbone.de_spiral.1.conv.layer", h), 1))
    h = h[:, :655]
    h = torch.relu(gn(_lin(sd, "human_backbone.de_spiral.2.conv.layer", h), 2))
    h = _lin(sd, "human_backbone.de_spiral.3.layer", h)  # [B,655,3]
    idx = torch.arange(N_PTS) // 2
    return h[:, idx]


# --------------------------------------------------------------------------------------
# PointNet++ (model/pcd_backbone/pointnet2.py:61-80, pointnet2_utils.py)
# --------------------------------------------------------------------------------------
def square_distance(src, dst):
    """pointnet2_utils.py:19-38: -2ab, then += |a|^2, then += |b|^2 (this order)."""
    d = -2 * torch.matmul(src, dst.transpose(1, 2))
    d += torch.sum(src**2, -1)[:, :, None]
    d += torch.sum(dst**2, -1)[:, None, :]
    return d


def farthest_point_sample(xyz, npoint, start):
    """pointnet2_utils.py:60-81 with the ``randint`` start injected.  dist uses the direct
    difference form, running min initialised to 1e10, argmax takes the first maximum."""
    C, N, _ = xyz.shape
    idx = torch.zeros(C, npoint, dtype=torch.long)
    dist = torch.full((C, N), 1e10, dtype=xyz.dtype)
    far = start.clone()
    ar = torch.arange(C)
    for i in range(npoint):
        idx[:, i] = far
        c = xyz[ar, far][:, None, :]
        d = torch.sum((xyz - c) ** 2, -1)
        dist = torch.minimum(dist, d)
        far = torch.max(dist, -1)[1]
    return idx


def ball_query(radius, nsample, xyz, new_xyz):
    """pointnet2_utils.py:84-104: first ``nsample`` indices (ascending) with d <= r^2, padded
    with the first hit.  (Selection by prefix count instead of the reference's full sort.)"""
    C, N, _ = xyz.shape
    d = square_distance(new_xyz, xyz)
    r2 = torch.tensor(radius**2, dtype=d.dtype)  # python double -> tensor dtype, as torch does for `>`
    inside = ~(d > r2)
    rank = torch.cumsum(inside, -1) - 1
    S = new_xyz.shape[1]
    out = torch.full((C, S, nsample), N, dtype=torch.long)
    sel = inside & (rank < nsample)
    c, s, n = torch.nonzero(sel, as_tuple=True)
    out[c, s, rank[c, s, n]] = n
    first = out[:, :, :1].expand(-1, -1, nsample)
    return torch.where(out == N, first, out)


def _gather(points, idx):
    """pointnet2_utils.py:41-57 index_points."""
    C = points.shape[0]
    ar = torch.arange(C).view(C, *([1] * (idx.dim() - 1)))
    return points[ar, idx]


def _batch_norm(sd, bn, y, train):
    """nn.BatchNorm{1,2}d on channel-last rows.  eval: running statistics.  train (a dict): batch statistics over every
    row (biased variance), and the running-stat update torch applies (momentum 0.1, unbiased variance,
    num_batches_tracked + 1) is recorded in train['updates']."""
    if train is None:
        mean, var = sd[bn + ".running_mean"], sd[bn + ".running_var"]
    else:
        flat = y.reshape(-1, y.shape[-1])
        n = flat.shape[0]
        mean = flat.mean(0)
        var = flat.var(0, unbiased=False)
        upd = train.setdefault("updates", {})
        upd[bn + ".running_mean"] = 0.9 * sd[bn + ".running_mean"] + 0.1 * mean
        upd[bn + ".running_var"] = 0.9 * sd[bn + ".running_var"] + 0.1 * var * (n / (n - 1))
        upd[bn + ".num_batches_tracked"] = sd[bn + ".num_batches_tracked"] + 1
    return (y - mean) / torch.sqrt(var + BN_EPS) * sd[bn + ".weight"] + sd[bn + ".bias"]


def _conv_bn_relu(sd, prefix, i, x, train=None):
    """1x1 conv + BatchNorm + ReLU on channel-last rows (pointnet2_utils.py:192-195,308-311)."""
    w = sd[f"{prefix}.mlp_convs.{i}.weight"]
    w = w.reshape(w.shape[0], w.shape[1])
    y = x @ w.T + sd[f"{prefix}.mlp_convs.{i}.bias"]
    return torch.relu(_batch_norm(sd, f"{prefix}.mlp_bns.{i}", y, train))


def set_abstraction(sd, name, npoint, radius, nsample, xyz, feats, start, trace=None, train=None):
    """pointnet2_utils.py:107-135,174-199: FPS -> ball query -> [rel-xyz || feat] -> 3x(conv+BN+ReLU) -> max."""
    fps_idx = farthest_point_sample(xyz, npoint, start)
    new_xyz = _gather(xyz, fps_idx)
    grp = ball_query(radius, nsample, xyz, new_xyz)
    g_xyz = _gather(xyz, grp) - new_xyz[:, :, None, :]
    h = torch.cat([g_xyz, _gather(feats, grp)], dim=-1)
    for i in range(3):
        h = _conv_bn_relu(sd, f"pcd_backbone.{name}", i, h, train)
    out = h.max(dim=2)[0]
    if trace is not None:
        trace[name + ".fps_idx"] = fps_idx
        trace[name + ".group_idx"] = grp
        trace[name + ".xyz"] = new_xyz
        trace[name + ".feat"] = out
    return new_xyz, out


def feature_propagation(sd, name, nlayer, xyz1, xyz2, feats1, feats2, trace=None, train=None):
    """pointnet2_utils.py:273-312: 3-NN inverse-distance interpolation (w = 1/(d+1e-8), normalised),
    concat skip features, conv+BN+ReLU chain."""
    d = square_distance(xyz1, xyz2)
    dk, ik = torch.topk(d, 3, dim=-1, largest=False, sorted=True)
    rec = 1.0 / (dk + 1e-8)
    w = rec / rec.sum(-1, keepdim=True)
    interp = (_gather(feats2, ik) * w[..., None]).sum(dim=2)
    h = interp if feats1 is None else torch.cat([feats1, interp], dim=-1)
    for i in range(nlayer):
        h = _conv_bn_relu(sd, f"pcd_backbone.{name}", i, h, train)
    if trace is not None:
        trace[name + ".nn_idx"] = ik
        trace[name + ".nn_w"] = w
        trace[name + ".feat"] = h
    return h


def pointnet2_backbone(sd, clouds, fps_starts, trace=None, train=None, drop_mask=None):
    """model/pcd_backbone/pointnet2.py:61-80: clouds[C,1024,3] -> [C,1024,3].  ``train`` (dict) switches the 22 BatchNorm layers to
    batch statistics; ``drop_mask[C,128,1024]`` (values 0 or 2) is the Dropout(0.5) mask of drop1 (:76) in train mode."""
    l0_xyz, l0_f = clouds, clouds
    xyzs, feats = [l0_xyz], [l0_f]
    for (name, npoint, r, ns), start in zip(SA_SPECS, fps_starts):
        nx, nf = set_abstraction(sd, name, npoint, r, ns, xyzs[-1], feats[-1], start, trace, train)
        xyzs.append(nx)
        feats.append(nf)
    f3 = feature_propagation(sd, "fp4", 2, xyzs[3], xyzs[4], feats[3], feats[4], trace, train)
    f2 = feature_propagation(sd, "fp3", 2, xyzs[2], xyzs[3], feats[2], f3, trace, train)
    f1 = feature_propagation(sd, "fp2", 2, xyzs[1], xyzs[2], feats[1], f2, trace, train)
    f0 = feature_propagation(sd, "fp1", 3, xyzs[0], xyzs[1], None, f1, trace, train)
    w1 = sd["pcd_backbone.conv1.weight"][:, :, 0]
    h = f0 @ w1.T + sd["pcd_backbone.conv1.bias"]
    h = torch.relu(_batch_norm(sd, "pcd_backbone.bn1", h, train))
    if drop_mask is not None:
        h = h * drop_mask.permute(0, 2, 1).to(h.dtype)
    return h @ sd["pcd_backbone.conv2.weight"][:, :, 0].T + sd["pcd_backbone.conv2.bias"]


# --------------------------------------------------------------------------------------
# scene branch: scrambles, collapsed point attention, pointwise translate, masked sum
# --------------------------------------------------------------------------------------
def scramble1(Fb, w):
    """model/sdm.py:191-193: within a sample, flat g = c*9+o holds F[b,o,c]*w[b,o]; the buffer
    is then re-read as g = o'*3072 + p*3 + d.  Explicit index form (the CUDA kernels use it)."""
    B = Fb.shape[0]
    g = torch.arange(N_OBJ * N_PTS * 3)
    o, c = g % N_OBJ, g // N_OBJ
    flat = Fb[:, o, c] * w[:, o]
    return flat.view(B, N_OBJ, N_PTS, 3)


def point_attention(sd, tr, p1, as_written=False):
    """model/sdm.py:95,187-188,194-196: nn.MultiheadAttention(E=12, 12 heads, head_dim 1, scale 1).
    All 1024 query rows of a (b,o) are the same vector, so one query per (b,o) is exact.
    ``as_written`` materialises the [9B*12,1024,1024] weights like the reference (baseline timing)."""
    B = tr.shape[0]
    E = 12
    bias = sd["pcd_attention.in_proj_bias"]
    qq = tr @ sd["pcd_attention.q_proj_weight"].T + bias[:E]  # [B,9,12]
    kk = p1 @ sd["pcd_attention.k_proj_weight"].T + bias[E : 2 * E]  # [B,9,1024,12]
    vv = p1 @ sd["pcd_attention.v_proj_weight"].T + bias[2 * E :]
    if as_written:
        q = qq[:, :, None, :].expand(-1, -1, N_PTS, -1).reshape(B * N_OBJ, N_PTS, E).permute(0, 2, 1)  # [9B,12,1024]
        k = kk.reshape(B * N_OBJ, N_PTS, E).permute(0, 2, 1)
        v = vv.reshape(B * N_OBJ, N_PTS, E).permute(0, 2, 1)
        a = torch.softmax(q[..., :, None] * k[..., None, :], dim=-1)  # [9B,12,1024,1024]
        ctx = torch.einsum("chqk,chk->cqh", a, v)  # [9B,1024,12]
        pa = ctx @ sd["pcd_attention.out_proj.weight"].T + sd["pcd_attention.out_proj.bias"]
        return pa.view(B, N_OBJ, N_PTS, E)
    a = torch.softmax(qq[:, :, None, :] * kk, dim=2)  # over the 1024 keys, per head
    ctx = (a * vv).sum(dim=2)  # [B,9,12]
    pa = ctx @ sd["pcd_attention.out_proj.weight"].T + sd["pcd_attention.out_proj.bias"]
    return pa[:, :, None, :].expand(-1, -1, N_PTS, -1)


def scramble2_masked_sum(pw, mask_global, b_offset):
    """model/sdm.py:199-202: global flat f = ((b*9+o)*1024+p)*3+d is multiplied by
    mask[(f div 9) mod B, f mod 9] (B = GLOBAL batch), then summed over objects."""
    B = pw.shape[0]
    Bg = mask_global.shape[0]
    per = N_OBJ * N_PTS * 3
    f = (torch.arange(B)[:, None] + b_offset) * per + torch.arange(per)[None, :]
    m = mask_global[(f // N_OBJ) % Bg, f % N_OBJ]
    return (pw.reshape(B, per) * m).view(B, N_OBJ, N_PTS, 3).sum(dim=1)


def point_net(sd, z, emb):
    """model/diffusion_utils.py:51-62,74-78,98-103,110: InputProcess then OutputProcess."""
    h = torch.sigmoid(_lin(sd, "input_process.pose_embedding.0", z))
    h = torch.sigmoid(_lin(sd, "input_process.pose_embedding.2", h))
    h = torch.cat([h, emb], dim=-1)
    h = torch.sigmoid(_lin(sd, "input_process.combination_extraction.0", h))
    h = torch.sigmoid(_lin(sd, "input_process.combination_extraction.2", h))
    h = _gelu(_lin(sd, "output_process.pose_final.0", h))
    return _gelu(_lin(sd, "output_process.pose_final.2", h))


# --------------------------------------------------------------------------------------
# forward / sampling / training
# --------------------------------------------------------------------------------------
def encode_conditions(sd, mask_global, given_objs, given_cats, text, fps_starts, b_offset=0,
                      as_written=False, trace=None, train=None, drop_mask=None):
    """Everything in model/sdm.py:147-203 that does not depend on x or t.
    Returns dict(enc, out_cat, w, pcd_out) where pcd_out is the quantity added to x in place."""
    B = given_objs.shape[0]
    enc = text_embedding(sd, text)
    out_cat = predict_category(sd, enc.detach())  # model/sdm.py:156: predict_cat sees enc_text.clone().detach()
    emb_cat = _gelu(_lin(sd, "embed_cat.0", given_cats))
    hm = human_decoder(sd, given_objs[:, 0])
    Fb = pointnet2_backbone(sd, given_objs.reshape(B * N_OBJ, N_PTS, 3), fps_starts, trace, train, drop_mask).reshape(B, N_OBJ, N_PTS * 3)
    w = object_attention_weights(sd, enc, emb_cat, mask_global, b_offset)
    tr = translation_params(sd, enc, emb_cat)
    p1 = scramble1(Fb, w)
    pa = point_attention(sd, tr, p1, as_written)
    pw = _gelu(_lin(sd, "point_wise_trans_layer.0", torch.cat([p1, pa], dim=-1)))
    scene = scramble2_masked_sum(pw, mask_global, b_offset)
    pcd_out = (scene + hm) / 2
    if trace is not None:
        trace.update(enc=enc, emb_cat=emb_cat, hm=hm, backbone=Fb, attn_w=w, tr=tr, p1=p1, pa=pa[:, :, 0], pw=pw, scene=scene)
    return {"enc": enc, "out_cat": out_cat[:, None, :], "w": w, "pcd_out": pcd_out}


def forward(sd, x, mask, t, given_objs, given_cats, text, fps_starts, mask_global=None, b_offset=0,
            as_written=False, trace=None, cond=None, train=None, drop_mask=None):
    """model/sdm.py:131-218.  MUTATES ``x`` in place (x += pcd_out, :204) like the reference.
    Returns (out_cat[B,1,C], x0[B,1024,3], guiding[B,1024,3])."""
    dtype = x.dtype
    if mask_global is None:
        mask_global = mask
    if cond is None:
        cond = encode_conditions(sd, mask_global.to(dtype), given_objs.to(dtype), given_cats.to(dtype), text.to(dtype),
                                 fps_starts, b_offset, as_written, trace, train, drop_mask)
    ts = timestep_embedding(sd, t)
    emb = upsample_embedding(sd, ts, cond["enc"])
    x += cond["pcd_out"]
    x0 = point_net(sd, x, emb)
    guiding = point_net(sd, cond["pcd_out"], emb)
    if trace is not None:
        trace.update(ts=ts, emb=emb, pcd_out=cond["pcd_out"], x0=x0, guiding=guiding)
    return cond["out_cat"], x0, guiding


def p_sample(sd, tables, x, mask, t, given_objs, given_cats, text, fps_starts, noise, **kw):
    """diffusion/gaussian_diffusion.py:282-393,501-561 for START_X / FIXED_SMALL / clip_denoised=False.
    ``x`` is mutated by the model and the MUTATED tensor is the x_t of the posterior mean (trap 1)."""
    out_cat, x0, guiding = forward(sd, x, mask, t, given_objs, given_cats, text, fps_starts, **kw)
    c1 = _extract(tables["posterior_mean_coef1"], t, x)
    c2 = _extract(tables["posterior_mean_coef2"], t, x)
    logvar = _extract(tables["posterior_log_variance_clipped"], t, x)
    mean = c1 * x0 + c2 * x
    nonzero = (t != 0).to(x.dtype).view(-1, 1, 1)
    sample = mean + nonzero * torch.exp(0.5 * logvar) * noise
    return {"sample": sample, "pred_xstart": x0, "mean": mean, "out_cat": out_cat, "guiding": guiding}


def p_sample_loop(sd, tables, x_T, mask, given_objs, given_cats, text, fps_starts_all, noise_all,
                  hoisted=False, dtype=torch.float32, **kw):
    """diffusion/gaussian_diffusion.py:684-759.  ``fps_starts_all[T,4,9B]`` and ``noise_all[T,B,1024,3]`` are
    consumed in loop order (first entry = the first executed step, i.e. t=T-1).
    hoisted=True encodes the conditions once with the FIRST step's FPS starts."""
    sdc = _cast(sd, dtype)
    T = len(tables["betas"])
    img = x_T.clone().to(dtype)
    cond = None
    out = None
    for k, i in enumerate(range(T - 1, -1, -1)):
        t = torch.full((img.shape[0],), i, dtype=torch.long)
        if hoisted and cond is None:
            mg = kw.get("mask_global", None)
            cond = encode_conditions(sdc, (mask if mg is None else mg).to(dtype), given_objs.to(dtype), given_cats.to(dtype),
                                     text.to(dtype), fps_starts_all[0], kw.get("b_offset", 0))
        out = p_sample(sdc, tables, img, mask, t, given_objs, given_cats, text,
                       None if hoisted else fps_starts_all[k], noise_all[k].to(dtype),
                       cond=cond if hoisted else None, **kw)
        img = out["sample"]
    return out


def q_sample(tables, x_start, t, noise):
    """diffusion/gaussian_diffusion.py:238-256."""
    return _extract(tables["sqrt_alphas_cumprod"], t, x_start) * x_start + _extract(
        tables["sqrt_one_minus_alphas_cumprod"], t, x_start) * noise


def chamfer_distance(x, y):
    """pytorch3d.loss.chamfer_distance (0.7.3) with default arguments: squared-L2 nearest
    neighbour in both directions, mean over points, mean over batch, the two directions summed.
    (Restated from the package's documented definition; the package is not installed.)"""
    d = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1)
    return d.min(2)[0].mean(1).mean() + d.min(1)[0].mean(1).mean()


def training_losses(sd, tables, x_start, mask, t, given_objs, given_cats, target_cat, text, fps_starts, noise,
                    lambda_cat=0.1, **kw):
    """diffusion/gaussian_diffusion.py:1256-1342 (MSE loss type, START_X), eval-mode BatchNorm.
    cat_loss = lambda_cat * CrossEntropy(softmaxed probabilities, argmax(target_cat)) (:1297-1301);
    'mse' is the chamfer distance between the model output and x_start (:1334)."""
    x_t = q_sample(tables, x_start.float(), t, noise)
    out_cat, x0, _ = forward(sd, x_t, mask, t, given_objs, given_cats, text, fps_starts, **kw)
    cat_loss = F.cross_entropy(out_cat[:, 0], target_cat.argmax(dim=1)) * lambda_cat
    mse = chamfer_distance(x0.float(), x_start.float())
    return {"cat_loss": cat_loss, "mse": mse, "loss": mse + cat_loss, "model_output": x0}


def training_grads(sd, tables, x_start, mask, t, given_objs, given_cats, target_cat, text, fps_starts, noise, drop_mask,
                   lambda_cat=0.1):
    """Gradient oracle for SURVEY 8(f) row 1 (run/train_sdm.py:78-84: ``loss.backward()`` in ``model.train()`` mode): autograd on
    this file's functional forward.  Returns (loss, {parameter name: gradient or None}) for every floating-point entry of ``sd``
    that is not a BatchNorm running statistic.  Pinned against the live reference by tests/golden/make_golden_grads.py."""
    leaves = {}
    for k, v in sd.items():
        if torch.is_floating_point(v) and not k.endswith(("running_mean", "running_var")):
            leaves[k] = v.detach().clone().requires_grad_(True)
    sd2 = {k: leaves.get(k, v) for k, v in sd.items()}
    terms = training_losses(sd2, tables, x_start, mask, t, given_objs, given_cats, target_cat, text, fps_starts, noise,
                            lambda_cat=lambda_cat, train={}, drop_mask=drop_mask)
    names = list(leaves)
    grads = torch.autograd.grad(terms["loss"], [leaves[k] for k in names], allow_unused=True)
    return terms["loss"].detach(), dict(zip(names, grads))
