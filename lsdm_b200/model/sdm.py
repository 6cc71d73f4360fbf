"""``SceneDiffusionModel`` -- drop-in mirror of the reference's ``model/sdm.py:18-218``.

Same constructor keywords, same ``state_dict`` keys and shapes, same
``forward(x, mask, timesteps, given_objs, given_cats, y=None, force_mask=False) -> (out_cat, x0)``
including its side effects: ``x`` is mutated in place (``x += pcd_out``, reference sdm.py:204),
``saved_cat`` and ``saved_guiding_points`` are set, and the four ``torch.randint`` FPS start draws
of ``farthest_point_sample`` (reference pointnet2_utils.py:72) are consumed from the CPU generator
in the same order and shapes.

The ``nn`` sub-modules below only HOLD parameters (so checkpoints load strictly and optimisers /
``.parameters()`` work); none of their ``forward`` methods is ever called.  All arithmetic runs in
``liblsdm_b200.so`` through :class:`lsdm_b200.engine.Engine`; without the CUDA library or a GPU,
``forward`` raises -- there is no PyTorch fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..synthetic import positional_table

N_POINTS, N_OBJ = 1024, 9
FPS_LEVEL_N = (1024, 1024, 256, 64)

_TORCH_RANDINT = torch.randint  # (tests feed FPS starts by patching torch.randint: the batched draw below then steps aside)
_BATCHED_DRAW_OK = None


def _batched_draw_is_sequential():
    """One ``randint(0, 1024, (k, 4, n))`` call yields the numbers -- and leaves the generator in the state -- of 4k consecutive
    ``randint(0, N_l, (n,))`` calls when torch's CPU ``randint`` consumes one 32-bit word per element in order and reduces it
    modulo the range (every N_l is a power of two dividing 1024, so ``word % N_l == (word % 1024) & (N_l - 1)``).  Checked once per
    process on a copy of the generator state (the caller's stream is left untouched); on any mismatch the per-call draws stay."""
    global _BATCHED_DRAW_OK
    if _BATCHED_DRAW_OK is None:
        ok = all(n & (n - 1) == 0 and max(FPS_LEVEL_N) % n == 0 for n in FPS_LEVEL_N)
        if ok:
            state = torch.get_rng_state()
            try:
                torch.manual_seed(20231017)
                seq = torch.stack([torch.stack([torch.randint(0, n, (37,), dtype=torch.long) for n in FPS_LEVEL_N]) for _ in range(3)])
                s_seq = torch.get_rng_state()
                torch.manual_seed(20231017)
                raw = torch.randint(0, max(FPS_LEVEL_N), (3, len(FPS_LEVEL_N), 37), dtype=torch.long)
                s_bat = torch.get_rng_state()
                mask = torch.tensor([n - 1 for n in FPS_LEVEL_N]).view(1, -1, 1)
                ok = bool(torch.equal(raw & mask, seq)) and bool(torch.equal(s_seq, s_bat))
            finally:
                torch.set_rng_state(state)
        _BATCHED_DRAW_OK = ok
    return _BATCHED_DRAW_OK


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the arithmetic of this layer runs in liblsdm_b200.so")


class _PositionalEncoding(_Holder):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.register_buffer("pe", positional_table(max_len, d_model))


class _TimestepEmbedder(_Holder):
    def __init__(self, latent_dim, sequence_pos_encoder):
        super().__init__()
        self.sequence_pos_encoder = sequence_pos_encoder
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.SiLU(), nn.Linear(latent_dim, latent_dim))


class _SetAbstraction(_Holder):
    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for c in mlp:
            self.mlp_convs.append(nn.Conv2d(last, c, 1))
            self.mlp_bns.append(nn.BatchNorm2d(c))
            last = c


class _FeaturePropagation(_Holder):
    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for c in mlp:
            self.mlp_convs.append(nn.Conv1d(last, c, 1))
            self.mlp_bns.append(nn.BatchNorm1d(c))
            last = c


class _PointNet2Backbone(_Holder):
    """Parameter layout of reference model/pcd_backbone/pointnet2.py:43-59."""

    def __init__(self, num_classes=3, dimension=3):
        super().__init__()
        self.sa1 = _SetAbstraction(dimension + 3, [32, 32, 64])
        self.sa2 = _SetAbstraction(64 + 3, [64, 64, 128])
        self.sa3 = _SetAbstraction(128 + 3, [128, 128, 256])
        self.sa4 = _SetAbstraction(256 + 3, [256, 256, 512])
        self.fp4 = _FeaturePropagation(768, [256, 256])
        self.fp3 = _FeaturePropagation(384, [256, 256])
        self.fp2 = _FeaturePropagation(320, [256, 128])
        self.fp1 = _FeaturePropagation(128, [128, 128, 128])
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)


class _GraphBlock(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _Holder()
        self.conv.layer = nn.Linear(cin, cout)
        nn.init.xavier_uniform_(self.conv.layer.weight)
        nn.init.constant_(self.conv.layer.bias, 0)
        self.norm = nn.GroupNorm(8, cout)


class _SpiralOut(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.layer = nn.Linear(cin, cout)
        nn.init.xavier_uniform_(self.layer.weight)
        nn.init.constant_(self.layer.bias, 0)


class _HumanDecoder(_Holder):
    """Parameter layout of reference posa/posa_models.py:292-317 (seq_length=1)."""

    def __init__(self):
        super().__init__()
        self.de_spiral = nn.Sequential(_GraphBlock(3, 64), _GraphBlock(64, 64), _GraphBlock(64, 64), _SpiralOut(64, 3))


class _InputProcess(_Holder):
    def __init__(self, input_feats, extract_dim):
        super().__init__()
        self.pose_embedding = nn.Sequential(nn.Linear(input_feats, extract_dim // 2), nn.Sigmoid(),
                                            nn.Linear(extract_dim // 2, extract_dim), nn.Sigmoid())
        self.combination_extraction = nn.Sequential(nn.Linear(extract_dim * 2, int(extract_dim * 1.5)), nn.Sigmoid(),
                                                    nn.Linear(int(extract_dim * 1.5), extract_dim), nn.Sigmoid())


class _OutputProcess(_Holder):
    def __init__(self, input_feats, extract_dim):
        super().__init__()
        self.pose_final = nn.Sequential(nn.Linear(extract_dim, extract_dim // 2), nn.GELU(),
                                        nn.Linear(extract_dim // 2, input_feats), nn.GELU())


class SceneDiffusionModel(nn.Module):
    def __init__(self, seg_len=256, modality='text', clip_version='ViT-B/32', clip_dim=768, dropout=0.1, n_layer=6,
                 n_head=8, f_vert=64, dim_ff=512, cat_emb=32, mesh_ds_dir="data/mesh_ds", posa_path=None, latent_dim=128,
                 cond_mask_prob=1.0, device=0, vert_dims=655, obj_cat=8, data_rep='rot6d', njoints=251, use_cuda=True,
                 pcd_points=1024, pcd_dim=128, xyz_dim=3, max_cats=13, translation_params=12, pcd_backbone_type="PNT2",
                 human_backbone_type="POSA", text_encoder_type="CLIP", **kwargs) -> None:
        super().__init__()
        # The kernels are specialised to the factory configuration of reference util/model_util.py:26-73.
        if (latent_dim, clip_dim, pcd_points, pcd_dim, xyz_dim, n_head, cat_emb, translation_params) != (128, 512, 1024, 3, 3, 8, 32, 12):
            raise NotImplementedError("lsdm_b200 implements the reference's factory configuration "
                                      "(latent 128, clip 512, 1024 points, pcd_dim 3, 8 heads, cat_emb 32, 12 translation params)")
        if pcd_backbone_type != "PNT2" or human_backbone_type != "POSA":
            raise NotImplementedError("only the default PNT2 / POSA backbones are on the accelerated path")
        if data_rep not in ('rot6d', 'xyz', 'hml_vec'):
            raise ValueError(data_rep)
        self.seg_len, self.pcd_points, self.clip_version, self.clip_dim = seg_len, pcd_points, clip_version, clip_dim
        self.latent_dim, self.pcd_dim, self.xyz_dim, self.extract_dim = latent_dim, pcd_dim, xyz_dim, latent_dim
        self.dropout, self.cond_mask_prob, self.data_rep = dropout, cond_mask_prob, data_rep
        self.input_feats = vert_dims * obj_cat
        self.n_head, self.translation_params, self.text_encoder_type = n_head, translation_params, text_encoder_type
        self.max_cats = max_cats
        self.device = "cuda:{}".format(device) if use_cuda else "cpu"
        self.modality = modality
        assert self.modality in ['text', 'audio', None]

        self.sequence_pos_encoder = _PositionalEncoding(latent_dim)
        self.embed_timestep = _TimestepEmbedder(latent_dim, self.sequence_pos_encoder)
        self.saved_cat = None
        self.embed_text = nn.Sequential(nn.Linear(clip_dim, clip_dim // 2), nn.GELU(), nn.Linear(clip_dim // 2, latent_dim * 2),
                                        nn.GELU(), nn.Linear(latent_dim * 2, latent_dim), nn.GELU())
        self.embed_cat = nn.Sequential(nn.Linear(max_cats, cat_emb), nn.GELU())
        self.predict_cat = nn.Sequential(nn.Linear(latent_dim, latent_dim // 2), nn.GELU(), nn.Linear(latent_dim // 2, latent_dim // 4),
                                         nn.GELU(), nn.Linear(latent_dim // 4, max_cats), nn.GELU(), nn.Softmax(dim=2))
        self.attn_layer = nn.MultiheadAttention(embed_dim=latent_dim, num_heads=n_head, kdim=cat_emb, vdim=pcd_points * pcd_dim,
                                                batch_first=True)
        self.translation_layer = nn.Sequential(nn.Linear(latent_dim + cat_emb, latent_dim), nn.GELU(),
                                               nn.Linear(latent_dim, translation_params), nn.GELU())
        self.point_wise_trans_layer = nn.Sequential(nn.Linear(translation_params + xyz_dim, xyz_dim), nn.GELU())
        self.pcd_attention = nn.MultiheadAttention(embed_dim=translation_params, num_heads=translation_params, kdim=xyz_dim,
                                                   vdim=xyz_dim, batch_first=True)
        self.pcd_backbone = _PointNet2Backbone(pcd_dim)
        self.human_backbone = _HumanDecoder()
        self.upsampling_layer = nn.Sequential(nn.Linear(1, 128), nn.GELU(), nn.Linear(128, 512), nn.GELU(),
                                              nn.Linear(512, pcd_points), nn.GELU())
        self.combine_extraction = nn.Sequential(nn.Linear(latent_dim * 2, latent_dim), nn.GELU())
        self.input_process = _InputProcess(xyz_dim, latent_dim)
        self.output_process = _OutputProcess(xyz_dim, latent_dim)
        self.saved_guiding_points = None

        # engine state (not part of the state dict)
        self._engine = None
        self._engine_sig = None
        self._weights_sig = None
        self._shard = None  # (batch_global, batch_offset)
        self._sync_bn_group = None
        self._text_encoder = None
        if use_cuda and torch.cuda.is_available():
            self.to(self.device)

    # ------------------------------------------------------------------ reference-surface helpers
    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts an unmodified reference checkpoint.  ``clip_model.*`` entries (the frozen CLIP tower the reference keeps as
        a sub-module, model/sdm.py:229-233) do not belong to this module's parameters: when a CUDA device is present they are
        loaded into the device text tower (``self.clip_text``), otherwise they are dropped."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith("clip_model.")}
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._weights_sig = None
        if any(k.startswith("clip_model.transformer.") for k in state_dict) and torch.cuda.is_available():
            self.load_clip_state_dict(state_dict)
        return out

    def load_clip_state_dict(self, state_dict, prefix="clip_model."):
        """Installs the CLIP ViT-B/32 text tower (openai/CLIP state-dict names under ``prefix``) on the device."""
        from .clip_text import ClipTextTower

        dev = torch.device(self.device)
        tower = ClipTextTower(dev if dev.type == "cuda" else None)
        tower.load_state_dict(state_dict, prefix)
        self.clip_text = tower
        return tower

    def set_tokenizer(self, fn):
        """``fn(list[str], context_length=22, truncate=True) -> int [B, 22]``: stands in for ``clip.tokenize`` (host-side BPE;
        its vocabulary file is not shipped here).  With it and a loaded text tower, ``y`` may be a list of strings."""
        self._tokenizer = fn

    def set_tokenizer_vocab(self, path):
        """Path of openai/CLIP's ``bpe_simple_vocab_16e6.txt.gz``: strings are then tokenised by ``lsdm_b200.model.clip_tokenizer``
        (the same BPE as ``clip.tokenize``).  Also picked up from the ``LSDM_CLIP_BPE`` environment variable."""
        from .clip_tokenizer import ClipBpeTokenizer

        self._tokenizer = ClipBpeTokenizer(path)

    def set_text_encoder(self, fn):
        """``fn(list[str]) -> [B,512]`` float tensor: an external replacement for the whole CLIP text path."""
        self._text_encoder = fn

    def set_shard(self, batch_global=None, batch_offset=0, sync_bn_group=None):
        """Data-parallel sharding: this process holds samples [offset, offset+B) of a global batch; ``mask`` arguments
        must then be the GLOBAL ``[batch_global, 9]`` mask (the reference's mask scrambles index it)."""
        self._shard = None if batch_global is None else (int(batch_global), int(batch_offset))
        self._sync_bn_group = sync_bn_group  # torch.distributed group (or True for the default group) for train-mode SyncBN

    def _encode_text(self, y):
        """``enc_text`` of model/sdm.py:147-150.  float ``[B,512]`` -> already-encoded text (synthetic runs); integer
        ``[B,77]`` -> ``clip.tokenize`` output, encoded by the device CLIP tower; list of strings -> tokenizer hook + tower
        (model/sdm.py:245-259), or the external encoder hook."""
        if torch.is_tensor(y) and y.is_floating_point():
            return y.float()
        tower = getattr(self, "clip_text", None)
        if torch.is_tensor(y):
            if tower is None:
                raise NotImplementedError("token ids need the CLIP text tower: load a checkpoint with clip_model.* entries or call "
                                          "load_clip_state_dict()")
            return tower.encode_text(y)
        if self._text_encoder is not None:
            return self._text_encoder(list(y)).float()
        tok = getattr(self, "_tokenizer", None)
        if tok is None and tower is not None:
            import os
            if os.environ.get("LSDM_CLIP_BPE"):
                self.set_tokenizer_vocab(os.environ["LSDM_CLIP_BPE"])
                tok = self._tokenizer
        if tower is None or tok is None:
            raise NotImplementedError("text strings need the CLIP BPE vocabulary (set_tokenizer_vocab(path) / LSDM_CLIP_BPE, or set_tokenizer(clip.tokenize)) and the "
                                      "CLIP text tower weights (load_clip_state_dict()); or pass the [B,512] embedding / [B,77] token ids "
                                      "as `y`, or install an external encoder with set_text_encoder()")
        # model/sdm.py:248-256: tokenise to 20 + 2 positions, zero-pad to the 77-token context
        t = torch.as_tensor(tok(list(y), context_length=22, truncate=True))
        t = torch.cat([t, torch.zeros(t.shape[0], tower.dims["ctx"] - t.shape[1], dtype=t.dtype)], dim=1)
        return tower.encode_text(t)

    def _sig(self):
        """Per-tensor (version, storage pointer) of every parameter / buffer.  In-place updates through ``.data`` (EMA,
        ``p.data.copy_``) do not bump a tensor's version: call :meth:`invalidate_weights` after those."""
        sig = []
        for m in self.modules():  # (one walk of the module tree; runs once per sampling / training call)
            for t in m._parameters.values():
                if t is not None:
                    sig.append((t._version, t.data_ptr()))
            for t in m._buffers.values():
                if t is not None:
                    sig.append((t._version, t.data_ptr()))
        return tuple(sig)

    def invalidate_weights(self):
        """Forces the next call to re-upload this module's weights into the device handle."""
        self._weights_sig = None

    sync_weights = invalidate_weights

    def engine(self, batch_local, device):
        """The ``lsdm_handle`` owner for this batch size / shard; (re)uploads weights when they changed."""
        from ..engine import Engine

        device = torch.device(device)
        if device.type != "cuda":  # host tensors in: the handle lives where the parameters live
            device = next(self.parameters()).device
            if device.type != "cuda":
                device = torch.device(self.device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        bg, off = self._shard if self._shard is not None else (batch_local, 0)
        sig = (int(batch_local), bg, off, str(device))
        if self._engine is None or self._engine_sig[3] != sig[3]:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(batch_local, self.max_cats, device, bg, off)
            self._weights_sig = None
        elif self._engine_sig != sig:
            self._engine.set_batch(batch_local, bg, off)
        self._engine_sig = sig
        ws = self._sig()
        if self._weights_sig != ws:
            self._engine.load_state_dict(self.state_dict())
            self._weights_sig = ws
        return self._engine

    def draw_fps_starts(self, batch_local):
        """The four CPU-generator draws of reference pointnet2_utils.py:72, at GLOBAL shape, sliced to the shard."""
        bg, off = self._shard if self._shard is not None else (batch_local, 0)
        draws = [torch.randint(0, n, (bg * N_OBJ,), dtype=torch.long) for n in FPS_LEVEL_N]
        return torch.stack([d.view(bg, N_OBJ)[off:off + batch_local].reshape(-1) for d in draws])

    def draw_fps_starts_steps(self, batch_local, n_steps, first=None):
        """``n_steps`` consecutive :meth:`draw_fps_starts` results as one ``[n_steps, 4, 9 * batch_local]`` tensor (same generator calls in
        the same order, drawn straight into the output); ``first``: an already drawn set for step 0."""
        bg, off = self._shard if self._shard is not None else (batch_local, 0)
        out = torch.empty(n_steps, len(FPS_LEVEL_N), batch_local * N_OBJ, dtype=torch.long)
        k0 = 0
        if first is not None:
            out[0] = first
            k0 = 1
        if n_steps - k0 > 1 and torch.randint is _TORCH_RANDINT and _batched_draw_is_sequential():
            # one generator call for the whole run of steps (~0.2 ms per 20 steps instead of ~1.3 ms of per-call overhead)
            raw = torch.randint(0, max(FPS_LEVEL_N), (n_steps - k0, len(FPS_LEVEL_N), bg * N_OBJ), dtype=torch.long)
            raw &= torch.tensor([n - 1 for n in FPS_LEVEL_N]).view(1, -1, 1)
            out[k0:] = raw.view(n_steps - k0, len(FPS_LEVEL_N), bg, N_OBJ)[:, :, off:off + batch_local].reshape(n_steps - k0, len(FPS_LEVEL_N), -1)
        elif bg == batch_local and off == 0:
            for k in range(k0, n_steps):
                for lvl, n in enumerate(FPS_LEVEL_N):
                    torch.randint(0, n, (bg * N_OBJ,), out=out[k, lvl])
        else:
            for k in range(k0, n_steps):
                out[k] = self.draw_fps_starts(batch_local)
        return out

    def draw_dropout_mask(self, batch_local, device):
        """The mask ``F.dropout`` would draw for the reference's ``[9B,128,1024]`` head activation (pointnet2.py:76, p = 0.5, values 0
        or 2; same generator, GLOBAL shape, sliced to the shard)."""
        bg, off = self._shard if self._shard is not None else (batch_local, 0)
        full = torch.nn.functional.dropout(torch.ones(bg * N_OBJ, 128, N_POINTS, device=device), 0.5, True)
        return full[off * N_OBJ:(off + batch_local) * N_OBJ].contiguous()

    def encode(self, mask, given_objs, given_cats, y, fps_start=None, device=None, drop_mask=None):
        """Step-invariant part of forward (reference sdm.py:147-203).  Returns the engine."""
        B = given_objs.shape[0]
        if device is None:
            device = given_objs.device if given_objs.is_cuda else next(self.parameters()).device
        eng = self.engine(B, device)
        if fps_start is None:
            fps_start = self.draw_fps_starts(B)
        if self.training:
            # model.train(): BatchNorm batch statistics over all 9B clouds + Dropout(0.5) in the backbone head.  The mask is what
            # F.dropout would draw for the reference's [9B,128,1024] activation (same generator, same shape).
            if self._shard is not None:
                # BatchNorm couples all 9B clouds of the GLOBAL batch: SyncBN over the data-parallel group
                if self._sync_bn_group is None:
                    raise NotImplementedError("sharded train-mode forward needs set_shard(..., sync_bn_group=<process group>) (SyncBN)")
                eng.set_allreduce(self._sync_bn_group if self._sync_bn_group is not True else None)
            if drop_mask is None:
                drop_mask = self.draw_dropout_mask(B, eng.device)
            eng.encode_conditions_train(self._encode_text(y), given_objs, given_cats, mask, fps_start, drop_mask)
            self._pull_bn_stats(eng)
        else:
            eng.encode_conditions(self._encode_text(y), given_objs, given_cats, mask, fps_start)
        return eng

    def _pull_bn_stats(self, eng):
        """Mirror the running statistics the library just updated into this module's BatchNorm buffers (nn.BatchNorm semantics)."""
        with torch.no_grad():
            for name, mod in self.pcd_backbone.named_modules():
                if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                    key = "pcd_backbone." + name
                    mod.running_mean.copy_(eng.read_weight(key + ".running_mean", mod.running_mean))
                    mod.running_var.copy_(eng.read_weight(key + ".running_var", mod.running_var))
                    mod.num_batches_tracked += 1
        self._weights_sig = self._sig()  # the handle already holds these values

    # ------------------------------------------------------------------ forward
    def forward(self, x, mask, timesteps, given_objs, given_cats, y=None, force_mask=False):
        """
        x: [bs, 1024, 3] noisy target cloud (MUTATED in place); mask: [bs, 9] (global mask when sharded);
        timesteps: [bs]; given_objs: [bs, 9, 1024, 3]; given_cats: [bs, 9, max_cats]; y: list[str] or [bs, 512] embedding.
        Returns (out_cat [bs, 1, max_cats], x0 [bs, 1024, 3]).
        """
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise ValueError("x must be a contiguous float32 CUDA tensor (it is updated in place)")
        eng = self.encode(mask, given_objs, given_cats, y, device=x.device)
        out_cat, x0, guiding = eng.forward(x, timesteps)
        self.saved_cat = out_cat.unsqueeze(1)
        self.saved_guiding_points = guiding
        return self.saved_cat, x0
