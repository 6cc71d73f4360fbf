"""Mirror of the reference's ``contact_former/transformer.py`` attention layer (same class names, constructor arguments,
parameter names / ``state_dict`` keys and ``forward`` signatures) whose forward runs in ``liblsdm_b200.so``
(``csrc/cf_mha.cu``): ``MultiHeadAttention`` (:44-103), ``PositionwiseFeedForward`` (:153-177), ``EncoderLayer`` /
``DecoderLayer`` (:180-207).  Inference (eval-mode) forward only: the Dropouts are identities; ``forward`` raises in
training mode and on CPU tensors -- there is no fallback path.

``precision``: "tf32" (default: tcgen05 TF32 projections and tensor-core attention, 3e-4 rel-L2 vs the fp32 oracle at the
reference's 256 x 655 shape, inside north_star's 1e-3), "3xtf32" / "fp32" (fp32-grade, 1e-6; attention on CUDA cores).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from .. import _lib
from ..engine import _ptr, _stream

_PREC = {"fp32": 0, "tf32": 1, "3xtf32": 2}


def _check(x, module):
    if module.training:
        raise NotImplementedError("lsdm_b200 contact_former layers implement the eval-mode forward only (call .eval())")
    if not x.is_cuda:
        raise _lib.LsdmError(_lib.ESTATE, "lsdm_b200 needs CUDA tensors: there is no CPU fallback")
    if x.dim() != 4 or x.shape[-1] != 64:
        raise ValueError(f"expected x of shape (bs, seg_len, n_verts, 64), got {tuple(x.shape)}")


class _Workspace:
    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes + 256 or self.buf.device != device:
            self.buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        return (self.buf.data_ptr() + 255) & ~255


class MultiHeadAttention(nn.Module):
    """Reference ``contact_former/transformer.py:44-103``."""

    def __init__(self, n_head, d_in, d_k, d_v, precision="tf32"):
        super().__init__()
        if (d_in, d_k, d_v) != (64, 64, 64):
            raise NotImplementedError("the CUDA layer covers the reference's configuration d_in = d_k = d_v = 64")
        self.n_head, self.d_in, self.d_k, self.d_v = n_head, d_in, d_k, d_v
        self.w_q = nn.Linear(d_in, n_head * d_k)
        self.w_k = nn.Linear(d_in, n_head * d_k)
        self.w_v = nn.Linear(d_in, n_head * d_v)
        self.temperature = np.power(d_k, 0.5)
        self.fc = nn.Linear(n_head * d_v, d_in)
        self.layer_norm = nn.LayerNorm(d_in)
        self.precision = precision
        self._ws = _Workspace()
        nn.init.normal_(self.w_q.weight, mean=0, std=np.sqrt(2.0 / (d_in + d_k)))
        nn.init.normal_(self.w_k.weight, mean=0, std=np.sqrt(2.0 / (d_in + d_k)))
        nn.init.normal_(self.w_v.weight, mean=0, std=np.sqrt(2.0 / (d_in + d_v)))
        nn.init.xavier_normal_(self.fc.weight)

    def forward(self, x, mask=None):
        _check(x, self)
        bs, seg_len, n_verts, _ = x.shape
        x = x.float().contiguous()
        dev = x.device
        m8, all_masked = None, 0
        if mask is not None:
            m8 = (mask.to(dev) != 0).to(torch.uint8).contiguous()
            if m8.shape != (bs, seg_len, seg_len):
                raise ValueError(f"mask must be (bs, seg_len, seg_len), got {tuple(mask.shape)}")
            all_masked = int(not bool(m8.any()))  # the reference's `mask.sum() == 0` (transformer.py:91), a host decision there too
        p = [t.detach().float().contiguous() for t in (self.w_q.weight, self.w_q.bias, self.w_k.weight, self.w_k.bias, self.w_v.weight,
                                                       self.w_v.bias, self.fc.weight, self.fc.bias, self.layer_norm.weight, self.layer_norm.bias)]
        w = _lib.CfMhaWeights(*[t.data_ptr() for t in p])
        lib = _lib.load()
        need = lib.lsdm_cf_workspace_bytes(bs, seg_len, n_verts, self.n_head)
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            ws = self._ws.get(need, dev)
            _lib.check(lib.lsdm_cf_mha_forward(C.byref(w), _ptr(x), _ptr(m8), all_masked, bs, seg_len, n_verts, self.n_head,
                                               _PREC[self.precision], C.c_void_p(ws), C.c_size_t(need), _ptr(out), _stream(dev)))
        self._keep = (p, m8, x)
        return out


class PositionwiseFeedForward(nn.Module):
    """Reference ``contact_former/transformer.py:153-177``."""

    def __init__(self, d_in, d_hid, precision="tf32"):
        super().__init__()
        if d_in != 64:
            raise NotImplementedError("the CUDA layer covers the reference's configuration d_in = 64")
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = nn.LayerNorm(d_in)
        self.d_hid, self.precision = d_hid, precision
        self._ws = _Workspace()
        torch.nn.init.xavier_uniform_(self.w_1.weight)
        torch.nn.init.xavier_uniform_(self.w_2.weight)

    def forward(self, x):
        _check(x, self)
        x = x.float().contiguous()
        dev, rows = x.device, x.numel() // 64
        p = [t.detach().float().contiguous() for t in (self.w_1.weight.squeeze(-1), self.w_1.bias, self.w_2.weight.squeeze(-1), self.w_2.bias,
                                                       self.layer_norm.weight, self.layer_norm.bias)]
        w = _lib.CfFfnWeights(*[t.data_ptr() for t in p], self.d_hid, 0)
        need = rows * (self.d_hid + 64) * 4 + 512
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            ws = self._ws.get(need, dev)
            _lib.check(_lib.load().lsdm_cf_ffn_forward(C.byref(w), _ptr(x), rows, _PREC[self.precision], C.c_void_p(ws), C.c_size_t(need),
                                                       _ptr(out), _stream(dev)))
        self._keep = (p, x)
        return out


class EncoderLayer(nn.Module):
    """Reference ``contact_former/transformer.py:180-192`` (``DecoderLayer`` :195-207 is identical)."""

    def __init__(self, n_head, d_in, d_k, d_v, precision="tf32"):
        super().__init__()
        self.self_attn = MultiHeadAttention(n_head, d_in, d_k, d_v, precision)
        self.pos_wise_ffnn = PositionwiseFeedForward(d_in, d_in, precision)

    def forward(self, x, mask=None):
        return self.pos_wise_ffnn(self.self_attn(x, mask))


DecoderLayer = EncoderLayer
