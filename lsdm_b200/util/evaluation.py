"""Evaluation metrics of the sampling path -- mirror of the reference's ``util/evaluation.py`` (same names, argument
meaning and return values) plus batched forms, computed on the device by ``liblsdm_b200.so``.

Reference call sites: ``run/test_sdm.py:186-207`` (Chamfer, EMD, F-score, top-k accuracy right after ``p_sample_loop``).
There is no CPU fallback: inputs are moved to the current CUDA device and every number comes from a CUDA kernel
(``lsdm_b200/csrc/eval_metrics.cu``, ``loss.cu``).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..engine import _ptr, _stream


def _dev(device=None):
    if not torch.cuda.is_available():
        raise _lib.LsdmError(_lib.ESTATE, "lsdm_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _cloud(t, device):
    t = torch.as_tensor(t)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.shape[-1] != 3:
        raise ValueError(f"expected [n,3] or [B,n,3] points, got {tuple(t.shape)}")
    return t.to(device=device, dtype=torch.float32).contiguous()


def emd_batch(x, y, device=None, return_assignment=False):
    """Batched ``emd``: x, y ``[B,n,3]`` -> float64 tensor ``[B]`` (and optionally the int32 matching ``[B,n]``)."""
    dev = _dev(device if device is not None else (x.device if torch.is_tensor(x) and x.is_cuda else None))
    x, y = _cloud(x, dev), _cloud(y, dev)
    if x.shape[0] != y.shape[0]:
        raise ValueError("batch mismatch")
    B, n, m = x.shape[0], x.shape[1], y.shape[1]
    out = torch.empty(B, dtype=torch.float64, device=dev)
    assign = torch.empty(B, n, dtype=torch.int32, device=dev) if return_assignment else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsdm_eval_emd(_ptr(x), _ptr(y), B, n, m, _ptr(out), _ptr(assign), None, _stream(dev)))
    return (out, assign) if return_assignment else out


def emd(x, y):
    """Reference ``util/evaluation.py:5-11``: mean matched distance of the min-cost assignment between two clouds
    (``[n,3]`` or ``[1,n,3]``); returns a Python float."""
    x, y = torch.as_tensor(x), torch.as_tensor(y)
    if x.dim() == 3:
        x, y = x.squeeze(0), y.squeeze(0)
    return float(emd_batch(x, y)[0].item())


def fscore_batch(gt, pr, th: float = 0.1, device=None):
    """Batched ``calculate_fscore``: ``[B,n,3]``, ``[B,m,3]`` -> float64 ``[B,3]`` = (fscore, precision, recall)."""
    dev = _dev(device if device is not None else (gt.device if torch.is_tensor(gt) and gt.is_cuda else None))
    gt, pr = _cloud(gt, dev), _cloud(pr, dev)
    B, n, m = gt.shape[0], gt.shape[1], pr.shape[1]
    counts = torch.empty(B, 2, dtype=torch.int32, device=dev)
    out = torch.empty(B, 3, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsdm_eval_fscore(_ptr(gt), _ptr(pr), B, n, m, C.c_double(float(th)), _ptr(counts), _ptr(out),
                                                _stream(dev)))
    return out


def calculate_fscore(gt_tensor, pr_tensor, th: float = 0.1):
    """Reference ``util/evaluation.py:28-52``: returns ``(fscore, precision, recall)`` as Python floats."""
    f, p, r = fscore_batch(gt_tensor, pr_tensor, th)[0].tolist()
    return f, p, r


def accuracy(output, target, topk=(1,)):
    """Reference ``util/evaluation.py:13-26``: precision@k in percent, one 0-d tensor per k."""
    dev = _dev(output.device if output.is_cuda else None)
    output = output.to(device=dev, dtype=torch.float32).contiguous()
    target = target.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    B, Cn = output.shape
    ks = torch.tensor([int(k) for k in topk], dtype=torch.int32, device=dev)
    correct = torch.empty(len(topk), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsdm_eval_topk(_ptr(output), _ptr(target), B, Cn, _ptr(ks), len(topk), _ptr(correct), _stream(dev)))
    return [c.float().mul_(100.0 / B) for c in correct]


def chamfer_batch(x, y, device=None):
    """Per-sample ``pytorch3d.loss.chamfer_distance`` (default arguments): float32 ``[B]``."""
    dev = _dev(device if device is not None else (x.device if torch.is_tensor(x) and x.is_cuda else None))
    x, y = _cloud(x, dev), _cloud(y, dev)
    B, n, m = x.shape[0], x.shape[1], y.shape[1]
    per = torch.empty(B, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsdm_eval_chamfer(_ptr(x), _ptr(y), B, n, m, _ptr(per), _stream(dev)))
    return per.sum(1)


def chamfer_distance(x, y):
    """``pytorch3d.loss.chamfer_distance(x, y)`` as called at ``run/test_sdm.py:187``: ``(loss, None)`` with the batch mean."""
    return chamfer_batch(x, y).mean(), None
