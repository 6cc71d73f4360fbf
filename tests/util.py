"""Shared helpers for the test-suite (fixtures are regenerated from seeds; see tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-30))


import contextlib


@contextlib.contextmanager
def injected_rng(fps_starts=None, noises=None):
    """Feed ``torch.randint`` (FPS starts, CPU) and ``torch.randn_like`` (sampling noise) from queues, in draw order,
    so that the CUDA path, the oracle and the golden fixtures all see the same randoms."""
    fq = list(fps_starts) if fps_starts is not None else None
    nq = list(noises) if noises is not None else None
    o_randint, o_randn_like = torch.randint, torch.randn_like

    def randint(*a, **k):
        if fq is not None:
            assert fq, "FPS start queue exhausted"
            v = fq.pop(0).clone()
            if k.get("out") is not None:  # (the library draws straight into its per-chunk buffer)
                k["out"].copy_(v)
                return k["out"]
            return v
        return o_randint(*a, **k)

    def randn_like(x, **k):
        if nq is not None:
            assert nq, "noise queue exhausted"
            return nq.pop(0).clone().to(x.device)
        return o_randn_like(x, **k)

    torch.randint, torch.randn_like = randint, randn_like
    try:
        yield
    finally:
        torch.randint, torch.randn_like = o_randint, o_randn_like
