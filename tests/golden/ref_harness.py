"""Import and run the UNMODIFIED reference (``/root/reference``) in this container.

Test/fixture infrastructure only (never imported by the product path, never
shipped to the GPU box: ``/root/reference`` does not exist there).  Follows
SURVEY.md Appendix C: four stub modules for absent third-party packages
(``clip``, ``pytorch3d``, ``trimesh``, ``openmesh``) and two monkeypatches
(``extract_spirals`` for seq_length=1, ``load_ds_us_param`` forced to CPU).
No file under ``/root/reference`` is modified or copied.

The only behavioural hook is RNG injection: ``torch.randint`` / ``torch.randn_like``
are wrapped (in a context manager) so that the FPS start indices and sampling noise the
reference draws can be supplied from the seeded fixtures, in draw order.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("LSDM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "diffusion"))


def _chamfer(x, y, **kw):
    d = torch.cdist(x, y) ** 2
    return d.min(2)[0].mean(1).mean() + d.min(1)[0].mean(1).mean(), None


class _FakeClip(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def encode_text(self, x):
        return x


def _install_stubs():
    if "clip" not in sys.modules:
        clip = types.ModuleType("clip")
        clip.load = lambda version, device=None, jit=False: (_FakeClip(), None)
        clip.tokenize = lambda texts, context_length=77, truncate=True: torch.zeros(len(texts), context_length, dtype=torch.long)
        clip.model = types.ModuleType("clip.model")
        clip.model.convert_weights = lambda m: None
        sys.modules["clip"] = clip
        sys.modules["clip.model"] = clip.model
    if "pytorch3d" not in sys.modules:
        p3d = types.ModuleType("pytorch3d")
        p3d.loss = types.ModuleType("pytorch3d.loss")
        p3d.loss.chamfer_distance = _chamfer
        sys.modules["pytorch3d"] = p3d
        sys.modules["pytorch3d.loss"] = p3d.loss
    if "trimesh" not in sys.modules:
        tm = types.ModuleType("trimesh")

        class _Mesh:
            def __init__(self, v, f):
                self.vertices, self.faces = v, f

        def load(path, process=False):
            v, f = [], []
            with open(path) as fh:
                for line in fh:
                    if line.startswith("v "):
                        v.append([float(t) for t in line.split()[1:4]])
                    elif line.startswith("f "):
                        f.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
            return _Mesh(np.asarray(v), np.asarray(f))

        tm.load = load
        sys.modules["trimesh"] = tm
    if "openmesh" not in sys.modules:
        om = types.ModuleType("openmesh")

        class TriMesh:
            def __init__(self, v, f):
                self.v, self.f = v, f

        om.TriMesh = TriMesh
        sys.modules["openmesh"] = om


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's modules (imported once)."""
    if _loaded:
        return _loaded["ns"]
    assert available(), "reference tree not present"
    _install_stubs()
    os.chdir(REF)  # posa_models.py:293 opens ./mesh_ds relative to cwd
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import posa.posa_utils as posa_utils
    import posa.posa_models as posa_models

    # exact for seq_length=1 (posa_utils.py:149,169): each spiral is the vertex itself
    posa_utils.extract_spirals = lambda mesh, L, dilation=1: [[i] for i in range(len(mesh.v))]
    _orig = posa_models.load_ds_us_param
    posa_models.load_ds_us_param = lambda d, level, L, use_cuda=True: _orig(d, level, L, use_cuda=False)

    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import util.model_util as model_util
        import model.sdm as sdm
        import diffusion.gaussian_diffusion as gd
        import diffusion.respace as respace
        import model.pcd_backbone.pointnet2_utils as pn2u
    ns = types.SimpleNamespace(model_util=model_util, sdm=sdm, gd=gd, respace=respace, pn2u=pn2u)
    _loaded["ns"] = ns
    return ns


def build_reference_model(state_dict, max_cats: int = 13):
    ns = load_reference()
    kw = ns.model_util.get_default_model_proxd() if max_cats == 13 else ns.model_util.get_default_model_humanise()
    model = ns.sdm.SceneDiffusionModel(**kw, use_cuda=False)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("clip_model.") for k in missing), missing
    model._encode_text_clip = lambda y: y.float()
    model.eval()
    return model


@contextlib.contextmanager
def injected_rng(fps_starts=None, noises=None, dropout_masks=None):
    """Feed ``torch.randint`` (FPS starts) and ``torch.randn_like`` (sampling noise)
    from queues, in the order the reference draws them."""
    fq = list(fps_starts) if fps_starts is not None else None
    nq = list(noises) if noises is not None else None
    o_randint, o_randn_like = torch.randint, torch.randn_like
    dq = list(dropout_masks) if dropout_masks is not None else None
    o_dropout = torch.nn.functional.dropout

    def dropout(x, p=0.5, training=True, inplace=False):
        if dq is not None and training:
            assert dq, "dropout mask queue exhausted"
            m = dq.pop(0)
            assert m.shape == x.shape, (m.shape, x.shape)
            return x * m.to(x.dtype)
        return o_dropout(x, p, training, inplace)

    torch.nn.functional.dropout = dropout

    def randint(*a, **k):
        if fq is not None:
            assert fq, "FPS start queue exhausted"
            v = fq.pop(0)
            size = a[2] if len(a) > 2 else k.get("size")
            assert tuple(v.shape) == tuple(size), (v.shape, size)
            return v.clone()
        return o_randint(*a, **k)

    def randn_like(x, **k):
        if nq is not None:
            assert nq, "noise queue exhausted"
            v = nq.pop(0)
            assert v.shape == x.shape
            return v.clone()
        return o_randn_like(x, **k)

    torch.randint, torch.randn_like = randint, randn_like
    try:
        yield
    finally:
        torch.randint, torch.randn_like = o_randint, o_randn_like
        torch.nn.functional.dropout = o_dropout
