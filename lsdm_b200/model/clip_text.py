"""Frozen CLIP text tower on the device -- stands where the reference keeps ``self.clip_model`` (``model/sdm.py:229-233,
245-277``): ``encode_text(tokens) -> [B, 512]`` float32, every operation a CUDA kernel of ``liblsdm_b200.so``
(``csrc/clip_text.cu``).  Tokenisation (``clip.tokenize``, host-side BPE with a vocabulary file that is not available
offline) is the caller's: pass its ``[B, 77]`` integer output.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..engine import _ptr, _stream

PREFIX = "clip_model."


class ClipTextTower:
    def __init__(self, device=None, precision="3xtf32"):
        if not torch.cuda.is_available():
            raise _lib.LsdmError(_lib.ESTATE, "lsdm_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _lib.check(self.lib.lsdm_clip_create(C.byref(h), self.device.index))
        self.h = h
        self._ws = None
        self.dims = None
        self.set_precision(precision)

    def set_precision(self, precision):
        code = {"fp32": 0, "tf32": 1, "3xtf32": 2}[precision]
        _lib.check(self.lib.lsdm_clip_set_precision(self.h, code))
        self.precision = precision

    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.lsdm_clip_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, state_dict, prefix=PREFIX):
        """Loads the text-side tensors of an openai/CLIP state dict (keys ``<prefix>token_embedding.weight`` ...; the visual
        tower, ``logit_scale`` and anything else are ignored; fp16 checkpoints are widened to fp32)."""
        keep = ("token_embedding.", "positional_embedding", "transformer.resblocks.", "ln_final.", "text_projection")
        n = 0
        with torch.cuda.device(self.device):
            for k, v in state_dict.items():
                if not k.startswith(prefix):
                    continue
                name = k[len(prefix):]
                if not name.startswith(keep):
                    continue
                t = v.detach().to(dtype=torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(self.lib.lsdm_clip_load_weight(self.h, name.encode(), _ptr(t), shape, t.dim(), _stream(self.device)))
                torch.cuda.current_stream(self.device).synchronize()  # `t` may be a temporary
                n += 1
            _lib.check(self.lib.lsdm_clip_finalize(self.h))
        d = [C.c_int32() for _ in range(6)]
        _lib.check(self.lib.lsdm_clip_dims(self.h, *[C.byref(x) for x in d]))
        self.dims = dict(zip(("width", "layers", "heads", "ctx", "vocab", "embed"), (int(x.value) for x in d)))
        return n

    def encode_text(self, tokens, seq_len=None):
        """``clip_model.encode_text(tokens).float()``.  ``tokens``: integer ``[B, ctx]``.  ``seq_len`` (default: the largest
        EOT position + 1) bounds the computed positions -- bit-identical to the full context because attention is causal."""
        if self.dims is None:
            raise _lib.LsdmError(_lib.ESTATE, "CLIP text tower has no weights (load_state_dict)")
        tokens = torch.as_tensor(tokens)
        if tokens.dim() != 2 or tokens.shape[1] != self.dims["ctx"] or tokens.is_floating_point():
            raise ValueError(f"tokens must be integer [B, {self.dims['ctx']}], got {tuple(tokens.shape)} {tokens.dtype}")
        if seq_len is None:
            seq_len = int(tokens.argmax(dim=-1).max()) + 1
        B = tokens.shape[0]
        tok = tokens.to(device=self.device, dtype=torch.int32).contiguous()
        need = self.lib.lsdm_clip_workspace_bytes(self.h, B, seq_len)
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 255) & ~255
        out = torch.empty(B, self.dims["embed"], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.lsdm_clip_encode_text(self.h, _ptr(tok), B, seq_len, C.c_void_p(base), C.c_size_t(need), _ptr(out),
                                                      _stream(self.device)))
        self._keep = tok
        return out

    def launch_count(self):
        return int(self.lib.lsdm_clip_launch_count(self.h))
