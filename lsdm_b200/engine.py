"""Thin PyTorch-side owner of one ``lsdm_handle``: device memory, streams and pointers only.

All arithmetic of the path happens inside ``liblsdm_b200.so``; this class allocates the workspace
with ``torch.empty``, hands ``data_ptr()``s and the current CUDA stream to the C ABI, and keeps
tensors alive while kernels that read them are in flight.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

N_POINTS, N_OBJ, CLIP_DIM = 1024, 9, 512


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on_device(fn):
    """Engine methods launch kernels on the handle's device: make it current for the call (the C ABI launches on the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)

    return wrapped


class Engine:
    def __init__(self, batch_local, n_cats=13, device=None, batch_global=None, batch_offset=0):
        if not torch.cuda.is_available():
            raise _lib.LsdmError(_lib.ESTATE, "lsdm_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_cats = int(n_cats)
        self.batch_local = int(batch_local)
        self.batch_global = int(batch_global if batch_global is not None else batch_local)
        self.batch_offset = int(batch_offset)
        cfg = _lib.Config(self.batch_local, self.batch_global, self.batch_offset, self.n_cats, self.device.index, 0)
        h = C.c_void_p()
        _lib.check(self.lib.lsdm_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self._ws = None
        self._keep = []
        self.T = 0
        self._alloc_workspace()
        # default: tensor cores (TF32 condition encoder with the fused SA / FP kernels, 3xTF32 x0 network); LSDM_PRECISION=fp32
        # selects the CUDA-core fp32 build of every dense layer (the parity tests run both).  The remaining switches exist for the
        # transparency legs of bench.py and the on/off equality tests; they all default to the fast path.
        self.set_precision(os.environ.get("LSDM_PRECISION", "tf32"))
        self.set_option("sa_fused", int(os.environ.get("LSDM_SA_FUSED", "1")))
        self.set_option("dedup_absent", int(os.environ.get("LSDM_DEDUP_ABSENT", "1")))
        self.set_option("x0_fused", int(os.environ.get("LSDM_X0_FUSED", "1")))
        self.set_option("sa1_compact", int(os.environ.get("LSDM_SA1_COMPACT", "1")))
        self.set_option("select_grid", int(os.environ.get("LSDM_SELECT_GRID", "9")))
        self.set_option("cond_stream", int(os.environ.get("LSDM_COND_STREAM", "1")))
        self.set_option("loop_invariants", int(os.environ.get("LSDM_LOOP_INVARIANTS", "15")))
        self.set_option("time_batch", int(os.environ.get("LSDM_TIME_BATCH", "1")))
        self.set_option("fps_compact", int(os.environ.get("LSDM_FPS_COMPACT", "1")))

    # ------------------------------------------------------------------ lifecycle
    @_on_device
    def _alloc_workspace(self):
        n = self.lib.lsdm_workspace_bytes(self.h)
        self._ws = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
        base = self._ws.data_ptr()
        aligned = (base + 255) & ~255
        _lib.check(self.lib.lsdm_set_workspace(self.h, C.c_void_p(aligned), C.c_size_t(n)))
        self.workspace_bytes = n

    @_on_device
    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.lsdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_batch(self, batch_local, batch_global=None, batch_offset=0):
        batch_global = batch_local if batch_global is None else batch_global
        realloc = batch_local != self.batch_local
        _lib.check(self.lib.lsdm_set_batch(self.h, batch_local, batch_global, batch_offset))
        self.batch_local, self.batch_global, self.batch_offset = batch_local, batch_global, batch_offset
        if realloc:
            torch.cuda.synchronize(self.device)
            self._alloc_workspace()

    def expected_keys(self):
        return [self.lib.lsdm_weight_key(self.h, i).decode() for i in range(self.lib.lsdm_num_weights(self.h))]

    @_on_device
    def load_state_dict(self, sd):
        """Uploads a reference ``model_state_dict`` (SURVEY.md Appendix B keys); clip_model.* is skipped."""
        st = _stream(self.device)
        for key in self.expected_keys():
            if key not in sd:
                raise KeyError(f"missing state-dict key {key}")
            t = sd[key]
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            if t.is_floating_point():
                t = t.detach().to(dtype=torch.float32).contiguous()
                self._keep.append(t)
                data = _ptr(t)
            else:
                data = None  # num_batches_tracked: accepted, ignored
            _lib.check(self.lib.lsdm_load_weight(self.h, key.encode(), data, shape, t.dim(), st))
        _lib.check(self.lib.lsdm_finalize_weights(self.h, st))
        torch.cuda.current_stream(self.device).synchronize()
        self._keep.clear()

    @_on_device
    def set_schedule(self, tables):
        """tables: dict of float64 numpy arrays (GaussianDiffusion attributes); cast to fp32 here, once."""
        names = ("posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped",
                 "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")
        arrs = [np.ascontiguousarray(np.asarray(tables[n], dtype=np.float64).astype(np.float32)) for n in names]
        T = len(arrs[0])
        ptrs = [C.c_void_p(a.ctypes.data) for a in arrs]
        _lib.check(self.lib.lsdm_set_schedule(self.h, *ptrs, T, _stream(self.device)))
        torch.cuda.current_stream(self.device).synchronize()
        self.T = T

    # ------------------------------------------------------------------ path
    def _f32(self, t, non_blocking=False):
        t = t.detach()
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            # non_blocking: a pinned host source is only enqueued; the caller waits for the copy before handing control back
            t = t.to(device=self.device, dtype=torch.float32, non_blocking=non_blocking and t.device.type == "cpu" and t.is_pinned()).contiguous()
        return t

    def _i64(self, t):
        t = t.detach()
        if t.dtype != torch.int64 or not t.is_contiguous():
            t = t.to(torch.int64).contiguous()
        return t  # host or device: the C side copies with cudaMemcpyDefault

    @_on_device
    def encode_conditions(self, text_emb, given_objs, given_cats, mask_global, fps_start):
        B = self.batch_local
        text_emb, given_objs, given_cats, mask_global = map(self._f32, (text_emb, given_objs, given_cats, mask_global))
        assert text_emb.shape == (B, CLIP_DIM), text_emb.shape
        assert given_objs.shape == (B, N_OBJ, N_POINTS, 3), given_objs.shape
        assert given_cats.shape == (B, N_OBJ, self.n_cats), given_cats.shape
        assert mask_global.shape == (self.batch_global, N_OBJ), mask_global.shape
        fps_start = self._i64(fps_start)
        assert fps_start.shape == (4, B * N_OBJ), fps_start.shape
        self._cond_keep = (text_emb, given_objs, given_cats, mask_global, fps_start)
        _lib.check(self.lib.lsdm_encode_conditions(self.h, _ptr(text_emb), _ptr(given_objs), _ptr(given_cats),
                                                   _ptr(mask_global), _ptr(fps_start), _stream(self.device)))
        if not fps_start.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()  # pageable host source must outlive the copy

    @_on_device
    def encode_conditions_train(self, text_emb, given_objs, given_cats, mask_global, fps_start, drop_mask):
        """model.train() condition encoder: BatchNorm batch statistics (+ running-stat update inside the handle) and the Dropout mask
        ``drop_mask[9B,128,1024]`` (0 or 2) of the backbone head."""
        B = self.batch_local
        text_emb, given_objs, given_cats, mask_global, drop_mask = map(self._f32, (text_emb, given_objs, given_cats, mask_global, drop_mask))
        assert drop_mask.shape == (B * N_OBJ, 128, N_POINTS), drop_mask.shape
        fps_start = self._i64(fps_start)
        assert fps_start.shape == (4, B * N_OBJ), fps_start.shape
        self._cond_keep = (text_emb, given_objs, given_cats, mask_global, fps_start, drop_mask)
        _lib.check(self.lib.lsdm_encode_conditions_train(self.h, _ptr(text_emb), _ptr(given_objs), _ptr(given_cats), _ptr(mask_global),
                                                         _ptr(fps_start), _ptr(drop_mask), _stream(self.device)))
        if not fps_start.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()

    def set_allreduce(self, process_group=None, enabled=True):
        """SyncBN for sharded train-mode forwards: BatchNorm (sum, sum-of-squares) buffers are all-reduced over ``process_group``
        (torch.distributed, NCCL) between the statistics and the normalisation kernels of every layer."""
        if not enabled:
            _lib.check(self.lib.lsdm_set_allreduce(self.h, None, None))
            self._allreduce_cb = None
            return
        import torch.distributed as dist

        def hook(_ctx, ptr, n):
            try:
                buf = None
                for owner in (self._ws, getattr(self, "_tape", None)):   # the statistics live in the workspace or in the training tape
                    if owner is not None and owner.data_ptr() <= int(ptr) and int(ptr) + 8 * n <= owner.data_ptr() + owner.numel():
                        off = int(ptr) - owner.data_ptr()
                        buf = owner[off:off + 8 * n].view(torch.float64)
                if buf is None:
                    raise RuntimeError("all-reduce hook: pointer outside the engine's buffers")
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=process_group)
                return 0
            except Exception:  # pragma: no cover
                import traceback
                traceback.print_exc()
                return 1

        self._allreduce_cb = _lib.ALLREDUCE_FN(hook)  # keep alive
        _lib.check(self.lib.lsdm_set_allreduce(self.h, C.cast(self._allreduce_cb, C.c_void_p), None))

    @_on_device
    def read_weight(self, key, like):
        """Current value of a state-dict entry inside the handle (BatchNorm running statistics after a train-mode forward)."""
        out = torch.empty(like.shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.lsdm_read_weight(self.h, key.encode(), _ptr(out), out.numel(), _stream(self.device)))
        return out

    @_on_device
    def denoise_step(self, x, t, noise, sample_out=None, want_x0=True, want_guiding=True, clip_denoised=False):
        """x is mutated in place (x += pcd_out).  Returns (sample, x0, guiding)."""
        B = self.batch_local
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape == (B, N_POINTS, 3)
        noise = self._f32(noise)
        t = self._i64(t)
        sample = torch.empty_like(x) if sample_out is None else sample_out
        x0 = torch.empty_like(x) if want_x0 else None
        gd = torch.empty_like(x) if want_guiding else None
        _lib.check(self.lib.lsdm_denoise_step(self.h, _ptr(x), _ptr(t), _ptr(noise), _ptr(sample), _ptr(x0), _ptr(gd),
                                              1 if clip_denoised else 0, _stream(self.device)))
        if not t.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()
        return sample, x0, gd

    @_on_device
    def forward(self, x, t):
        """x is mutated in place.  Returns (out_cat[B,C], x0, guiding)."""
        B = self.batch_local
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape == (B, N_POINTS, 3)
        t = self._i64(t)
        out_cat = torch.empty(B, self.n_cats, device=self.device)
        x0 = torch.empty_like(x)
        gd = torch.empty_like(x)
        _lib.check(self.lib.lsdm_forward(self.h, _ptr(x), _ptr(t), _ptr(out_cat), _ptr(x0), _ptr(gd), _stream(self.device)))
        if not t.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()
        return out_cat, x0, gd

    @_on_device
    def sample_loop(self, x, text_emb, given_objs, given_cats, mask_global, fps_start_all, noise_all, t_first, hoisted=False,
                    clip_denoised=False):
        """Runs len(noise_all) consecutive steps t_first, t_first-1, ... in place on x.  Returns (x0, guiding) of the last."""
        B = self.batch_local
        text_emb, given_objs, given_cats, mask_global, noise_all = map(
            self._f32, (text_emb, given_objs, given_cats, mask_global, noise_all))
        n = noise_all.shape[0]
        fps_start_all = self._i64(fps_start_all).to(self.device)
        assert fps_start_all.shape[1:] == (4, B * N_OBJ) and fps_start_all.shape[0] >= (1 if hoisted else n)
        assert noise_all.shape == (n, B, N_POINTS, 3)
        x0 = torch.empty_like(x)
        gd = torch.empty_like(x)
        self._cond_keep = (text_emb, given_objs, given_cats, mask_global, fps_start_all, noise_all)
        _lib.check(self.lib.lsdm_sample_loop(self.h, _ptr(x), _ptr(text_emb), _ptr(given_objs), _ptr(given_cats),
                                             _ptr(mask_global), _ptr(fps_start_all), _ptr(noise_all), int(t_first), int(n),
                                             1 if hoisted else 0, 1 if clip_denoised else 0, _ptr(x0), _ptr(gd),
                                             _stream(self.device)))
        return x0, gd

    @_on_device
    def out_cat(self):
        o = torch.empty(self.batch_local, self.n_cats, device=self.device)
        _lib.check(self.lib.lsdm_get_out_cat(self.h, _ptr(o), _stream(self.device)))
        return o

    @_on_device
    def pcd_out(self):
        o = torch.empty(self.batch_local, N_POINTS, 3, device=self.device)
        _lib.check(self.lib.lsdm_get_pcd_out(self.h, _ptr(o), _stream(self.device)))
        return o

    @_on_device
    def q_sample(self, x_start, t, noise):
        x_start, noise = self._f32(x_start), self._f32(noise)
        t = self._i64(t)
        out = torch.empty_like(x_start)
        _lib.check(self.lib.lsdm_q_sample(self.h, _ptr(x_start), _ptr(t), _ptr(noise), _ptr(out), _stream(self.device)))
        if not t.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()
        return out

    @_on_device
    def chamfer(self, x, y):
        """pytorch3d.loss.chamfer_distance(x, y)[0] with default arguments (scalar tensor)."""
        x, y = self._f32(x), self._f32(y)
        B, n, _ = x.shape
        m = y.shape[1]
        sums = torch.zeros(2, device=self.device)
        _lib.check(self.lib.lsdm_chamfer(self.h, _ptr(x), _ptr(y), B, n, m, _ptr(sums), _stream(self.device)))
        return sums.sum() / B

    @_on_device
    def cat_loss(self, probs, target_cat):
        probs, target_cat = self._f32(probs), self._f32(target_cat)
        B = probs.shape[0]
        s = torch.zeros(1, device=self.device)
        _lib.check(self.lib.lsdm_cat_loss(self.h, _ptr(probs), _ptr(target_cat), B, _ptr(s), _stream(self.device)))
        return s[0] / B

    @_on_device
    # ------------------------------------------------------------------ training backward
    def weight_slots(self):
        """{state-dict key: (offset, numel)} of the flat gradient buffer ``lsdm_training_backward`` accumulates into."""
        if getattr(self, "_slots", None) is None:
            slots = {}
            off, num = C.c_int64(), C.c_int64()
            for i in range(self.lib.lsdm_num_weights(self.h)):
                _lib.check(self.lib.lsdm_weight_slot(self.h, i, C.byref(off), C.byref(num)))
                if off.value >= 0:
                    slots[self.lib.lsdm_weight_key(self.h, i).decode()] = (off.value, num.value)
            self._slots = slots
        return self._slots

    @_on_device
    def training_backward(self, x_start, t, noise, text_emb, given_objs, given_cats, mask_global, target_cat, fps_start, drop_mask,
                          lambda_cat, g_mse, g_cat, want_x0=False):
        """Taped train-mode forward + reverse sweep inside the library.  Returns (flat gradient [lsdm_grad_floats], losses[2], x0)."""
        B = self.batch_local
        x_start, noise, text_emb, given_objs, given_cats, mask_global, target_cat, drop_mask = map(
            self._f32, (x_start, noise, text_emb, given_objs, given_cats, mask_global, target_cat, drop_mask))
        t = self._i64(t).to(self.device)
        fps_start = self._i64(fps_start)
        assert fps_start.shape == (4, B * N_OBJ) and drop_mask.shape == (B * N_OBJ, 128, N_POINTS)
        need = int(self.lib.lsdm_train_tape_bytes(self.h))
        if getattr(self, "_tape", None) is None or self._tape.numel() < need + 256:
            self._tape = None
            self._tape = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = (self._tape.data_ptr() + 255) & ~255
        grads = torch.zeros(int(self.lib.lsdm_grad_floats(self.h)), device=self.device)
        losses = torch.zeros(2, device=self.device)
        x0 = torch.empty(B, N_POINTS, 3, device=self.device) if want_x0 else None
        self._cond_keep = (x_start, noise, text_emb, given_objs, given_cats, mask_global, target_cat, drop_mask, t, fps_start)
        _lib.check(self.lib.lsdm_training_backward(self.h, _ptr(x_start), _ptr(t), _ptr(noise), _ptr(text_emb), _ptr(given_objs), _ptr(given_cats),
                                                   _ptr(mask_global), _ptr(target_cat), _ptr(fps_start), _ptr(drop_mask), float(lambda_cat),
                                                   float(g_mse), float(g_cat), C.c_void_p(base), C.c_size_t(need), _ptr(grads), _ptr(losses),
                                                   _ptr(x0), _stream(self.device)))
        if not fps_start.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()
        return grads, losses, x0

    def debug_tensor(self, name, dtype=torch.float32):
        n = self.lib.lsdm_debug_tensor(self.h, name.encode(), None, 0, None)
        if n < 0:
            _lib.check(int(n))
        out = torch.empty(int(n), dtype=dtype, device=self.device)
        r = self.lib.lsdm_debug_tensor(self.h, name.encode(), _ptr(out), out.numel() * out.element_size(), _stream(self.device))
        if r < 0:
            _lib.check(int(r))
        return out

    def launch_count(self):
        return int(self.lib.lsdm_launch_count(self.h))

    KCLASSES = ("gemm", "fps", "ball_query", "sa_gather", "three_nn", "fp_combine", "head3", "cond", "scene", "denoise", "other")

    def profile_begin(self):
        _lib.check(self.lib.lsdm_profile_begin(self.h))

    @_on_device
    def profile_end(self):
        """Returns ({class: ms}, {class: launches}, gemm_flops) measured with CUDA events inside the library."""
        n = len(self.KCLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        fl = C.c_double()
        _lib.check(self.lib.lsdm_profile_end(self.h, ms, cnt, n, C.byref(fl)))
        return dict(zip(self.KCLASSES, ms)), dict(zip(self.KCLASSES, cnt)), fl.value

    def profile_report(self):
        """Per-kernel rows of the tensor-core class from the last profiled pass: [{tag, launches, ms, flops}]."""
        rows = []
        for line in self.lib.lsdm_profile_report(self.h).decode().splitlines():
            tag, n, ms, fl = line.split("\t")
            rows.append({"tag": tag, "launches": int(n), "ms": float(ms), "flops": float(fl)})
        return rows

    PRECISIONS = {"fp32": (0, 0), "tf32": (1, 2), "tf32-all": (1, 1), "3xtf32": (2, 2)}

    def set_precision(self, mode):
        """'fp32': CUDA-core fp32 everywhere.  'tf32' (tensor cores): TF32 for the condition encoder, 3xTF32 for the
        per-step x0 network.  'tf32-all': TF32 everywhere.  '3xtf32': split-TF32 everywhere (~fp32 accuracy)."""
        enc, step = self.PRECISIONS[mode]
        _lib.check(self.lib.lsdm_set_precision(self.h, enc, step))
        self.precision = mode

    def set_option(self, name, value):
        _lib.check(self.lib.lsdm_set_option(self.h, name.encode(), int(value)))

    @_on_device
    def debug_gemm(self, A, W, bias=None, act=0, group_max=False, tf32=False, bias_mode=1, precision=None):
        """One linear layer through the library's GEMM (test hook)."""
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty(M // 32 if group_max else M, N, device=self.device)
        _lib.check(self.lib.lsdm_debug_gemm(self.h, _ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(out), out.stride(0), _ptr(bias),
                                            bias_mode, M, N, K, act, 1 if group_max else 0,
                                            (1 if tf32 else 0) if precision is None else precision, _stream(self.device)))
        return out
