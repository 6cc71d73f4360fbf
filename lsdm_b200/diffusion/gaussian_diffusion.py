"""``GaussianDiffusion`` -- mirror of the LIVE subset of the reference's
``diffusion/gaussian_diffusion.py`` (schedules :22-66, tables :122-204, ``q_sample`` :238,
``q_posterior_mean_variance`` :258, ``p_mean_variance`` :282, ``p_sample`` :501,
``p_sample_loop[_progressive]`` :611/:684, ``training_losses`` :1256).

Host side keeps only what the reference keeps on the host: the float64 schedule tables (numpy) and the
loop control.  The per-step arithmetic -- model forward, ``x += pcd_out``, posterior mean, ancestral noise --
is ONE call into ``liblsdm_b200.so`` (``lsdm_denoise_step`` / ``lsdm_sample_loop``).  RNG is consumed like the
reference does: four CPU ``torch.randint`` FPS-start draws per model call, then ``torch.randn_like(x)``.

Unsupported reference options raise ``NotImplementedError`` instead of silently doing something else:
EPSILON / PREVIOUS_X mean types, learned variances, KL losses, ``cond_fn``, ``denoised_fn`` (no reference
caller uses them: util/model_util.py:127-163, run/test_sdm.py:156-176).
"""
from __future__ import annotations

import enum
import math
from copy import deepcopy

import numpy as np
import torch as th


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.):
    """Reference gaussian_diffusion.py:22-47."""
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """Reference gaussian_diffusion.py:50-66."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """Reference gaussian_diffusion.py:1585-1598 (float64 gather, then fp32 cast).  API-compat helper only:
    the kernels read fp32 copies of the tables uploaded once by ``lsdm_set_schedule``."""
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


class GaussianDiffusion:
    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False, lambda_rcxyz=0.,
                 lambda_vel=0., lambda_pose=1., lambda_orient=1., lambda_loc=1., data_rep='rot6d', lambda_root_vel=0.,
                 lambda_vel_rcxyz=0., lambda_fc=0., lambda_cat=0.05):
        if model_mean_type != ModelMeanType.START_X:
            raise NotImplementedError("only START_X (predict_xstart=True, util/model_util.py:129) is on the accelerated path")
        if model_var_type not in (ModelVarType.FIXED_SMALL, ModelVarType.FIXED_LARGE):
            raise NotImplementedError("learned variances are not used by SDM (util/model_util.py:133)")
        if loss_type not in (LossType.MSE, LossType.RESCALED_MSE):
            raise NotImplementedError("KL loss types are not used by SDM (util/model_util.py:137)")
        if rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=False in the reference factory (util/model_util.py:134)")
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        self.rescale_timesteps, self.data_rep = rescale_timesteps, data_rep
        if data_rep != 'rot_vel' and lambda_pose != 1.:
            raise ValueError('lambda_pose is relevant only when training on velocities!')
        self.lambda_pose, self.lambda_orient, self.lambda_loc = lambda_pose, lambda_orient, lambda_loc
        self.lambda_rcxyz, self.lambda_vel, self.lambda_root_vel = lambda_rcxyz, lambda_vel, lambda_root_vel
        self.lambda_vel_rcxyz, self.lambda_fc, self.lambda_cat = lambda_vel_rcxyz, lambda_fc, lambda_cat
        if max(lambda_rcxyz, lambda_vel, lambda_root_vel, lambda_vel_rcxyz, lambda_fc) > 0.:
            raise NotImplementedError("geometric losses are zero in the SDM configuration (util/model_util.py:76-85)")

        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert betas.ndim == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    # ------------------------------------------------------------------ engine plumbing
    def _tables_for_engine(self):
        if self.model_var_type == ModelVarType.FIXED_LARGE:  # gaussian_diffusion.py:338-343
            logvar = np.log(np.append(self.posterior_variance[1], self.betas[1:]))
        else:
            logvar = self.posterior_log_variance_clipped
        return {"posterior_mean_coef1": self.posterior_mean_coef1, "posterior_mean_coef2": self.posterior_mean_coef2,
                "posterior_log_variance_clipped": logvar, "sqrt_alphas_cumprod": self.sqrt_alphas_cumprod,
                "sqrt_one_minus_alphas_cumprod": self.sqrt_one_minus_alphas_cumprod}

    @staticmethod
    def _unwrap(model):
        return getattr(model, "model", model)

    def _engine(self, model, batch, device):
        eng = self._unwrap(model).engine(batch, device)
        if getattr(eng, "_sched_owner", None) is not self:
            eng.set_schedule(self._tables_for_engine())
            eng._sched_owner = self
        return eng

    def _variance_tables(self):
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            v = np.append(self.posterior_variance[1], self.betas[1:])
            return v, np.log(v)
        return self.posterior_variance, self.posterior_log_variance_clipped

    @staticmethod
    def _check_x(x):
        if not (x.is_cuda and x.dtype == th.float32 and x.is_contiguous()):
            raise ValueError("x must be a contiguous float32 CUDA tensor (the model updates it in place)")

    # ------------------------------------------------------------------ q
    def q_sample(self, x_start, t, noise=None, model=None):
        """Reference gaussian_diffusion.py:238-256.  With ``model`` the fused CUDA kernel is used."""
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        if model is not None and x_start.is_cuda:
            return self._engine(model, x_start.shape[0], x_start.device).q_sample(x_start, t, noise)
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """Reference gaussian_diffusion.py:258-280 (API-compat helper; the sampling path fuses this into the kernel)."""
        assert x_start.shape == x_t.shape
        mean = (_extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + _extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        var = _extract_into_tensor(self.posterior_variance, t, x_t.shape)
        logvar = _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    # ------------------------------------------------------------------ p
    def _scale_timesteps(self, t):
        return t

    def _fused_step(self, model, x, mask, t, given_objs, given_cats, y, noise, clip_denoised, fps_start=None):
        self._check_x(x)
        net = self._unwrap(model)
        eng = self._engine(model, x.shape[0], x.device)
        net.encode(mask, given_objs, given_cats, y, fps_start, device=x.device)
        sample, x0, guiding = eng.denoise_step(x, t, noise, clip_denoised=clip_denoised)
        net.saved_cat = eng.out_cat().unsqueeze(1)
        net.saved_guiding_points = guiding
        return sample, x0

    def p_mean_variance(self, model, x, mask, t, given_objs, given_cats, y, clip_denoised=True, denoised_fn=None,
                        model_kwargs=None):
        """Reference gaussian_diffusion.py:282-393.  ``x`` is mutated by the model; the mutated tensor is the x_t of the mean."""
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is not used by any SDM caller")
        B = x.shape[0]
        assert t.shape == (B,)
        mean, x0 = self._fused_step(model, x, mask, t, given_objs, given_cats, y, th.zeros_like(x), clip_denoised)
        var, logvar = self._variance_tables()
        return {"mean": mean, "variance": _extract_into_tensor(var, t, x.shape),
                "log_variance": _extract_into_tensor(logvar, t, x.shape), "pred_xstart": x0}

    def p_sample(self, model, x, mask, t, given_objs, given_cats, y, clip_denoised=True, denoised_fn=None, cond_fn=None,
                 model_kwargs=None, const_noise=False):
        """Reference gaussian_diffusion.py:501-561: one fused CUDA step."""
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn guidance is not used by any SDM caller")
        self._check_x(x)
        net = self._unwrap(model)
        fps_start = net.draw_fps_starts(x.shape[0])  # model-internal draws come first (pointnet2_utils.py:72) ...
        noise = th.randn_like(x)                      # ... then the sampling noise (gaussian_diffusion.py:545)
        if const_noise:
            noise = noise[[0]].repeat(x.shape[0], 1, 1)
        sample, x0 = self._fused_step(model, x, mask, t, given_objs, given_cats, y, noise, clip_denoised, fps_start)
        return {"sample": sample, "pred_xstart": x0}

    def p_sample_loop(self, model, shape, mask, given_objs, given_cats, y, noise=None, clip_denoised=True, denoised_fn=None,
                      cond_fn=None, model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """Reference gaussian_diffusion.py:611-682.  The call every reference driver makes (run/test_sdm.py:166-182,
        run/train_sdm.py:132-148, run/scene_edit.py:296-313) -- ``dump_steps=None``, no progress bar -- runs the whole loop
        inside the library (``lsdm_sample_loop``: three-stream pipeline, no per-step Python); intermediate dumps and the
        tqdm bar need the per-step generator and take :meth:`p_sample_loop_progressive`.  Both give the same tensors for the
        same RNG state (tests/test_gpu_parity.py)."""
        if dump_steps is None and not progress and not const_noise:
            if denoised_fn is not None or cond_fn is not None or cond_fn_with_grad or randomize_class:
                raise NotImplementedError("denoised_fn / cond_fn guidance / randomize_class are not used by any SDM caller")
            return self._sample_loop_library(model, shape, mask, given_objs, given_cats, y, noise=noise, clip_denoised=clip_denoised,
                                             device=device, skip_timesteps=skip_timesteps, init_image=init_image)
        final = None
        dump = [] if dump_steps is not None else None
        for i, sample in enumerate(self.p_sample_loop_progressive(
                model, shape, mask, given_objs, given_cats, y, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress, skip_timesteps=skip_timesteps,
                init_image=init_image, randomize_class=randomize_class, cond_fn_with_grad=cond_fn_with_grad,
                const_noise=const_noise)):
            if dump is not None and i in dump_steps:
                dump.append(deepcopy(sample["sample"]))
            final = sample
        return dump if dump is not None else final["sample"]

    def p_sample_loop_progressive(self, model, shape, mask, given_objs, given_cats, y, noise=None, clip_denoised=True,
                                  denoised_fn=None, cond_fn=None, model_kwargs=None, device=None, progress=False,
                                  skip_timesteps=0, init_image=None, randomize_class=False, cond_fn_with_grad=False,
                                  const_noise=False):
        """Reference gaussian_diffusion.py:684-759."""
        if cond_fn is not None or cond_fn_with_grad or randomize_class:
            raise NotImplementedError("guidance / randomize_class are not used by any SDM caller")
        if device is None:
            device = next(self._unwrap(model).parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = th.zeros_like(img)
        indices = list(range(self.num_timesteps - skip_timesteps))[::-1]
        if init_image is not None:
            my_t = th.ones([shape[0]], device=device, dtype=th.long) * indices[0]
            img = self.q_sample(init_image, my_t, img)
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            with th.no_grad():
                out = self.p_sample(model, img, mask, t, given_objs, given_cats, y, clip_denoised=clip_denoised,
                                    denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs, const_noise=const_noise)
                yield out
                img = out["sample"]

    def _sample_loop_library(self, model, shape, mask, given_objs, given_cats, y, noise=None, clip_denoised=True, device=None,
                             skip_timesteps=0, init_image=None, hoisted=False, chunk=None):
        """The T-step ancestral loop of reference gaussian_diffusion.py:684-759 with the loop body inside ``lsdm_sample_loop``.
        RNG is consumed as the reference does: ``th.randn(*shape)`` for x_T, then per step four CPU-generator FPS-start draws and
        one device-generator ``randn_like``; per chunk of ``chunk`` steps they are drawn up front (the two generators are
        independent streams, so the per-generator order is the reference's) and uploaded once.  ``skip_timesteps`` /
        ``init_image`` follow :716-731: x_T = q_sample(init_image or zeros, t_first, noise)."""
        net = self._unwrap(model)
        if device is None:
            device = next(net.parameters()).device
        device = th.device(device)
        B = shape[0]
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        caller_owns_img = noise is not None
        n_total = self.num_timesteps - skip_timesteps
        if n_total <= 0:
            raise ValueError("skip_timesteps leaves no timestep to run")
        if skip_timesteps and init_image is None:
            init_image = th.zeros_like(img)
        eng = self._engine(model, B, device)
        with th.no_grad():
            if init_image is not None:
                my_t = th.ones([B], device=device, dtype=th.long) * (n_total - 1)
                # a fresh tensor: the caller's noise is not touched  (the engine is at hand: no second weight-signature walk)
                img = eng.q_sample(init_image, my_t, img) if img.is_cuda else self.q_sample(init_image, my_t, img, model=model)
                caller_owns_img = False
            self._check_x(img)
            # conditions go to the device once per loop (the reference's callers move them before the call)
            # (pinned host buffers are only enqueued; `uploaded` is waited for before this function returns)
            text = eng._f32(net._encode_text(y), non_blocking=True)
            mask_d, objs_d, cats_d = (eng._f32(v, non_blocking=True) for v in (mask, given_objs, given_cats))
            uploaded = th.cuda.Event()
            uploaded.record(th.cuda.current_stream(device))
            fps0 = net.draw_fps_starts(B)
            if caller_owns_img or hoisted or n_total == 1:
                # first step through encode + denoise_step: the caller's tensor (when `noise` is given) only sees the
                # first model call's in-place `x += pcd_out`, and the sample goes to a fresh tensor, as in the reference
                nz0 = th.randn_like(img)
                t0 = th.full((B,), n_total - 1, device=img.device, dtype=th.long)
                net.encode(mask_d, objs_d, cats_d, text, fps0, device=img.device)
                cur, x0, gd = eng.denoise_step(img, t0, nz0, clip_denoised=clip_denoised)
                done, fps0_used = 1, True
            else:
                # x_T is this function's own tensor: every step, the first included, runs inside the pipelined library loop
                cur, done, fps0_used = img, 0, False
            while done < n_total:
                # steps per library call: the host draws a call's FPS starts up front while the GPU still runs the
                # previous call, so only the first call's draws are exposed -- it is kept short; every call pays one pipeline
                # fill (~3 ms), so the later ones are long
                n = min(chunk if chunk else (50 if done < 50 else 250), n_total - done)
                # the device-generator draws are launched first so that the GPU produces them while the host draws the FPS starts
                # (two independent generators: each one's own order is the reference's)
                nz = th.empty(n, *img.shape, device=img.device)
                for k in range(n):
                    nz[k] = th.randn_like(img)
                if hoisted:
                    fps = fps0[None]
                else:
                    fps = net.draw_fps_starts_steps(B, n, first=None if fps0_used else fps0)
                fps0_used = True
                x0, gd = eng.sample_loop(cur, text, objs_d, cats_d, mask_d, fps, nz, n_total - 1 - done, hoisted, clip_denoised)
                done += n
        net.saved_cat = eng.out_cat().unsqueeze(1)
        net.saved_guiding_points = gd
        uploaded.synchronize()  # (long done: the library call waited for the cloud classification, which follows the uploads)
        return cur

    def p_sample_loop_fused(self, model, shape, mask, given_objs, given_cats, y, noise=None, clip_denoised=True, device=None,
                            skip_timesteps=0, init_image=None, hoisted=False, chunk=None):
        """:meth:`p_sample_loop`'s library loop with its two extra knobs exposed: ``hoisted=True`` encodes the conditions once
        with the first step's FPS starts (an algorithmic optimisation that changes the random draw the backbone sees,
        SURVEY.md 7.0 -- not the reference's per-step behaviour) and ``chunk`` (steps per library call; default: 50 for the first call, then 250)."""
        return self._sample_loop_library(model, shape, mask, given_objs, given_cats, y, noise=noise, clip_denoised=clip_denoised,
                                         device=device, skip_timesteps=skip_timesteps, init_image=init_image, hoisted=hoisted,
                                         chunk=chunk)

    def ddim_sample_loop(self, *args, **kwargs):
        """Referenced but never called by the reference (run/test_sdm.py:160-164); its own implementation raises
        TypeError with the SDM signature (gaussian_diffusion.py:761-784,908-926)."""
        raise NotImplementedError("ddim_sample_loop is dead in the reference (pre-SDM signature); use p_sample_loop with a "
                                  "respaced SpacedDiffusion")

    # ------------------------------------------------------------------ training
    def training_losses(self, model, cf, mask, t, given_objs, given_cats, target_cat, y=None, noise=None):
        """Reference gaussian_diffusion.py:1256-1342 (MSE loss type, START_X): ``{'cat_loss','mse','loss'}`` as scalar tensors.
        Under ``torch.no_grad()`` (or with frozen parameters) only the forward runs.  Otherwise the scalars carry an autograd
        node whose backward is ``lsdm_training_backward`` (taped fp32 forward + reverse sweep inside the library), so that the
        reference's ``loss.backward()`` / ``mp_trainer.backward(loss)`` (run/train_sdm.py:78-84) fills ``.grad`` of every
        trainable parameter."""
        net = self._unwrap(model)
        x_start = cf
        if noise is None:
            noise = th.randn_like(x_start)
        params = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        if th.is_grad_enabled() and params:
            if not net.training:
                # eval-mode forward value; a .backward() on it says what is missing instead of torch's opaque "does not require grad"
                with th.no_grad():
                    terms = self.training_losses(model, cf, mask, t, given_objs, given_cats, target_cat, y=y, noise=noise)
                return {k: _NoEvalBackward.apply(v, params[0][1]) for k, v in terms.items()}
            mse, ce = _TrainingLossFn.apply(self, model, x_start, mask, t, given_objs, given_cats, target_cat, y, noise,
                                            [n for n, _ in params], *[p for _, p in params])
            cat_loss = ce * self.lambda_cat
            return {"cat_loss": cat_loss, "mse": mse, "loss": mse + cat_loss}
        eng = self._engine(model, x_start.shape[0], x_start.device)
        x_t = eng.q_sample(x_start.float(), t, noise.float())
        out_cat, model_output = net(x_t, mask, self._scale_timesteps(t), given_objs, given_cats, y)
        terms = {}
        cat_loss = eng.cat_loss(out_cat.squeeze(1), target_cat) * self.lambda_cat
        terms["cat_loss"] = cat_loss
        assert model_output.shape == x_start.shape
        terms["mse"] = eng.chamfer(model_output.float(), x_start.float())
        terms["loss"] = terms["mse"] + cat_loss
        return terms


# attn_layer's value / output projections never reach the loss (only the attention WEIGHTS are used, model/sdm.py:182): the
# reference leaves their .grad at None
_DEAD_PARAMS = ("attn_layer.v_proj_weight", "attn_layer.out_proj.weight", "attn_layer.out_proj.bias")


class _NoEvalBackward(th.autograd.Function):
    """Carries an eval-mode loss value; its backward raises (the library's backward implements model.train(): BatchNorm batch
    statistics, as run/train_sdm.py trains)."""

    @staticmethod
    def forward(ctx, value, _param):
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("backward through training_losses is implemented for model.train() (BatchNorm batch statistics, as "
                                  "run/train_sdm.py trains), not for model.eval()")


class _TrainingLossFn(th.autograd.Function):
    """(chamfer, mean cross-entropy) of one training forward, with the library's backward."""

    @staticmethod
    def forward(ctx, diffusion, model, x_start, mask, t, given_objs, given_cats, target_cat, y, noise, names, *params):
        net = diffusion._unwrap(model)
        B = x_start.shape[0]
        eng = diffusion._engine(model, B, x_start.device)
        text = net._encode_text(y)
        # RNG exactly as the reference's forward consumes it: four CPU FPS-start draws, then the Dropout(0.5) mask of the head
        fps = net.draw_fps_starts(B)
        drop = net.draw_dropout_mask(B, eng.device)
        with th.no_grad():
            x_t = eng.q_sample(x_start.float(), t, noise.float())
            net.encode(mask, given_objs, given_cats, text, fps, device=eng.device, drop_mask=drop)   # updates the BatchNorm running stats
            out_cat, x0, guiding = eng.forward(x_t, diffusion._scale_timesteps(t))
            net.saved_cat, net.saved_guiding_points = out_cat.unsqueeze(1), guiding
            ce = eng.cat_loss(out_cat, target_cat)
            mse = eng.chamfer(x0, x_start.float())
        ctx.saved = (diffusion, model, x_start, mask, t, given_objs, given_cats, target_cat, text, noise, fps, drop, list(names),
                     [tuple(p.shape) for p in params])
        return mse, ce

    @staticmethod
    def backward(ctx, g_mse, g_ce):
        diffusion, model, x_start, mask, t, given_objs, given_cats, target_cat, text, noise, fps, drop, names, shapes = ctx.saved
        net = diffusion._unwrap(model)
        eng = diffusion._engine(model, x_start.shape[0], x_start.device)
        if net._shard is not None and net._sync_bn_group is not None:
            eng.set_allreduce(net._sync_bn_group if net._sync_bn_group is not True else None)
        flat, _, _ = eng.training_backward(x_start, t, noise, text, given_objs, given_cats, mask, target_cat, fps, drop, 1.0,
                                           float(g_mse), float(g_ce))
        slots = eng.weight_slots()
        grads = []
        for n, shp in zip(names, shapes):
            if n in _DEAD_PARAMS or n not in slots:
                grads.append(None)
                continue
            off, num = slots[n]
            grads.append(flat[off:off + num].view(shp))
        return (None,) * 11 + tuple(grads)
