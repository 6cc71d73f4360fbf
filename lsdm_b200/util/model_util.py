"""Factory -- mirror of the reference's ``util/model_util.py`` (same names, arguments and defaults), so that
``run/test_sdm.py`` / ``run/train_sdm.py`` work unchanged once their import points here."""
from __future__ import annotations

from ..diffusion import gaussian_diffusion as gd
from ..diffusion.respace import SpacedDiffusion, space_timesteps
from ..model.sdm import SceneDiffusionModel


def load_model_wo_clip(model, state_dict):
    """Reference util/model_util.py:10-13."""
    missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
    assert len(unexpected_keys) == 0
    assert all([k.startswith('clip_model.') for k in missing_keys])


def get_default_model_proxd():
    """Reference util/model_util.py:26-48."""
    return {'seq_len': 256, 'modality': 'text', 'clip_version': 'ViT-B/32', 'clip_dim': 512, 'dropout': 0.1, 'n_layer': 6,
            'n_head': 8, 'f_vert': 64, 'dim_ff': 512, 'd_hid': 256, 'mesh_ds_dir': "data/mesh_ds", 'posa_path': None,
            'latent_dim': 128, 'pcd_dim': 3, 'cond_mask_prob': 1.0, 'device': 0, 'vert_dims': 655, 'obj_cat': 8,
            'data_rep': 'rot6d', 'njoints': 251}


def get_default_model_humanise():
    """Reference util/model_util.py:50-73."""
    d = get_default_model_proxd()
    d['max_cats'] = 11
    return d


def get_default_diffusion():
    """Reference util/model_util.py:76-85."""
    return {"lambda_fc": 0.0, "lambda_rcxyz": 0.0, "lambda_vel": 0.0, "lambda_cat": 0.1, "noise_schedule": "cosine",
            "sigma_small": True}


def create_gaussian_diffusion(args, timestep_respacing=''):
    """Reference util/model_util.py:127-163: 1000 cosine steps, x0 prediction, fixed small variance, MSE, no respacing.
    ``timestep_respacing`` (hard-coded '' in the reference) is exposed for BASELINE config 3 ('ddim100')."""
    steps = 1000
    betas = gd.get_named_beta_schedule(args['noise_schedule'], steps, 1.)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing),
        betas=betas,
        model_mean_type=gd.ModelMeanType.START_X,
        model_var_type=gd.ModelVarType.FIXED_SMALL if args['sigma_small'] else gd.ModelVarType.FIXED_LARGE,
        loss_type=gd.LossType.MSE,
        rescale_timesteps=False,
        lambda_vel=args['lambda_vel'],
        lambda_rcxyz=args['lambda_rcxyz'],
        lambda_fc=args['lambda_fc'],
        lambda_cat=args['lambda_cat'],
    )


def create_model_and_diffusion(datatype):
    """Reference util/model_util.py:16-23."""
    kw = get_default_model_proxd() if datatype == "proxd" else get_default_model_humanise()
    model = SceneDiffusionModel(**kw)
    diffusion = create_gaussian_diffusion(get_default_diffusion())
    return model, diffusion
