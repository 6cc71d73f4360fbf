"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Runs only in the build container (needs /root/reference).  Inputs and weights
come from ``lsdm_b200.synthetic`` seeds, so fixtures hold reference OUTPUTS only.
Every array is produced by live reference code (``model/sdm.py``,
``diffusion/gaussian_diffusion.py``, ``diffusion/respace.py``, ``model/pcd_backbone/*``,
``posa/posa_models.py``) through ``ref_harness`` -- see that file for the stubs.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness as rh  # noqa: E402
from lsdm_b200 import synthetic as syn  # noqa: E402

warnings.filterwarnings("ignore")

# seeds of record (tests regenerate the same inputs from these)
SEED_W = 0
CASES = {}


def npf(t):
    return t.detach().cpu().numpy()


def case_forward(kind, B=3, seed_in=1, seed_rng=2):
    """One full forward at t=[999,500,0] with per-stage captures."""
    sd = syn.make_state_dict(SEED_W, kind)
    m = rh.build_reference_model(sd)
    inp = syn.make_inputs(seed_in, B)
    fps, _ = syn.make_step_randoms(seed_rng, B, 1)
    caps = {}
    hk = [
        m.pcd_backbone.register_forward_hook(lambda mod, i, o: caps.__setitem__("backbone", o)),
        m.human_backbone.register_forward_hook(lambda mod, i, o: caps.__setitem__("hm", o)),
        m.attn_layer.register_forward_hook(lambda mod, i, o: caps.__setitem__("attn_w", o[1][:, 0])),
        m.pcd_attention.register_forward_hook(lambda mod, i, o: caps.__setitem__("pa", o[0][:, 0])),
        m.combine_extraction.register_forward_hook(lambda mod, i, o: caps.__setitem__("emb", o)),
        m.point_wise_trans_layer.register_forward_hook(lambda mod, i, o: caps.__setitem__("pw", o)),
        m.translation_layer.register_forward_hook(lambda mod, i, o: caps.__setitem__("tr", o)),
        m.embed_text.register_forward_hook(lambda mod, i, o: caps.__setitem__("enc", o)),
    ]
    # record the discrete selections of the first two clouds (wrap, do not modify, the reference fns)
    ns = rh.load_reference()
    sel = {"fps": [], "ball": []}
    o_fps, o_ball = ns.pn2u.farthest_point_sample, ns.pn2u.query_ball_point

    def fps_w(xyz, npoint):
        r = o_fps(xyz, npoint)
        sel["fps"].append(r[:2].clone())
        return r

    def ball_w(radius, nsample, xyz, new_xyz):
        r = o_ball(radius, nsample, xyz, new_xyz)
        sel["ball"].append(r[:2].clone())
        return r

    ns.pn2u.farthest_point_sample, ns.pn2u.query_ball_point = fps_w, ball_w
    x = inp["x_T"].clone()
    t = torch.tensor([999, 500, 0][:B])
    try:
        with torch.no_grad(), rh.injected_rng(fps_starts=list(fps[0])):
            out_cat, x0 = m(x, inp["mask"], t, inp["given_objs"], inp["given_cats"], inp["text_emb"])
    finally:
        ns.pn2u.farthest_point_sample, ns.pn2u.query_ball_point = o_fps, o_ball
        for h in hk:
            h.remove()
    out = {
        "t": t.numpy(),
        "out_cat": npf(out_cat),
        "x0": npf(x0),
        "x_mutated": npf(x),
        "guiding": npf(m.saved_guiding_points),
        "backbone": npf(caps["backbone"]),
        "hm": npf(caps["hm"]),
        "attn_w": npf(caps["attn_w"]),
        "pa": npf(caps["pa"]),
        "tr": npf(caps["tr"]),
        "enc": npf(caps["enc"]),
        "emb_sub": npf(caps["emb"][:, ::8]),
        "pw_sub": npf(caps["pw"][:, :, ::4]),
    }
    for lvl in range(4):
        out[f"fps_idx{lvl}"] = sel["fps"][lvl].numpy().astype(np.int16)
        out[f"ball_idx{lvl}"] = sel["ball"][lvl].numpy().astype(np.int16)
    return out


def case_p_sample(kind, B=2, seed_in=3, seed_rng=4):
    """BASELINE config 1: one p_sample at t=999, B=2 (gaussian_diffusion.py:501-561)."""
    ns = rh.load_reference()
    sd = syn.make_state_dict(SEED_W, kind)
    m = rh.build_reference_model(sd)
    diff = ns.model_util.create_gaussian_diffusion(ns.model_util.get_default_diffusion())
    inp = syn.make_inputs(seed_in, B)
    fps, noise = syn.make_step_randoms(seed_rng, B, 1)
    x = inp["x_T"].clone()
    t = torch.full((B,), 999, dtype=torch.long)
    with torch.no_grad(), rh.injected_rng(fps_starts=list(fps[0]), noises=[noise[0]]):
        out = diff.p_sample(m, x, inp["mask"], t, inp["given_objs"], inp["given_cats"], inp["text_emb"], clip_denoised=False)
    return {"sample": npf(out["sample"]), "pred_xstart": npf(out["pred_xstart"]), "x_mutated": npf(x),
            "saved_cat": npf(m.saved_cat), "guiding": npf(m.saved_guiding_points)}


def case_loop(kind, B=2, sections="8", seed_in=5, seed_rng=6):
    """Respaced ancestral loop through SpacedDiffusion.p_sample_loop (respace.py:64-130,
    gaussian_diffusion.py:611-759); the model sees the compact ts (respace.py:130)."""
    ns = rh.load_reference()
    sd = syn.make_state_dict(SEED_W, kind)
    m = rh.build_reference_model(sd)
    betas = ns.gd.get_named_beta_schedule("cosine", 1000, 1.0)
    keep = ns.respace.space_timesteps(1000, sections)
    diff = ns.respace.SpacedDiffusion(
        use_timesteps=keep, betas=betas, model_mean_type=ns.gd.ModelMeanType.START_X,
        model_var_type=ns.gd.ModelVarType.FIXED_SMALL, loss_type=ns.gd.LossType.MSE, rescale_timesteps=False,
        lambda_vel=0.0, lambda_rcxyz=0.0, lambda_fc=0.0, lambda_cat=0.1)
    T = diff.num_timesteps
    inp = syn.make_inputs(seed_in, B)
    fps, noise = syn.make_step_randoms(seed_rng, B, T)
    x_T = inp["x_T"].clone()
    with torch.no_grad(), rh.injected_rng(fps_starts=[v for s in fps for v in s], noises=list(noise)):
        sample = diff.p_sample_loop(m, (B, 1024, 3), inp["mask"], inp["given_objs"], inp["given_cats"], inp["text_emb"],
                                    noise=x_T, clip_denoised=False)
    return {"sample": npf(sample), "T": np.int64(T), "keep": np.array(sorted(keep)),
            "saved_cat": npf(m.saved_cat), "guiding": npf(m.saved_guiding_points), "x_T_after": npf(x_T)}


def case_train(kind, B=4, seed_in=7, seed_rng=8):
    """training_losses forward in eval mode (gaussian_diffusion.py:1256-1342). The chamfer term goes
    through the harness's cdist stub of pytorch3d (absent package) -- see DESIGN.md."""
    ns = rh.load_reference()
    sd = syn.make_state_dict(SEED_W, kind)
    m = rh.build_reference_model(sd)
    diff = ns.model_util.create_gaussian_diffusion(ns.model_util.get_default_diffusion())
    inp = syn.make_inputs(seed_in, B, training=True)
    fps, noise = syn.make_step_randoms(seed_rng, B, 1)
    with torch.no_grad(), rh.injected_rng(fps_starts=list(fps[0])):
        terms = diff.training_losses(m, inp["x_start"].clone(), inp["mask"], inp["t"], inp["given_objs"], inp["given_cats"],
                                     inp["target_cat"], y=inp["text_emb"], noise=noise[0])
    return {k: np.float64(float(terms[k])) for k in ("cat_loss", "mse", "loss")}


def case_train_forward(kind="wellcond", B=3, seed_in=13, seed_rng=14, seed_drop=15):
    """model.train() forward: BatchNorm batch statistics + running-stat updates, Dropout(0.5) in the backbone head with an injected
    mask (model/pcd_backbone/pointnet2.py:76), and training_losses on top (gaussian_diffusion.py:1256-1342)."""
    ns = rh.load_reference()
    sd = syn.make_state_dict(SEED_W, kind)
    m = rh.build_reference_model(sd)
    m.train()
    diff = ns.model_util.create_gaussian_diffusion(ns.model_util.get_default_diffusion())
    inp = syn.make_inputs(seed_in, B, training=True)
    fps, noise = syn.make_step_randoms(seed_rng, B, 1)
    mask = syn.make_dropout_mask(seed_drop, B)
    with torch.no_grad(), rh.injected_rng(fps_starts=list(fps[0]), dropout_masks=[mask]):
        terms = diff.training_losses(m, inp["x_start"].clone(), inp["mask"], inp["t"], inp["given_objs"], inp["given_cats"],
                                     inp["target_cat"], y=inp["text_emb"], noise=noise[0])
    new = m.state_dict()
    out = {k: np.float64(float(terms[k])) for k in ("cat_loss", "mse", "loss")}
    for key in ("pcd_backbone.sa1.mlp_bns.0", "pcd_backbone.sa2.mlp_bns.2", "pcd_backbone.sa4.mlp_bns.1", "pcd_backbone.fp4.mlp_bns.0",
                "pcd_backbone.fp2.mlp_bns.1", "pcd_backbone.fp1.mlp_bns.2", "pcd_backbone.bn1"):
        out[key + ".running_mean"] = npf(new[key + ".running_mean"])
        out[key + ".running_var"] = npf(new[key + ".running_var"])
        out[key + ".num_batches_tracked"] = npf(new[key + ".num_batches_tracked"])
    out["saved_cat"] = npf(m.saved_cat)
    out["guiding"] = npf(m.saved_guiding_points)
    return out


def case_tables():
    """Schedule tables (gaussian_diffusion.py:166-202) for T=1000 and two respacings (respace.py:8-87)."""
    ns = rh.load_reference()
    out = {}
    betas = ns.gd.get_named_beta_schedule("cosine", 1000, 1.0)
    for tag, sections in (("full", [1000]), ("ddim100", "ddim100"), ("s100", [100]), ("s10_15_20", [10, 15, 20])):
        keep = ns.respace.space_timesteps(1000 if tag != "s10_15_20" else 300, sections)
        b = betas if tag != "s10_15_20" else ns.gd.get_named_beta_schedule("linear", 300, 1.0)
        d = ns.respace.SpacedDiffusion(
            use_timesteps=keep, betas=b, model_mean_type=ns.gd.ModelMeanType.START_X,
            model_var_type=ns.gd.ModelVarType.FIXED_SMALL, loss_type=ns.gd.LossType.MSE)
        out[tag + ".keep"] = np.array(sorted(keep))
        out[tag + ".timestep_map"] = np.array(d.timestep_map)
        for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped",
                  "posterior_variance", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            out[tag + "." + k] = getattr(d, k)
    return out


def case_shard(kind="wellcond", B=4, seed_in=9, seed_rng=10):
    """Global-batch forward used to check sharded index math."""
    sd = syn.make_state_dict(SEED_W, kind)
    m = rh.build_reference_model(sd)
    inp = syn.make_inputs(seed_in, B)
    fps, _ = syn.make_step_randoms(seed_rng, B, 1)
    x = inp["x_T"].clone()
    t = torch.full((B,), 37, dtype=torch.long)
    with torch.no_grad(), rh.injected_rng(fps_starts=list(fps[0])):
        out_cat, x0 = m(x, inp["mask"], t, inp["given_objs"], inp["given_cats"], inp["text_emb"])
    return {"x0": npf(x0), "x_mutated": npf(x), "out_cat": npf(out_cat)}


def main():
    jobs = {
        "tables": case_tables,
        "forward_wellcond": lambda: case_forward("wellcond"),
        "forward_default": lambda: case_forward("default"),
        "psample_wellcond": lambda: case_p_sample("wellcond"),
        "psample_default": lambda: case_p_sample("default"),
        "loop8_wellcond": lambda: case_loop("wellcond"),
        "train_wellcond": lambda: case_train("wellcond"),
        "shard_wellcond": case_shard,
        "trainmode_wellcond": case_train_forward,
        "trainmode_b4_wellcond": lambda: case_train_forward("wellcond", 4, 16, 17, 18),
    }
    only = sys.argv[1:]
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        out = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
