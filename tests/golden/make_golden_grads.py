"""Golden GRADIENTS of the training step from the UNMODIFIED reference (build container only) -- groundwork for SURVEY 8(f)
row 1 (backward pass): `training_losses(...)["loss"].backward()` in `model.train()` mode, as run/train_sdm.py:78-84 does.

    python tests/golden/make_golden_grads.py      ->  tests/golden/train_grads_wellcond.npz

Holds, for every parameter of the 228-entry state dict: the L2 norm and the sum of its gradient (or NaN when the reference
leaves `.grad` None -- dead parameters such as attn_layer's v_proj / out_proj), plus the full gradient of a few small tensors
spread over the network.  Inputs / weights are regenerated from seeds (lsdm_b200.synthetic), as for every other fixture.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402
from lsdm_b200 import synthetic as syn  # noqa: E402

FULL = ["output_process.pose_final.2.weight", "output_process.pose_final.2.bias", "input_process.pose_embedding.0.weight",
        "embed_timestep.time_embed.0.bias", "predict_cat.4.bias", "pcd_attention.in_proj_bias", "point_wise_trans_layer.0.weight",
        "pcd_backbone.conv2.weight", "pcd_backbone.sa1.mlp_bns.0.weight", "pcd_backbone.fp4.mlp_convs.1.bias",
        "translation_layer.2.bias", "embed_text.4.bias", "upsampling_layer.0.weight"]
CASE = dict(B=2, seed_in=21, seed_rng=22, seed_drop=23)


def main():
    ns = rh.load_reference()
    sd = syn.make_state_dict(0, "wellcond")
    m = rh.build_reference_model(sd)
    m.train()
    diff = ns.model_util.create_gaussian_diffusion(ns.model_util.get_default_diffusion())
    B = CASE["B"]
    inp = syn.make_inputs(CASE["seed_in"], B, training=True)
    fps, noise = syn.make_step_randoms(CASE["seed_rng"], B, 1)
    mask = syn.make_dropout_mask(CASE["seed_drop"], B)
    with rh.injected_rng(fps_starts=list(fps[0]), dropout_masks=[mask]):
        terms = diff.training_losses(m, inp["x_start"].clone(), inp["mask"], inp["t"], inp["given_objs"], inp["given_cats"],
                                     inp["target_cat"], y=inp["text_emb"], noise=noise[0])
        terms["loss"].backward()
    out = {"loss": np.float64(float(terms["loss"]))}
    names, norms, sums = [], [], []
    for k, p in m.named_parameters():
        if k.startswith("clip_model."):
            continue
        names.append(k)
        if p.grad is None:
            norms.append(np.nan)
            sums.append(np.nan)
        else:
            norms.append(float(p.grad.double().norm()))
            sums.append(float(p.grad.double().sum()))
        if k in FULL:
            out["grad/" + k] = p.grad.detach().numpy().astype(np.float32)
    out["names"] = np.array(names)
    out["norms"] = np.array(norms)
    out["sums"] = np.array(sums)
    np.savez_compressed(os.path.join(HERE, "train_grads_wellcond.npz"), **out)
    dead = [n for n, v in zip(names, norms) if np.isnan(v)]
    print("loss", out["loss"], "params", len(names), "dead (grad None):", dead)
    print("largest grads:", sorted(zip(norms, names), reverse=True)[:5])


if __name__ == "__main__":
    main()
