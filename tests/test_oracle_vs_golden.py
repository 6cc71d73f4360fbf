"""Pins oracle/lsdm_oracle.py against outputs of the unmodified reference (tests/golden/*.npz,
written by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import lsdm_oracle as O
from lsdm_b200 import synthetic as syn
from util import golden, rel_l2

TOL = 2e-5  # fp32 CPU restatement vs fp32 CPU reference: summation-order noise only


@pytest.mark.parametrize("kind", ["wellcond", "default"])
def test_forward_stages(kind):
    g = golden("forward_" + kind)
    sd = syn.make_state_dict(0, kind)
    B = 3
    inp = syn.make_inputs(1, B)
    fps, _ = syn.make_step_randoms(2, B, 1)
    x = inp["x_T"].clone()
    tr = {}
    oc, x0, guiding = O.forward(sd, x, inp["mask"], torch.from_numpy(g["t"]), inp["given_objs"], inp["given_cats"],
                                inp["text_emb"], list(fps[0]), trace=tr)
    # discrete selections are bit-exact (integer work)
    for lvl, name in enumerate(("sa1", "sa2", "sa3", "sa4")):
        assert np.array_equal(tr[name + ".fps_idx"][:2].numpy(), g[f"fps_idx{lvl}"].astype(np.int64)), name
        assert np.array_equal(tr[name + ".group_idx"][:2].numpy(), g[f"ball_idx{lvl}"].astype(np.int64)), name
    assert rel_l2(tr["enc"], g["enc"]) < TOL
    assert rel_l2(tr["tr"], g["tr"]) < TOL
    assert rel_l2(tr["attn_w"], g["attn_w"]) < TOL
    assert rel_l2(tr["hm"], g["hm"]) < TOL
    assert rel_l2(tr["backbone"].reshape(B * 9, 1024, 3), g["backbone"]) < TOL
    assert rel_l2(tr["pa"].reshape(B * 9, 12), g["pa"]) < TOL
    assert rel_l2(tr["pw"][:, :, ::4], g["pw_sub"]) < TOL
    assert rel_l2(tr["emb"][:, ::8], g["emb_sub"]) < TOL
    assert rel_l2(x, g["x_mutated"]) < TOL  # in-place x += pcd_out (trap 1)
    assert rel_l2(oc, g["out_cat"]) < TOL
    assert rel_l2(x0, g["x0"]) < TOL
    assert rel_l2(guiding, g["guiding"]) < TOL


@pytest.mark.parametrize("kind", ["wellcond", "default"])
def test_p_sample_config1(kind):
    """BASELINE config 1: 1-step p_sample, batch 2."""
    g = golden("psample_" + kind)
    sd = syn.make_state_dict(0, kind)
    tables = O.diffusion_tables(O.cosine_betas(1000))
    inp = syn.make_inputs(3, 2)
    fps, noise = syn.make_step_randoms(4, 2, 1)
    x = inp["x_T"].clone()
    t = torch.full((2,), 999, dtype=torch.long)
    out = O.p_sample(sd, tables, x, inp["mask"], t, inp["given_objs"], inp["given_cats"], inp["text_emb"], list(fps[0]), noise[0])
    assert rel_l2(out["sample"], g["sample"]) < TOL
    assert rel_l2(out["pred_xstart"], g["pred_xstart"]) < TOL
    assert rel_l2(x, g["x_mutated"]) < TOL
    assert rel_l2(out["out_cat"], g["saved_cat"]) < TOL
    assert rel_l2(out["guiding"], g["guiding"]) < TOL


def test_respaced_loop():
    g = golden("loop8_wellcond")
    sd = syn.make_state_dict(0, "wellcond")
    keep = O.space_timesteps(1000, "8")
    assert keep == list(g["keep"])
    tables = O.diffusion_tables(O.respaced_betas(O.cosine_betas(1000), keep))
    T = int(g["T"])
    assert T == 8
    inp = syn.make_inputs(5, 2)
    fps, noise = syn.make_step_randoms(6, 2, T)
    out = O.p_sample_loop(sd, tables, inp["x_T"], inp["mask"], inp["given_objs"], inp["given_cats"], inp["text_emb"], fps, noise)
    assert rel_l2(out["sample"], g["sample"]) < TOL
    assert rel_l2(out["guiding"], g["guiding"]) < TOL
    assert rel_l2(out["out_cat"], g["saved_cat"]) < TOL


def test_training_losses():
    g = golden("train_wellcond")
    sd = syn.make_state_dict(0, "wellcond")
    tables = O.diffusion_tables(O.cosine_betas(1000))
    inp = syn.make_inputs(7, 4, training=True)
    fps, noise = syn.make_step_randoms(8, 4, 1)
    terms = O.training_losses(sd, tables, inp["x_start"], inp["mask"], inp["t"], inp["given_objs"], inp["given_cats"],
                              inp["target_cat"], inp["text_emb"], list(fps[0]), noise[0])
    for k in ("cat_loss", "mse", "loss"):
        assert abs(float(terms[k]) - float(g[k])) <= 2e-5 * abs(float(g[k])), k


def test_schedule_tables():
    g = golden("tables")
    cos = O.cosine_betas(1000)
    for tag, sections, betas, T0 in (("full", [1000], cos, 1000), ("ddim100", "ddim100", cos, 1000),
                                     ("s100", [100], cos, 1000), ("s10_15_20", [10, 15, 20], O.linear_betas(300), 300)):
        keep = O.space_timesteps(T0, sections)
        assert keep == list(g[tag + ".keep"]) == list(g[tag + ".timestep_map"])
        tb = O.diffusion_tables(O.respaced_betas(betas, keep))
        for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped",
                  "posterior_variance", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            np.testing.assert_allclose(tb[k], g[tag + "." + k], rtol=1e-12, atol=0)
    assert list(g["ddim100.keep"]) == list(range(0, 1000, 10))


def test_sharded_index_math():
    """A shard [lo:hi) computed from local data + GLOBAL mask + global offsets reproduces the rows of the
    global-batch reference forward; naive sharding does not."""
    g = golden("shard_wellcond")
    sd = syn.make_state_dict(0, "wellcond")
    B = 4
    inp = syn.make_inputs(9, B)
    fps, _ = syn.make_step_randoms(10, B, 1)
    t = torch.full((B,), 37, dtype=torch.long)
    for lo, hi in ((0, 2), (2, 4), (1, 2)):
        x = inp["x_T"][lo:hi].clone()
        starts = [s.view(B, 9)[lo:hi].reshape(-1) for s in fps[0]]
        _, x0, _ = O.forward(sd, x, inp["mask"][lo:hi], t[lo:hi], inp["given_objs"][lo:hi], inp["given_cats"][lo:hi],
                             inp["text_emb"][lo:hi], starts, mask_global=inp["mask"], b_offset=lo)
        assert rel_l2(x0, g["x0"][lo:hi]) < TOL
        assert rel_l2(x, g["x_mutated"][lo:hi]) < TOL
    x = inp["x_T"][2:4].clone()
    starts = [s.view(B, 9)[2:4].reshape(-1) for s in fps[0]]
    O.forward(sd, x, inp["mask"][2:4], t[2:4], inp["given_objs"][2:4], inp["given_cats"][2:4], inp["text_emb"][2:4], starts)
    assert rel_l2(x, g["x_mutated"][2:4]) > 1e-3  # naive sharding is visibly wrong


def test_as_written_attention_equals_collapsed():
    sd = syn.make_state_dict(0, "wellcond")
    rs = np.random.RandomState(0)
    tr = torch.from_numpy(rs.standard_normal((1, 9, 12)).astype(np.float32))
    p1 = torch.from_numpy(rs.standard_normal((1, 9, 1024, 3)).astype(np.float32))
    a = O.point_attention(sd, tr, p1, as_written=True)
    b = O.point_attention(sd, tr, p1, as_written=False)
    assert rel_l2(a, b) < 1e-5


def test_train_mode_forward_and_bn_updates():
    """model.train(): BatchNorm batch statistics, running-stat updates (momentum 0.1, unbiased var), Dropout(0.5) mask injected."""
    g = golden("trainmode_wellcond")
    sd = syn.make_state_dict(0, "wellcond")
    B = 3
    inp = syn.make_inputs(13, B, training=True)
    fps, noise = syn.make_step_randoms(14, B, 1)
    mask = syn.make_dropout_mask(15, B)
    tables = O.diffusion_tables(O.cosine_betas(1000))
    tr = {}
    terms = O.training_losses(sd, tables, inp["x_start"], inp["mask"], inp["t"], inp["given_objs"], inp["given_cats"], inp["target_cat"],
                              inp["text_emb"], list(fps[0]), noise[0], train=tr, drop_mask=mask)
    for k in ("cat_loss", "mse", "loss"):
        assert abs(float(terms[k]) - float(g[k])) <= 2e-5 * abs(float(g[k])), k
    n = 0
    for k in g:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel_l2(tr["updates"][k], g[k]) < TOL, k
            n += 1
        elif k.endswith("num_batches_tracked"):
            assert int(tr["updates"][k]) == int(g[k])
    assert n == 14 and len([k for k in tr["updates"] if k.endswith("running_mean")]) == 22


def test_training_gradients_match_reference():
    """Groundwork for SURVEY 8(f) row 1: autograd on the oracle's functional forward reproduces the gradients the unmodified
    reference computes with loss.backward() in model.train() mode (norm and sum of every parameter's gradient, full tensors for
    13 parameters spread over the network, and the same set of dead parameters)."""
    from golden.make_golden_grads import CASE, FULL

    g = golden("train_grads_wellcond")
    sd = syn.make_state_dict(0, "wellcond")
    B = CASE["B"]
    inp = syn.make_inputs(CASE["seed_in"], B, training=True)
    fps, noise = syn.make_step_randoms(CASE["seed_rng"], B, 1)
    drop = syn.make_dropout_mask(CASE["seed_drop"], B)
    tables = O.diffusion_tables(O.cosine_betas(1000))
    loss, grads = O.training_grads(sd, tables, inp["x_start"], inp["mask"], inp["t"], inp["given_objs"], inp["given_cats"],
                                   inp["target_cat"], inp["text_emb"], list(fps[0]), noise[0], drop)
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    names = [str(n) for n in g["names"]]
    assert set(names) <= set(grads)
    for n, norm, s in zip(names, g["norms"], g["sums"]):
        gr = grads[n]
        if np.isnan(norm):
            assert gr is None or float(gr.abs().max()) == 0.0, n    # dead in the reference: no gradient reaches it
            continue
        assert gr is not None, n
        # conv biases in front of a train-mode BatchNorm have a mathematically zero gradient (the batch mean absorbs them): both
        # sides hold rounding noise there (<= 2e-5 against weight gradients of 1e-2 .. 1), hence the absolute term
        assert abs(float(gr.double().norm()) - norm) <= 2e-3 * norm + 3e-5, (n, float(gr.double().norm()), norm)
    for n in FULL:
        if ".mlp_convs." in n and n.endswith(".bias"):
            continue  # zero gradient up to rounding noise (see above)
        assert rel_l2(grads[n], g["grad/" + n]) < 2e-3, n
