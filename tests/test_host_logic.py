"""CPU tests of the host side: C-ABI surface, reference-shaped Python API, schedule tables, sharding plan (gloo, world 2)."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from lsdm_b200 import synthetic as syn
from util import golden, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from lsdm_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "lsdm_b200.h")).read()
    declared = set(re.findall(r"LSDM_API[^;()]*?\b(lsdm_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _lib.load()  # no compute calls: loading needs no GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert b"sm_100a" in lib.lsdm_version()


def test_cabi_argument_errors_without_gpu():
    from lsdm_b200 import _lib

    lib = _lib.load()
    assert lib.lsdm_create(None, None) == _lib.EINVAL
    assert b"null" in lib.lsdm_last_error()
    assert lib.lsdm_num_weights(None) == 0
    assert lib.lsdm_workspace_bytes(None) == 0
    assert lib.lsdm_launch_count(None) == 0


def test_cabi_argument_errors_of_the_widened_rows_without_gpu():
    """Evaluation metrics, CLIP tower and ContactFormer layer entry points validate their arguments before touching the device."""
    import ctypes as C

    from lsdm_b200 import _lib

    lib = _lib.load()
    assert lib.lsdm_eval_emd(None, None, 1, 8, 8, None, None, None, None) == _lib.EINVAL
    one = C.c_void_p(16)  # never dereferenced: the size checks come first
    assert lib.lsdm_eval_emd(one, one, 1, 8, 9, one, None, None, None) == _lib.EINVAL and b"same number" in lib.lsdm_last_error()
    assert lib.lsdm_eval_emd(one, one, 1, 2048, 2048, one, None, None, None) == _lib.EINVAL and b"1024" in lib.lsdm_last_error()
    assert lib.lsdm_eval_fscore(one, one, 1, 5000, 8, C.c_double(0.1), one, one, None) == _lib.EINVAL
    assert lib.lsdm_eval_topk(one, one, 0, 13, one, 1, one, None) == _lib.EINVAL
    assert lib.lsdm_clip_create(None, 0) == _lib.EINVAL
    assert lib.lsdm_clip_finalize(None) == _lib.EINVAL
    assert lib.lsdm_clip_workspace_bytes(None, 4, 22) == 0 and lib.lsdm_clip_launch_count(None) == 0
    assert lib.lsdm_clip_encode_text(None, None, 1, 22, None, 0, None, None) == _lib.EINVAL
    assert lib.lsdm_cf_workspace_bytes(0, 16, 5, 8) == 0 and lib.lsdm_cf_workspace_bytes(1, 256, 655, 8) > 0
    assert lib.lsdm_cf_mha_forward(None, None, None, 0, 1, 16, 5, 8, 1, None, 0, None, None) == _lib.EINVAL
    assert lib.lsdm_cf_set_option(b"no_such_option", 1) == _lib.EINVAL
    w = _lib.CfMhaWeights()
    assert lib.lsdm_cf_mha_forward(C.byref(w), one, None, 0, 1, 16, 5, 8, 1, one, 0, one, None) == _lib.EINVAL  # null weights


def test_state_dict_contract():
    from lsdm_b200.util.model_util import create_model_and_diffusion

    for datatype, cats in (("proxd", 13), ("humanise", 11)):
        model, diffusion = create_model_and_diffusion(datatype)
        sd = syn.make_state_dict(0, "wellcond", cats)
        own = model.state_dict()
        assert set(own) == set(sd)
        for k in own:
            assert own[k].shape == sd[k].shape, k
        ckpt = dict(sd)
        ckpt["clip_model.positional_embedding"] = torch.zeros(77, 512)  # reference checkpoints carry the CLIP tower
        model.load_state_dict(ckpt)  # strict
        assert diffusion.num_timesteps == 1000
        assert hasattr(diffusion, "ddim_sample_loop")
        assert sum(p.numel() for p in model.parameters()) == 2418986 - (13 - cats) * (32 + 32 + 1)


def test_schedule_tables_match_reference():
    from lsdm_b200.diffusion import gaussian_diffusion as gd
    from lsdm_b200.diffusion.respace import SpacedDiffusion, space_timesteps

    g = golden("tables")
    cos = gd.get_named_beta_schedule("cosine", 1000)
    for tag, sections, betas, T0 in (("full", [1000], cos, 1000), ("ddim100", "ddim100", cos, 1000), ("s100", [100], cos, 1000),
                                     ("s10_15_20", [10, 15, 20], gd.get_named_beta_schedule("linear", 300), 300)):
        keep = space_timesteps(T0, sections)
        d = SpacedDiffusion(use_timesteps=keep, betas=betas, model_mean_type=gd.ModelMeanType.START_X,
                            model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
        assert sorted(keep) == list(g[tag + ".keep"])
        assert d.timestep_map == list(g[tag + ".timestep_map"])
        for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped", "posterior_variance",
                  "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            np.testing.assert_allclose(getattr(d, k), g[tag + "." + k], rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")


def test_unsupported_options_fail_loudly():
    from lsdm_b200.diffusion import gaussian_diffusion as gd
    from lsdm_b200.model.sdm import SceneDiffusionModel
    from lsdm_b200.util.model_util import get_default_model_proxd

    b = gd.get_named_beta_schedule("cosine", 10)
    with pytest.raises(NotImplementedError):
        gd.GaussianDiffusion(betas=b, model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
    with pytest.raises(NotImplementedError):
        gd.GaussianDiffusion(betas=b, model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.LEARNED, loss_type=gd.LossType.MSE)
    with pytest.raises(NotImplementedError):
        SceneDiffusionModel(**{**get_default_model_proxd(), "latent_dim": 256})
    m = SceneDiffusionModel(**get_default_model_proxd())
    inp = syn.make_inputs(1, 1)
    with pytest.raises(ValueError):  # CPU tensors: no fallback
        m(inp["x_T"], inp["mask"], torch.zeros(1, dtype=torch.long), inp["given_objs"], inp["given_cats"], inp["text_emb"])
    with pytest.raises(NotImplementedError):
        m._encode_text(["a chair"])


def test_fps_start_draws_follow_reference_order_and_shard():
    from lsdm_b200.model.sdm import SceneDiffusionModel
    from lsdm_b200.util.model_util import get_default_model_proxd

    m = SceneDiffusionModel(**get_default_model_proxd())
    torch.manual_seed(5)
    ref = [torch.randint(0, n, (4 * 9,), dtype=torch.long) for n in (1024, 1024, 256, 64)]
    torch.manual_seed(5)
    full = m.draw_fps_starts(4)
    assert torch.equal(full, torch.stack(ref))
    m.set_shard(4, 1)
    torch.manual_seed(5)
    part = m.draw_fps_starts(2)
    assert torch.equal(part, torch.stack([r.view(4, 9)[1:3].reshape(-1) for r in ref]))


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lsdm_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    B = 4
    per = B // world
    lo = rank * per
    sd = syn.make_state_dict(0, "wellcond")
    inp = syn.make_inputs(9, B)  # every rank regenerates the global tensors and slices its shard
    fps, _ = syn.make_step_randoms(10, B, 1)
    starts = [s.view(B, 9)[lo:lo + per].reshape(-1) for s in fps[0]]
    t = torch.full((per,), 37, dtype=torch.long)
    x = inp["x_T"][lo:lo + per].clone()
    _, x0, _ = O.forward(sd, x, inp["mask"][lo:lo + per], t, inp["given_objs"][lo:lo + per], inp["given_cats"][lo:lo + per],
                         inp["text_emb"][lo:lo + per], starts, mask_global=inp["mask"], b_offset=lo)
    out = [torch.empty_like(x0) for _ in range(world)]
    dist.all_gather(out, x0.contiguous())  # the path's single exchange step
    if rank == 0:
        q.put(torch.cat(out).numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_plan_gloo():
    """world_size 2 over gloo: shard -> local compute with global mask/offset -> one all-gather == global-batch reference."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = golden("shard_wellcond")
    assert rel_l2(got, g["x0"]) < 2e-5


def test_fps_kernel_has_no_contracted_fma():
    """The FPS distance is (dx^2 + dy^2) + dz^2 with separately rounded products (reference pointnet2_utils.py:60-83).
    ptxas contracts packed mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the rounding modifiers
    (tools/probe/packed_fma_probe.cu), which silently changes 21 % of the distances; the kernel therefore adds in scalar.
    Guard: the compiled FPS kernel must not contain any fused multiply-add."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lsdm_b200", "liblsdm_b200.so")
    if not os.path.exists(cuobjdump) or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    fps = [b for b in blocks if "fps4_kernel" in b.split("\n", 1)[0]]
    assert fps, "fps4_kernel not found in the library"
    for b in fps:
        assert not re.search(r"\bFFMA2?\b", b), "FPS kernel contains a fused multiply-add (contracted mul+add)"


def test_clip_bpe_tokenizer_matches_transformers_on_a_synthetic_vocabulary(tmp_path):
    """lsdm_b200.model.clip_tokenizer (restatement of clip.tokenize; the clip package and its vocabulary file are absent) against
    transformers.CLIPTokenizer -- an independent implementation of the same BPE -- on a merge table learned from a toy corpus,
    written in the clip package's file format."""
    import collections
    import gzip

    import torch
    from transformers import CLIPTokenizer

    from lsdm_b200.model.clip_tokenizer import ClipBpeTokenizer, _byte_alphabet

    corpus = ("a person sits on the chair next to the table . the sofa is in front of the television , a lamp stands behind the bed "
              "place a small wooden cabinet near the window and the door ; it's the person's desk , they've moved it 3 times in 2021 !").split()
    _, alphabet = _byte_alphabet()
    words = collections.Counter(tuple(list(w[:-1]) + [w[-1] + "</w>"]) for w in corpus)
    merges = []
    for _ in range(120):   # plain BPE training: most frequent pair first
        pairs = collections.Counter()
        for w, c in words.items():
            for p in zip(w, w[1:]):
                pairs[p] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1]); i += 2
                else:
                    out.append(w[i]); i += 1
            new[tuple(out)] += c
        words = new
    path = tmp_path / "bpe_toy.txt.gz"
    with gzip.open(path, "wb") as f:
        f.write(("#version: toy\n" + "\n".join(a + " " + b for a, b in merges) + "\n").encode("utf-8"))
    mine = ClipBpeTokenizer(str(path))
    symbols = alphabet + [s + "</w>" for s in alphabet] + [a + b for a, b in merges] + ["<|startoftext|>", "<|endoftext|>"]
    assert len(set(symbols)) == len(symbols) and mine.sot == len(symbols) - 2 and mine.eot == len(symbols) - 1
    ref = CLIPTokenizer(vocab={s: i for i, s in enumerate(symbols)}, merges=[(a, b) for a, b in merges])
    texts = ["A person sits on the chair next to the table.", "it's the person's desk, they've moved it 3 times in 2021!",
             "  place   a small wooden cabinet\tnear the window  ", "café naïve über 12ab", "the sofa&amp;the bed", "zzz qqq ???", ""]
    import html
    for t in texts:   # (clip's text cleaning un-escapes HTML entities twice; transformers' tokenizer does not: done here for it)
        assert mine.encode(t) == ref(html.unescape(html.unescape(t)), add_special_tokens=False)["input_ids"], t
    out = mine.tokenize(texts, context_length=22, truncate=True)
    assert out.shape == (len(texts), 22) and out.dtype == torch.long
    for row, t in zip(out, texts):
        ids = [mine.sot] + mine.encode(t) + [mine.eot]
        if len(ids) > 22:
            ids = ids[:22]
            ids[-1] = mine.eot
        assert row[:len(ids)].tolist() == ids and (row[len(ids):] == 0).all()
    with pytest.raises(RuntimeError):
        mine.tokenize(["the chair " * 40], context_length=22, truncate=False)
    with pytest.raises(FileNotFoundError):
        ClipBpeTokenizer(str(tmp_path / "missing.gz"))


@pytest.mark.parametrize("shard", [None, (8, 2)])
def test_fps_start_draws_for_a_run_of_steps_equal_the_per_step_draws(shard):
    """p_sample_loop draws a library call's FPS starts up front: one generator call for the whole run of steps must yield the
    numbers of the reference's per-step ``randint(0, N, (9B,))`` calls (pointnet2_utils.py:72), at GLOBAL shape when sharded, and
    leave the CPU generator in the same state; the one-time self-check must not consume from the caller's stream."""
    from lsdm_b200.model import sdm

    class Draws:
        _shard = shard
        draw_fps_starts = sdm.SceneDiffusionModel.draw_fps_starts
        draw_fps_starts_steps = sdm.SceneDiffusionModel.draw_fps_starts_steps

    d, B, K = Draws(), (64 if shard is None else 3), 20
    sdm._BATCHED_DRAW_OK = None  # the self-check runs inside this test's stream
    torch.manual_seed(11)
    torch.rand(5)
    seq = torch.stack([d.draw_fps_starts(B) for _ in range(K)])
    state = torch.get_rng_state()
    for with_first in (False, True):
        torch.manual_seed(11)
        torch.rand(5)
        first = d.draw_fps_starts(B) if with_first else None
        got = d.draw_fps_starts_steps(B, K, first=first)
        assert torch.equal(got, seq)
        assert torch.equal(torch.get_rng_state(), state)
    assert sdm._BATCHED_DRAW_OK is True  # (False would mean torch's CPU randint changed: the per-call path is then used)


def test_fps_on_live_points_only_is_the_same_selection():
    """numpy model of the compacting FPS of pointnet_select.cu (level 0 keeps all N points: every 128 rounds the points at
    distance 0 are dropped and the live ones re-dealt in index order; argmax = first maximum, index 0 when every tracked
    distance is 0) against the oracle's farthest_point_sample: random clouds, duplicated points, lattice ties, coincident points."""
    import lsdm_oracle as O

    rs = np.random.RandomState(3)
    clouds = [rs.uniform(-0.5, 0.5, (256, 3)).astype(np.float32)]
    dup = rs.uniform(-0.5, 0.5, (256, 3)).astype(np.float32)
    dup[128:] = dup[:128]
    clouds.append(dup)
    clouds.append((rs.randint(0, 4, (256, 3)) * 0.25).astype(np.float32))
    one = np.zeros((256, 3), np.float32)
    one[7] = 3.0
    clouds.append(one)
    xyz = torch.from_numpy(np.stack(clouds))
    start = torch.tensor([5, 200, 17, 0])
    ref = O.farthest_point_sample(xyz, 256, start).numpy()
    for c, pts in enumerate(np.stack(clouds)):
        n = len(pts)
        idx = np.arange(n)               # tracked points: original indices, ascending
        dist = np.full(n, 1e10, np.float32)
        far, out = int(start[c]), []
        for it in range(n):
            if it and it % 32 == 0:      # compaction (the kernel: every 128 of 1024 rounds)
                keep = dist != 0
                idx, dist = idx[keep], dist[keep]
            out.append(far)
            d = pts[idx] - pts[far]
            d = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            dist = np.minimum(dist, d.astype(np.float32))
            far = int(idx[np.argmax(dist)]) if len(dist) and dist.max() > 0 else 0
        assert out == list(ref[c]), c


def test_level1_in_cloud_order_is_the_reference_result_permuted():
    """The argument behind `loop_invariants` bit 2, checked on the oracle: sa1 keeps all N points, so its output for
    the reference's FPS-ordered centroids is the cloud-order output gathered by the FPS indices -- also when duplicate points make the
    FPS order repeat index 0 instead of being a permutation --, and the level-1 ball query run on cloud-order points gives the
    reference's groups with the FPS indices composed in (same neighbours, same padding, same pooled features at level 2)."""
    import lsdm_oracle as O

    sd = syn.make_state_dict(0, "wellcond")
    rs = np.random.RandomState(7)
    N = 1024
    pts = rs.uniform(-0.5, 0.5, (3, N, 3)).astype(np.float32)
    pts[1, 512:] = pts[1, :512]                       # every point twice: the level-0 FPS order is not a permutation
    pts[2] = (rs.randint(0, 6, (N, 3)) * 0.2 - 0.5)   # lattice: distance ties, duplicates, full balls
    xyz = torch.from_numpy(pts)
    start0, start1 = torch.tensor([5, 900, 17]), torch.tensor([3, 77, 1000])
    # reference order: sa1 with the FPS draw, then sa2 on its output
    tr = {}
    l1_xyz, l1_f = O.set_abstraction(sd, "sa1", 1024, 0.1, 32, xyz, xyz, start0, trace=tr)
    _, l2_f = O.set_abstraction(sd, "sa2", 256, 0.2, 32, l1_xyz, l1_f, start1, trace=tr)
    pi0 = tr["sa1.fps_idx"]
    assert len(set(pi0[1].tolist())) < N              # the duplicated cloud really repeats indices
    # cloud order: centroids = the cloud's own points, no FPS
    grp_c = O.ball_query(0.1, 32, xyz, xyz)
    h = torch.cat([O._gather(xyz, grp_c) - xyz[:, :, None, :], O._gather(xyz, grp_c)], dim=-1)
    for i in range(3):
        h = O._conv_bn_relu(sd, "pcd_backbone.sa1", i, h)
    f1_cloud = h.max(dim=2)[0]
    assert torch.equal(O._gather(f1_cloud, pi0), l1_f)
    assert torch.equal(O._gather(xyz, pi0), l1_xyz)
    # level 2 from cloud-order level 1: groups found on the FPS-ordered coordinates (the truncation to the first 32 in index order is
    # in FPS order), stored as cloud indices; features and coordinates gathered from the cloud-order arrays through them
    fps1 = tr["sa2.fps_idx"]
    new_xyz = O._gather(l1_xyz, fps1)
    grp = O.ball_query(0.2, 32, l1_xyz, new_xyz)
    assert torch.equal(grp, tr["sa2.group_idx"])
    grp_composed = O._gather(pi0, grp.reshape(3, -1)).reshape(grp.shape)
    h = torch.cat([O._gather(xyz, grp_composed) - new_xyz[:, :, None, :], O._gather(f1_cloud, grp_composed)], dim=-1)
    for i in range(3):
        h = O._conv_bn_relu(sd, "pcd_backbone.sa2", i, h)
    assert torch.equal(h.max(dim=2)[0], l2_f)
