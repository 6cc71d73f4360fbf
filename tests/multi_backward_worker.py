"""Worker for the 2-GPU training-step test (launched by torchrun): global batch 4 over 2 ranks.  Each rank back-propagates the loss
of its shard (SyncBN statistics AND the BatchNorm backward sums all-reduced over NCCL inside the library), the gradients are
averaged over the ranks with one all-reduce (what FusedAdamW(all_reduce=True) does), and rank 0 checks them against the
gradients of the same global batch computed unsharded on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd
from util import injected_rng, rel_l2


def run(m, diff, inp, fps, noise, drop, lo, hi, Bg):
    sl = slice(lo, hi)
    m.draw_dropout_mask = lambda B, dev: drop[lo * 9:hi * 9].contiguous().to(dev)
    m.zero_grad(set_to_none=True)
    with injected_rng(fps_starts=list(fps[0])):   # drawn at GLOBAL shape, sliced inside
        terms = diff.training_losses(m, inp["x_start"][sl].contiguous(), inp["mask"], inp["t"][sl].contiguous(), inp["given_objs"][sl].contiguous(),
                                     inp["given_cats"][sl].contiguous(), inp["target_cat"][sl].contiguous(), y=inp["text_emb"][sl].contiguous(),
                                     noise=noise[0][sl].contiguous().cuda())
    terms["loss"].backward()
    torch.cuda.synchronize()
    return terms


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Bg = 4
    per = Bg // world
    lo = rank * per
    inp = {k: v.cuda() for k, v in syn.make_inputs(16, Bg, training=True).items()}
    fps, noise = syn.make_step_randoms(17, Bg, 1)
    drop = syn.make_dropout_mask(18, Bg)
    diff = create_gaussian_diffusion(get_default_diffusion())
    m = SceneDiffusionModel(**{**get_default_model_proxd(), "device": local})
    m.load_state_dict(syn.make_state_dict(0, "wellcond"))
    m.train()
    m.set_shard(Bg, lo, sync_bn_group=True)
    terms = run(m, diff, inp, fps, noise, drop, lo, lo + per, Bg)
    names = [n for n, p in m.named_parameters() if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for n, p in m.named_parameters() if p.grad is not None])
    loss = terms["loss"].detach().clone()
    dist.all_reduce(flat)
    dist.all_reduce(loss)
    flat /= world
    loss /= world
    if rank == 0:
        m1 = SceneDiffusionModel(**{**get_default_model_proxd(), "device": local})
        m1.load_state_dict(syn.make_state_dict(0, "wellcond"))
        m1.train()
        t1 = run(m1, diff, inp, fps, noise, drop, 0, Bg, Bg)
        assert abs(float(loss) - float(t1["loss"])) < 1e-4 * abs(float(t1["loss"])), (float(loss), float(t1["loss"]))
        off, worst = 0, ("", 0.0)
        ref = dict(m1.named_parameters())
        for n in names:
            g1 = ref[n].grad
            k = g1.numel()
            got = flat[off:off + k].view_as(g1)
            off += k
            if (".mlp_convs." in n or n == "pcd_backbone.conv1.bias") and n.endswith(".bias"):
                continue  # mathematically zero in front of a train-mode BatchNorm
            e = rel_l2(got.cpu(), g1.cpu())
            if e > worst[1]:
                worst = (n, e)
            assert e < 3e-3, (n, e)
        print("worst sharded-vs-unsharded gradient rel-L2:", worst)
    dist.barrier()
    if rank == 0:
        print("SHARDED_BACKWARD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
