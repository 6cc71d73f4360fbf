"""Worker for the 2-GPU SyncBN test (launched by torchrun): each rank runs the train-mode forward of its shard of a global batch
of 4 with the BatchNorm statistics all-reduced over NCCL, and checks against the fixture written by the unmodified reference."""
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
from util import golden, rel_l2


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = golden("trainmode_b4_wellcond")
    Bg = 4
    per = Bg // world
    lo = rank * per
    m = SceneDiffusionModel(**{**get_default_model_proxd(), "device": local})
    m.load_state_dict(syn.make_state_dict(0, "wellcond"))
    m.train()
    m.set_shard(Bg, lo, sync_bn_group=True)
    diff = create_gaussian_diffusion(get_default_diffusion())
    inp = {k: v.cuda() for k, v in syn.make_inputs(16, Bg, training=True).items()}
    fps, noise = syn.make_step_randoms(17, Bg, 1)
    drop = syn.make_dropout_mask(18, Bg).cuda()
    sl = slice(lo, lo + per)
    dev = torch.device("cuda", local)
    eng = diff._engine(m, per, dev)
    x_t = eng.q_sample(inp["x_start"][sl].contiguous(), inp["t"][sl].contiguous(), noise[0][sl].contiguous().cuda())
    starts = torch.stack([s.view(Bg, 9)[sl].reshape(-1) for s in fps[0]])
    m.encode(inp["mask"], inp["given_objs"][sl].contiguous(), inp["given_cats"][sl].contiguous(), inp["text_emb"][sl].contiguous(), starts,
             device=dev, drop_mask=drop[lo * 9:(lo + per) * 9].contiguous())
    out_cat, x0, guiding = m._engine.forward(x_t, inp["t"][sl].contiguous())
    torch.cuda.synchronize()
    assert rel_l2(guiding.cpu(), g["guiding"][sl]) < 1e-3, rel_l2(guiding.cpu(), g["guiding"][sl])
    sd = m.state_dict()
    for k in g:
        if k.endswith("running_mean") or k.endswith("running_var"):
            r = rel_l2(sd[k].cpu(), g[k])
            assert r < 5e-3, (k, r)
    # losses: global means = mean over ranks of the per-shard means (equal shards)
    cat = eng.cat_loss(out_cat, inp["target_cat"][sl].contiguous()) * 0.1
    mse = eng.chamfer(x0, inp["x_start"][sl].contiguous())
    both = torch.stack([cat, mse])
    dist.all_reduce(both)
    both /= world
    assert abs(float(both[0]) - float(g["cat_loss"])) < 1e-3 * float(g["cat_loss"])
    assert abs(float(both[1]) - float(g["mse"])) < 1e-3 * float(g["mse"])
    dist.barrier()
    if rank == 0:
        print("SYNCBN_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
