"""Bitwise determinism of the pipelined STRICT loop over many steps, with the uniform-cloud closed forms on / off."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsdm_b200 import synthetic as syn  # noqa: E402
from lsdm_b200.model.sdm import SceneDiffusionModel  # noqa: E402
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd  # noqa: E402

B, K = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 23
dev = torch.device("cuda", 0)
model = SceneDiffusionModel(**get_default_model_proxd())
model.load_state_dict(syn.make_state_dict(0, "wellcond"))
model.eval()
diff = create_gaussian_diffusion(get_default_diffusion())
inp = syn.make_inputs(1234, B)
fps, noise = syn.make_step_randoms(4321, B, K)
g = {k: v.to(dev) for k, v in inp.items()}
fps_d, noise_d = fps.view(K, 4, B * 9).contiguous().to(dev), noise.to(dev)
eng = diff._engine(model, B, dev)


def run(flag, split=None):
    eng.set_option("select_uniform", flag)
    x = g["x_T"].clone()
    cuts = [0, K] if split is None else [0, split, K]
    for a, b in zip(cuts[:-1], cuts[1:]):
        eng.sample_loop(x, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_d[a:b], noise_d[a:b], 999 - a, False)
    torch.cuda.synchronize()
    return x


ref = run(1)
for name, flag, split in (("on/on", 1, None), ("on split 3", 1, 3), ("off", 0, None), ("off again", 0, None), ("off split 3", 0, 3)):
    y = run(flag, split)
    d = (y - ref).abs().max().item()
    print(f"{name:12s} equal={torch.equal(y, ref)} max|diff|={d:.3e} rel={(y - ref).norm().item() / ref.norm().item():.3e}")
eng.set_option("select_uniform", 1)
