"""default (de-duplicated) run -> dedup_absent = 0 run, repeated, under several loop_invariants settings: how often the second differs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd

B, W, K, T = 64, 5, 20, 1000
TR = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
model = SceneDiffusionModel(**{**get_default_model_proxd(), "device": 0})
model.load_state_dict(syn.make_state_dict(0, "wellcond"))
model.eval()
diff = create_gaussian_diffusion(get_default_diffusion())
inp = syn.make_inputs(1234, B)
fps, noise = syn.make_step_randoms(4321, B, W + K)
g = {k: v.to(dev) for k, v in inp.items()}
fps_d, noise_d = fps.to(dev), noise.to(dev)
eng = diff._engine(model, B, dev)
absent = (inp["given_objs"].abs().sum((2, 3)) == 0)


def run():
    x = g["x_T"].clone()
    for first, n in ((0, W), (W, K)):
        eng.sample_loop(x, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_d[first:first + n], noise_d[first:first + n], T - 1 - first, False)
    torch.cuda.synchronize()
    return x


for inv in (0, 11, 4):
    eng.set_option("loop_invariants", inv)
    ref = run()
    bad = 0
    for i in range(TR):
        run()
        eng.set_option("dedup_absent", 0)
        x = run()
        eng.set_option("dedup_absent", 1)
        if not torch.equal(x, ref):
            bad += 1
            d = (x - ref).abs().amax(dim=(1, 2))
            rows = torch.nonzero(d > 0).flatten().tolist()
            print(f"  inv={inv} transition {i}: samples {rows[:8]} max {float(d.max()):.3g} absent-of-first {absent[rows[0]].int().tolist()}", flush=True)
    print(f"loop_invariants={inv}: {bad} of {TR} transitions differ", flush=True)
eng.set_option("loop_invariants", 15)
