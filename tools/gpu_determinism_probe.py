"""Runs the bench workload's W + K steps under several option settings, several times each, and reports whether the final x is
bit-identical to the default run (dedup / select_uniform / time_batch legs must be; loop_invariants 0 differs by the split sums)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, K, T = 5, 20, 1000
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


def run():
    x = g["x_T"].clone()
    for first, n in ((0, W), (W, K)):
        eng.sample_loop(x, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_d[first:first + n], noise_d[first:first + n], T - 1 - first, False)
    torch.cuda.synchronize()
    return x


ref = run()
for name, opts in (("default", {}), ("dedup_absent=0", {"dedup_absent": 0}), ("time_batch=0", {"time_batch": 0}),
                   ("dedup_absent=0,time_batch=0", {"dedup_absent": 0, "time_batch": 0}), ("loop_invariants=13", {"loop_invariants": 13}),
                   ("dedup_absent=0,loop_invariants=11", {"dedup_absent": 0, "loop_invariants": 11}),
                   ("dedup_absent=0,loop_invariants=14", {"dedup_absent": 0, "loop_invariants": 14}),
                   ("dedup_absent=0,loop_invariants=7", {"dedup_absent": 0, "loop_invariants": 7})):
    for k, v in opts.items():
        eng.set_option(k, v)
    res = []
    for _ in range(4):
        x = run()
        res.append((bool(torch.equal(x, ref)), float((x - ref).abs().max())))
    for k in opts:
        eng.set_option(k, {"dedup_absent": 1, "time_batch": 1, "loop_invariants": 15}[k])
    print(name, res, flush=True)
