"""Repeats bench.py's `all_clouds_encoded` leg (dedup_absent = 0, W + K steps in two calls) many times in one process, after the
legs bench.py runs before it, and reports every run whose final x differs from the default run: which samples, how much."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd

B, W, K, T = 64, 5, 20, 1000
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 16
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


def run(hoisted=False, sync_between=True):
    x = g["x_T"].clone()
    for first, n in ((0, W), (W, K)):
        eng.sample_loop(x, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_d[first:first + n] if not hoisted else fps_d[:1],
                        noise_d[first:first + n], T - 1 - first, hoisted)
        if sync_between:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return x


ref = run()
run(hoisted=True)
eng.set_option("select_uniform", 0)
print("uniform off identical:", bool(torch.equal(run(), ref)), flush=True)
eng.set_option("select_uniform", 1)
for name, opts in (("dedup_absent=0", {"dedup_absent": 0}), ("dedup_absent=0,loop_invariants=11", {"dedup_absent": 0, "loop_invariants": 11}),
                   ("dedup_absent=0,loop_invariants=7(no guiding skip)", {"dedup_absent": 0, "loop_invariants": 7}),
                   ("dedup_absent=0,time_batch=0", {"dedup_absent": 0, "time_batch": 0})):
    for k, v in opts.items():
        eng.set_option(k, v)
    bad = 0
    for i in range(REPS):
        x = run(sync_between=(i % 2 == 0))
        if not torch.equal(x, ref):
            bad += 1
            d = (x - ref).abs().amax(dim=(1, 2))
            rows = torch.nonzero(d > 0).flatten().tolist()
            print(f"  {name} run {i}: {len(rows)} samples differ {rows[:12]} max {float(d.max()):.3g} median of differing {float(d[d > 0].median()):.3g}", flush=True)
    for k in opts:
        eng.set_option(k, {"dedup_absent": 1, "time_batch": 1, "loop_invariants": 15}[k])
    print(f"{name}: {bad} of {REPS} runs differ", flush=True)

# transitions: a de-duplicated (default) run immediately before every dedup-off run, as in bench.py where the leg is the first
# dedup-off call of the process
bad = 0
TR = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for i in range(TR):
    if not torch.equal(run(), ref):
        print(f"  default run {i} differs", flush=True)
    eng.set_option("dedup_absent", 0)
    x = run()
    eng.set_option("dedup_absent", 1)
    if not torch.equal(x, ref):
        bad += 1
        d = (x - ref).abs().amax(dim=(1, 2))
        rows = torch.nonzero(d > 0).flatten().tolist()
        print(f"  transition {i}: {len(rows)} samples differ {rows[:16]} max {float(d.max()):.3g}", flush=True)
if TR:
    print(f"transitions default -> dedup_absent=0: {bad} of {TR} differ", flush=True)
