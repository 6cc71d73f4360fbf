"""Stage timeline of lsdm_sample_loop (LSDM_TIMELINE=1): python tools/gpu_timeline.py [B] [steps]"""
import os, sys
os.environ["LSDM_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
model = SceneDiffusionModel(**get_default_model_proxd())
model.load_state_dict(syn.make_state_dict(0, "wellcond"))
model.eval()
diff = create_gaussian_diffusion(get_default_diffusion())
inp = {k: v.cuda() for k, v in syn.make_inputs(1234, B).items()}
fps, noise = syn.make_step_randoms(4321, B, K)
eng = diff._engine(model, B, torch.device("cuda", 0))
for dedup in (1, 0):
    eng.set_option("dedup_absent", dedup)
    for rep in range(2):
        print(f"=== dedup_absent={dedup} rep={rep}", file=sys.stderr, flush=True)
        x = inp["x_T"].clone()
        eng.sample_loop(x, inp["text_emb"], inp["given_objs"], inp["given_cats"], inp["mask"], fps.cuda(), noise.cuda(), 999, False)
        torch.cuda.synchronize()
