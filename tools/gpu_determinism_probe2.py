"""Which intermediate differs first between two identical runs of the pipelined STRICT loop (debug taps of the last step)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsdm_b200 import synthetic as syn  # noqa: E402
from lsdm_b200.model.sdm import SceneDiffusionModel  # noqa: E402
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd  # noqa: E402

B, K = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
INT = [f"fps_idx{l}" for l in range(4)] + [f"ball_idx{l}" for l in range(4)] + [f"nn_idx{l}" for l in range(4)]
FLT = [f"nn_w{l}" for l in range(4)] + ["hm", "attn_w", "tr", "enc", "l1_feat", "l2_feat", "l3_feat", "l4_feat", "fp4_feat", "fp3_feat", "fp2_feat",
                                        "backbone", "pa", "pw", "pcd_out", "emb_cat", "x0", "guiding"]


def run(flag, k):
    eng.set_option("select_uniform", flag)
    x = g["x_T"].clone()
    eng.sample_loop(x, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_d[:k], noise_d[:k], 999, False)
    torch.cuda.synchronize()
    taps = {n: eng.debug_tensor(n, torch.int32).clone() for n in INT}
    taps.update({n: eng.debug_tensor(n, torch.float32).clone() for n in FLT})
    taps["x"] = x
    return taps


for flag in (0, 1):
    for k in (1, 2, 3, K):
        a, b = run(flag, k), run(flag, k)
        bad = [(n, float((a[n].double() - b[n].double()).abs().max())) for n in a if not torch.equal(a[n], b[n])]
        print(f"select_uniform={flag} steps={k}: differing taps:", bad if bad else "none")
eng.set_option("select_uniform", 1)
