"""One or more training steps (training_losses -> backward -> FusedAdamW) on synthetic data: python tools/gpu_train_step.py [B] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.optim import FusedAdamW
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = SceneDiffusionModel(**get_default_model_proxd())
m.load_state_dict(syn.make_state_dict(0, "wellcond"))
m.train()
diff = create_gaussian_diffusion(get_default_diffusion())
g = {k: v.cuda() for k, v in syn.make_inputs(5, B, training=True).items()}
opt = FusedAdamW(m.parameters(), lr=1e-3)
for it in range(K + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    terms = diff.training_losses(m, g["x_start"], g["mask"], g["t"], g["given_objs"], g["given_cats"], g["target_cat"], y=g["text_emb"])
    terms["loss"].backward()
    opt.step()
    torch.cuda.synchronize()
    print(f"step {it}: loss {float(terms['loss']):.5f}  {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
