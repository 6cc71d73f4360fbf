"""Per-tensor gradient error of lsdm_training_backward vs the autograd oracle (python tools/gpu_grad_diag.py [B])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import lsdm_oracle as O
from lsdm_b200 import synthetic as syn
from util import injected_rng, rel_l2
from test_gpu_training_backward import _train_model, _run_backward

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sd = syn.make_state_dict(0, "wellcond")
inp = syn.make_inputs(21, B, training=True)
fps, noise = syn.make_step_randoms(22, B, 1)
drop = syn.make_dropout_mask(23, B)
m, diff = _train_model()
terms = _run_backward(m, diff, inp, fps, noise, drop)
tables = O.diffusion_tables(O.cosine_betas(1000))
F64 = len(sys.argv) > 2 and sys.argv[2] == "f64"
cv = (lambda x: x.double()) if F64 else (lambda x: x)
loss, ref = O.training_grads(O._cast(sd, torch.float64) if F64 else sd, tables, cv(inp["x_start"]), cv(inp["mask"]), inp["t"], cv(inp["given_objs"]),
                             cv(inp["given_cats"]), cv(inp["target_cat"]), cv(inp["text_emb"]), list(fps[0]), cv(noise[0]), cv(drop))
print("loss", float(terms["loss"]), float(loss))
rows = []
for n, p in m.named_parameters():
    r = ref.get(n)
    if r is None:
        rows.append((0.0, n, "dead", 0.0 if p.grad is None else float(p.grad.abs().max())))
        continue
    rows.append((rel_l2(p.grad.cpu(), r), n, float(r.norm()), float(p.grad.norm())))
for e, n, a, b in sorted(rows, key=lambda r: -r[0] if isinstance(r[0], float) else 0)[:60]:
    print(f"{e:.3e}  {n:60s} ref {a}  got {b}")
