"""Per-stage CUDA-vs-oracle error table (non-asserting diagnostic; run on the GPU box, output to gpurun_out/)."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch

import lsdm_oracle as O
from lsdm_b200 import synthetic as syn
from lsdm_b200.model.sdm import SceneDiffusionModel
from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd
from util import injected_rng, rel_l2


def main(kind="wellcond", B=3):
    print("device", torch.cuda.get_device_name(0), "kind", kind)
    sd = syn.make_state_dict(0, kind)
    inp = syn.make_inputs(1, B)
    fps, noise = syn.make_step_randoms(2, B, 1)
    t = torch.tensor([999, 500, 0][:B])
    tr = {}
    xo = inp["x_T"].clone()
    oc, x0o, go = O.forward(sd, xo, inp["mask"], t, inp["given_objs"], inp["given_cats"], inp["text_emb"], list(fps[0]), trace=tr)
    m = SceneDiffusionModel(**get_default_model_proxd())
    m.load_state_dict(sd)
    m.eval()
    g = {k: v.cuda() for k, v in inp.items()}
    x = g["x_T"].clone()
    with injected_rng(fps_starts=list(fps[0])):
        out_cat, x0 = m(x, g["mask"], t.cuda(), g["given_objs"], g["given_cats"], g["text_emb"])
    torch.cuda.synchronize()
    eng = m._engine
    C = B * 9
    for lvl, (name, npnt) in enumerate((("sa1", 1024), ("sa2", 256), ("sa3", 64), ("sa4", 16))):
        a = eng.debug_tensor(f"fps_idx{lvl}", torch.int32).view(C, npnt).cpu().long()
        print(f"fps_idx{lvl}: mismatches {(a != tr[name + '.fps_idx']).sum().item()} / {a.numel()}")
        a = eng.debug_tensor(f"ball_idx{lvl}", torch.int32).view(C, npnt, 32).cpu().long()
        print(f"ball_idx{lvl}: mismatches {(a != tr[name + '.group_idx']).sum().item()} / {a.numel()}")
    a = eng.debug_tensor("nn_idx0", torch.int32).view(C, 64, 3).cpu().long()
    print(f"nn_idx(fp4): mismatches {(a != tr['fp4.nn_idx']).sum().item()} / {a.numel()}")
    a = eng.debug_tensor("nn_idx3", torch.int32).view(C, 1024, 3).cpu().long()
    print(f"nn_idx(fp1): mismatches {(a != tr['fp1.nn_idx']).sum().item()} / {a.numel()}")
    pairs = [("enc", tr["enc"]), ("tr", tr["tr"]), ("attn_w", tr["attn_w"]), ("hm", tr["hm"]), ("l1_feat", tr["sa1.feat"]),
             ("l2_feat", tr["sa2.feat"]), ("l3_feat", tr["sa3.feat"]), ("l4_feat", tr["sa4.feat"]), ("fp4_feat", tr["fp4.feat"]),
             ("fp3_feat", tr["fp3.feat"]), ("fp2_feat", tr["fp2.feat"]), ("backbone", tr["backbone"]), ("pa", tr["pa"]),
             ("pw", tr["pw"]), ("pcd_out", tr["pcd_out"])]
    for name, ref in pairs:
        try:
            got = eng.debug_tensor(name).cpu().view(ref.shape)
            print(f"{name:10s} rel_l2 {rel_l2(got, ref):.3e}   |ref| {float(ref.norm()):.3e}  nan {int(torch.isnan(got).sum())}")
        except Exception:
            traceback.print_exc()
    emb = eng.debug_tensor("emb_cat").view(B * 1024, 256)[:, 128:].cpu().reshape(B, 1024, 128)
    print(f"emb        rel_l2 {rel_l2(emb, tr['emb']):.3e}")
    print(f"x mutated  rel_l2 {rel_l2(x.cpu(), xo):.3e}")
    print(f"out_cat    rel_l2 {rel_l2(out_cat.cpu(), oc):.3e}")
    print(f"x0         rel_l2 {rel_l2(x0.cpu(), x0o):.3e}")
    print(f"guiding    rel_l2 {rel_l2(m.saved_guiding_points.cpu(), go):.3e}")
    print("launches", eng.launch_count(), "workspace MB", eng.workspace_bytes / 2**20)


if __name__ == "__main__":
    for kind in ("wellcond", "default"):
        try:
            main(kind)
        except Exception:
            traceback.print_exc()
