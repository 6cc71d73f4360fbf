"""Timing of the device CLIP text tower (ViT-B/32 text side, synthetic weights) next to the oracle on the host cores.
   python tools/gpu_clip_bench.py [B]   -> one JSON line"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clip_oracle as CO  # noqa: E402  (CPU baseline leg only)
from lsdm_b200 import synthetic as syn  # noqa: E402
from lsdm_b200.model.clip_text import ClipTextTower  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = syn.make_clip_state_dict(0)
tok = syn.make_clip_tokens(5, B)
L = int(tok.argmax(1).max()) + 1
res = {"B": B, "seq_len_computed": L}
for prec in ("3xtf32", "tf32", "fp32"):
    tower = ClipTextTower(precision=prec)
    tower.load_state_dict(sd, prefix="")
    tokd = tok.cuda()
    tower.encode_text(tokd, seq_len=L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = tower.encode_text(tokd, seq_len=L)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows = B * L
    flops = 2.0 * rows * 12 * (512 * 1536 + 512 * 512 + 2 * 512 * 2048)
    res[prec] = {"ms": ms, "dense_tflops": flops / (ms * 1e-3) / 1e12}
    if prec == "3xtf32":
        got = out.cpu()
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    CO.encode_text(sd, tok[:, :L])
    t0 = time.perf_counter()
    ref = CO.encode_text(sd, tok[:, :L])
    cpu_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    CO.encode_text(sd, tok)
    cpu77_ms = (time.perf_counter() - t0) * 1e3
res["cpu_oracle_ms_truncated"] = cpu_ms
res["cpu_oracle_ms_full_context"] = cpu77_ms
res["cpu_threads"] = os.cpu_count()
res["rel_l2_3xtf32_vs_oracle"] = float((got - ref).norm() / ref.norm())
print(json.dumps(res))
