"""Timing of the ContactFormer attention layer (1 x 256 frames x 655 vertices x 64, 8 heads) next to the oracle on the host.
   python tools/gpu_cf_bench.py   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cf_oracle as FO  # noqa: E402  (CPU baseline leg only)
from golden.make_golden_cf import cf_state_dict  # noqa: E402
from lsdm_b200.contact_former.transformer import EncoderLayer  # noqa: E402

bs, S, V, H = 1, 256, 655, 8
sd = cf_state_dict(9)
x = torch.from_numpy(np.random.RandomState(1).standard_normal((bs, S, V, 64)).astype(np.float32))
xc = x.cuda()
rows = bs * S * V
flops_mha = 2.0 * rows * (3 * 64 * H * 64 + H * 64 * 64) + 2.0 * 2 * bs * V * H * S * S * 64
flops_ffn = 2.0 * rows * 2 * 64 * 64
res = {"shape": [bs, S, V, 64], "heads": H, "gflop_mha": flops_mha / 1e9, "gflop_ffn": flops_ffn / 1e9}
for prec in ("3xtf32", "tf32", "fp32"):
    layer = EncoderLayer(H, 64, 64, 64, precision=prec)
    layer.load_state_dict(sd)
    layer = layer.cuda().eval()
    with torch.no_grad():
        layer(xc)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(5):
            y = layer.self_attn(xc)
        e[1].record()
        for _ in range(5):
            z = layer.pos_wise_ffnn(y)
        e[2].record()
        torch.cuda.synchronize()
    ms_mha, ms_ffn = e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5
    res[prec] = {"mha_ms": ms_mha, "mha_tflops": flops_mha / ms_mha / 1e9, "ffn_ms": ms_ffn,
                 "layer_io_gbs": 2 * rows * 64 * 4 * 2 / ((ms_mha + ms_ffn) * 1e-3) / 1e9}
    res[prec]["_out"] = z.cpu()
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    t0 = time.perf_counter()
    ref = FO.encoder_layer(sd, x)
    res["cpu_oracle_layer_ms"] = (time.perf_counter() - t0) * 1e3
res["cpu_threads"] = os.cpu_count()
for prec in ("3xtf32", "tf32", "fp32"):
    got = res[prec].pop("_out")
    res[prec]["rel_l2_vs_oracle"] = float((got - ref).norm() / ref.norm())
print(json.dumps(res))
