"""Timing of the device evaluation metrics next to the reference's host implementation (scipy Hungarian).
   python tools/gpu_eval_bench.py [B]   -> one JSON line"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import eval_oracle as EO  # noqa: E402  (CPU baseline leg only)
from lsdm_b200 import _lib  # noqa: E402
from lsdm_b200.engine import _ptr, _stream  # noqa: E402
from lsdm_b200.util import evaluation as E  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
r = np.random.RandomState(3)
x = torch.from_numpy((r.rand(B, 1024, 3) - 0.5).astype(np.float32)).cuda()
y = torch.from_numpy((r.randn(B, 1024, 3) * 0.25).astype(np.float32)).cuda()
dev = x.device
out = torch.empty(B, dtype=torch.float64, device=dev)
rounds = torch.empty(B, dtype=torch.int32, device=dev)
lib = _lib.load()


def run():
    _lib.check(lib.lsdm_eval_emd(_ptr(x), _ptr(y), B, 1024, 1024, _ptr(out), None, _ptr(rounds), _stream(dev)))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
e0.record()
for _ in range(10):
    E.fscore_batch(x, y, 0.1)
    E.chamfer_batch(x, y)
e1.record()
torch.cuda.synchronize()
ms_fc = e0.elapsed_time(e1) / 10
xc, yc = x[:4].cpu().numpy(), y[:4].cpu().numpy()
t0 = time.perf_counter()
ref = [EO.emd(xc[b], yc[b]) for b in range(4)]
cpu_ms = (time.perf_counter() - t0) / 4 * 1e3
err = max(abs(float(out[b]) - ref[b]) / ref[b] for b in range(4))
print(json.dumps({"B": B, "emd_ms_batch": ms, "emd_ms_per_sample": ms / B, "rounds_mean": float(rounds.float().mean()),
                  "rounds_max": int(rounds.max()), "fscore_plus_chamfer_ms_batch": ms_fc, "cpu_scipy_emd_ms_per_sample": cpu_ms,
                  "emd_speedup_vs_1core_scipy": cpu_ms / (ms / B), "max_rel_err_vs_scipy": err}))
