"""Golden vectors for the evaluation metrics from the UNMODIFIED reference ``util/evaluation.py`` (build container only).

    python tests/golden/make_golden_eval.py      ->  tests/golden/eval_metrics.npz

``open3d`` is absent: it is stubbed so that the module imports; ``calculate_fscore`` (the only open3d user) is NOT
recorded (parity unpinned, see oracle/eval_oracle.py).  ``emd`` and ``accuracy`` run as written.  Inputs are regenerated
from seeds by ``eval_inputs`` below, so the fixture holds reference OUTPUTS only.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("LSDM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def eval_inputs(seed=77, B=6, n=1024, C=13):
    """Cloud pairs of several shapes (uniform, blob vs spread, near-identical, duplicated grid points) + class scores."""
    r = np.random.RandomState(seed)
    x = (r.rand(B, n, 3) - 0.5).astype(np.float32)
    y = (r.rand(B, n, 3) - 0.5).astype(np.float32)
    x[1] = (r.randn(n, 3) * 0.02).astype(np.float32)                       # collapsed prediction vs spread target
    y[2] = x[2] + (r.randn(n, 3) * 1e-3).astype(np.float32)                # near-identical
    x[3] = np.round(x[3] * 8) / 8                                          # many duplicate points / tied distances
    y[3] = np.round(y[3] * 8) / 8
    x[4] = (r.randn(n, 3) * 0.3).astype(np.float32)                        # gaussian vs uniform
    y[5] = x[5][r.permutation(n)]                                          # exact permutation: emd == 0
    scores = r.randn(64, C).astype(np.float32)
    target = r.randint(0, C, size=(64,)).astype(np.int64)
    return x, y, scores, target


def main():
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    sys.path.insert(0, REF)
    from util import evaluation as ref_eval  # the reference's own file

    x, y, scores, target = eval_inputs()
    emds = np.array([ref_eval.emd(torch.from_numpy(x[b:b + 1]), torch.from_numpy(y[b:b + 1])) for b in range(len(x))])
    small = np.array([ref_eval.emd(torch.from_numpy(x[b, :k]), torch.from_numpy(y[b, :k])) for b, k in ((0, 1), (0, 2), (0, 33), (4, 257))])
    acc = [float(v) for v in ref_eval.accuracy(torch.from_numpy(scores), torch.from_numpy(target), topk=(1, 3, 5))]
    np.savez(os.path.join(HERE, "eval_metrics.npz"), emd=emds, emd_small=small, acc=np.array(acc))
    print("emd", emds, "small", small, "acc", acc)


if __name__ == "__main__":
    main()
