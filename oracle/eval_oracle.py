"""CPU restatement of the reference's evaluation metrics -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file; the product path
(``lsdm_b200/util/evaluation.py`` -> ``liblsdm_b200.so``) never does.

Pinning: ``emd`` and ``accuracy`` are pinned against the live reference's ``util/evaluation.py`` (imported in the build
container with an ``open3d`` stub; ``tests/golden/make_golden_eval.py`` -> ``tests/golden/eval_metrics.npz``).
``calculate_fscore`` calls ``open3d.geometry.PointCloud.compute_point_cloud_distance`` -- open3d is absent here and not under
``/root/reference`` (requirements.txt, unpinned) -> *parity unpinned* for that third-party call; restated from its published
definition: for every point of the source cloud, the Euclidean distance to its nearest neighbour in the target cloud (double).
``chamfer``: pytorch3d 0.7.3 ``chamfer_distance`` default arguments (parity unpinned, see lsdm_oracle.py).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment
from scipy.spatial.distance import cdist


def emd(x, y):
    """util/evaluation.py:5-11: cdist (float64 Euclidean) + Hungarian; mean matched distance."""
    x, y = np.asarray(x), np.asarray(y)
    if x.ndim == 3:
        x, y = x[0], y[0]
    d = cdist(x, y)
    a = linear_sum_assignment(d)
    return d[a].sum() / min(len(x), len(y))


def nn_distance(src, dst):
    """open3d compute_point_cloud_distance(src -> dst): nearest-neighbour Euclidean distance of every src point, float64."""
    return cdist(np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)).min(1)


def calculate_fscore(gt, pr, th=0.1):
    """util/evaluation.py:28-52 (precision from gt->pr distances d1, recall from pr->gt distances d2, as written)."""
    d1 = nn_distance(gt, pr)
    d2 = nn_distance(pr, gt)
    if len(d1) and len(d2):
        recall = float((d2 < th).sum()) / float(len(d2))
        precision = float((d1 < th).sum()) / float(len(d1))
        fscore = 2 * recall * precision / (recall + precision) if recall + precision > 0 else 0
    else:
        fscore = precision = recall = 0
    return fscore, precision, recall


def accuracy(output, target, topk=(1,)):
    """util/evaluation.py:13-26: precision@k in percent.  ``output`` [B,C] scores, ``target`` [B] class ids."""
    output, target = np.asarray(output), np.asarray(target)
    order = np.argsort(-output, axis=1, kind="stable")
    res = []
    for k in topk:
        hit = (order[:, :k] == target[:, None]).any(1).sum()
        res.append(100.0 * hit / len(target))
    return res


def chamfer_per_sample(x, y):
    """pytorch3d chamfer_distance defaults per sample: mean_p min_q |p-q|^2 + mean_q min_p |p-q|^2."""
    out = []
    for a, b in zip(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)):
        d = cdist(a, b) ** 2
        out.append(d.min(1).mean() + d.min(0).mean())
    return np.asarray(out)
