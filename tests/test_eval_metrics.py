"""Evaluation metrics (reference util/evaluation.py, run/test_sdm.py:186-207).

CPU: the oracle restatement against golden outputs of the live reference (tests/golden/make_golden_eval.py).
GPU: the CUDA kernels (through the C ABI via lsdm_b200.util.evaluation) against the oracle and the same golden values.

Tolerances: EMD -- the auction's matching is within 2^-26 x extent per point of the optimum and the value is summed in
double, so 1e-6 relative (+1e-9 absolute for emd == 0) is asserted; F-score / accuracy are counts -> exact;
Chamfer is fp32 -> 1e-5 relative.
"""
import numpy as np
import pytest
import torch

import eval_oracle as EO
from golden.make_golden_eval import eval_inputs
from util import golden


def test_oracle_emd_and_accuracy_match_reference_golden():
    g = golden("eval_metrics")
    x, y, scores, target = eval_inputs()
    got = np.array([EO.emd(x[b], y[b]) for b in range(len(x))])
    np.testing.assert_allclose(got, g["emd"], rtol=1e-12, atol=1e-15)
    small = np.array([EO.emd(x[b, :k], y[b, :k]) for b, k in ((0, 1), (0, 2), (0, 33), (4, 257))])
    np.testing.assert_allclose(small, g["emd_small"], rtol=1e-12)
    np.testing.assert_allclose(EO.accuracy(scores, target, (1, 3, 5)), g["acc"], rtol=0, atol=1e-5)


def test_oracle_fscore_definition():
    # hand-checkable case: gt on a line, pr shifted by 0.05 with one far outlier
    gt = np.stack([np.arange(10) * 1.0, np.zeros(10), np.zeros(10)], 1)
    pr = gt + np.array([0.05, 0, 0])
    pr[0] = [100, 100, 100]
    f, p, r = EO.calculate_fscore(gt, pr, 0.1)
    assert p == 0.9 and r == 0.9 and abs(f - 0.9) < 1e-12  # gt[0] has no predicted point within 0.1; pr[0] is far from every gt
    assert EO.calculate_fscore(gt, gt + 5.0, 0.1) == (0, 0.0, 0.0)


@pytest.mark.gpu
def test_emd_vs_reference_golden_and_oracle():
    from lsdm_b200.util import evaluation as E

    g = golden("eval_metrics")
    x, y, _, _ = eval_inputs()
    got, assign = E.emd_batch(torch.from_numpy(x), torch.from_numpy(y), return_assignment=True)
    got = got.cpu().numpy()
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, g["emd"], rtol=1e-6, atol=1e-9)
    a = assign.cpu().numpy()
    for b in range(len(x)):  # a perfect matching, and the value is the mean matched distance
        assert sorted(a[b].tolist()) == list(range(x.shape[1]))
        d = np.linalg.norm(x[b].astype(np.float64) - y[b][a[b]].astype(np.float64), axis=1).mean()
        assert abs(d - got[b]) <= 1e-12 + 1e-12 * d
    # single-pair reference signature (util/evaluation.py:5), [1,n,3] and [n,3]
    assert abs(E.emd(torch.from_numpy(x[:1]), torch.from_numpy(y[:1])) - g["emd"][0]) <= 1e-6 * g["emd"][0]
    for (b, k), ref in zip(((0, 1), (0, 2), (0, 33), (4, 257)), g["emd_small"]):
        assert abs(E.emd(torch.from_numpy(x[b, :k]), torch.from_numpy(y[b, :k])) - ref) <= 1e-6 * ref


@pytest.mark.gpu
def test_emd_properties_full_batch():
    from lsdm_b200.util import evaluation as E

    r = np.random.RandomState(5)
    B, n = 64, 1024
    x = torch.from_numpy((r.rand(B, n, 3) - 0.5).astype(np.float32)).cuda()
    y = torch.from_numpy((r.randn(B, n, 3) * 0.25).astype(np.float32)).cuda()
    e_xy = E.emd_batch(x, y)
    e_yx = E.emd_batch(y, x)
    assert torch.isfinite(e_xy).all()
    # symmetric in its arguments (the optimum is; both runs are within the auction's bound of it)
    assert float((e_xy - e_yx).abs().max()) <= 1e-6 * float(e_xy.max())
    # a cloud against a permutation of itself matches at zero cost; translation by v costs exactly |v| (x -> x+v is optimal)
    perm = torch.stack([x[b][torch.from_numpy(r.permutation(n)).cuda()] for b in range(B)])
    assert float(E.emd_batch(x, perm).abs().max()) == 0.0
    v = torch.tensor([0.3, -0.4, 1.2], device="cuda")
    shifted = E.emd_batch(x, perm + v)
    assert float((shifted - float(v.double().norm())).abs().max()) <= 1e-5
    # deterministic
    assert torch.equal(e_xy, E.emd_batch(x, y))
    # never below the Chamfer-style lower bound (mean nearest-neighbour distance), never above the identity matching
    nn_lb = torch.cdist(x.double(), y.double()).min(2)[0].mean(1)
    ident = (x.double() - y.double()).norm(dim=2).mean(1)
    assert (e_xy >= nn_lb - 1e-9).all() and (e_xy <= ident + 1e-9).all()
    # spot-check 4 samples against the oracle (scipy Hungarian, ~0.1 s each)
    for b in (0, 17, 40, 63):
        ref = EO.emd(x[b].cpu().numpy(), y[b].cpu().numpy())
        assert abs(float(e_xy[b]) - ref) <= 1e-6 * ref


@pytest.mark.gpu
def test_emd_rejects_unequal_clouds():
    from lsdm_b200 import _lib
    from lsdm_b200.util import evaluation as E

    with pytest.raises(_lib.LsdmError):
        E.emd_batch(torch.zeros(1, 8, 3), torch.zeros(1, 9, 3))
    with pytest.raises(_lib.LsdmError):
        E.emd_batch(torch.zeros(1, 1025, 3), torch.zeros(1, 1025, 3))


@pytest.mark.gpu
def test_fscore_accuracy_chamfer_vs_oracle():
    from lsdm_b200.util import evaluation as E

    g = golden("eval_metrics")
    x, y, scores, target = eval_inputs()
    for th in (0.1, 0.03, 0.5):
        got = E.fscore_batch(torch.from_numpy(x), torch.from_numpy(y), th).cpu().numpy()
        ref = np.array([EO.calculate_fscore(x[b], y[b], th) for b in range(len(x))], dtype=np.float64)
        np.testing.assert_allclose(got, ref, rtol=1e-15, atol=0)
    # ragged sizes + the single-pair signature
    f = E.calculate_fscore(torch.from_numpy(x[0, :700]), torch.from_numpy(y[0, :333]), 0.1)
    np.testing.assert_allclose(f, EO.calculate_fscore(x[0, :700], y[0, :333], 0.1), rtol=1e-15)
    acc = [float(a) for a in E.accuracy(torch.from_numpy(scores), torch.from_numpy(target), topk=(1, 3, 5))]
    np.testing.assert_allclose(acc, g["acc"], rtol=0, atol=1e-4)
    ch = E.chamfer_batch(torch.from_numpy(x), torch.from_numpy(y)).cpu().numpy()
    np.testing.assert_allclose(ch, EO.chamfer_per_sample(x, y), rtol=1e-5, atol=1e-9)
    loss, normals = E.chamfer_distance(torch.from_numpy(x), torch.from_numpy(y))
    assert normals is None and abs(float(loss) - EO.chamfer_per_sample(x, y).mean()) <= 1e-5 * float(loss)


def _auction_model(x, y, eps0=1 << 27, eps_final=16, theta=8):
    """numpy model of eval_metrics.cu::emd_auction_kernel: same integer scaling, epsilon schedule, Jacobi rounds and
    (bid << 10 | 1023 - person) conflict resolution -- a CPU check of the ALGORITHM (the kernel itself is tested on the GPU)."""
    n = len(x)
    ext = max(np.abs(x).max(), np.abs(y).max())
    _, e2 = np.frexp(np.float32(ext))
    scale = np.float32(2.0 ** (28 - int(e2)))
    d = np.sqrt(((x[:, None, :] - y[None, :, :]).astype(np.float32) ** 2).sum(-1, dtype=np.float32)).astype(np.float32)
    c = np.rint((d * scale).astype(np.float64)).astype(np.int64)
    assert c.max() < 2 ** 30
    price = np.zeros(n, np.int64)
    eps, rounds = eps0, 0
    while True:
        owner = -np.ones(n, np.int64)
        assigned = -np.ones(n, np.int64)
        while True:
            L = np.nonzero(assigned < 0)[0]
            if len(L) == 0:
                break
            rounds += 1
            v = -c[L] - price[None, :]
            j1 = np.argmax(v, 1)
            v1 = v[np.arange(len(L)), j1]
            v[np.arange(len(L)), j1] = np.iinfo(np.int64).min
            v2 = v.max(1) if n > 1 else v1
            bid = price[j1] + (v1 - v2 if n > 1 else 0) + eps
            key = (bid << 10) | (1023 - L)
            best = {}
            for k, j in zip(key, j1):
                best[j] = max(best.get(j, 0), int(k))
            for j, k in best.items():
                w = 1023 - (k & 1023)
                if owner[j] >= 0:
                    assigned[owner[j]] = -1
                owner[j], assigned[w], price[j] = w, j, k >> 10
        if eps <= eps_final:
            break
        eps = max(eps // theta, eps_final)
    assert price.max() < 2 ** 53  # the 64-bit bid key has room
    return np.sqrt(((x.astype(np.float64) - y[assigned].astype(np.float64)) ** 2).sum(-1)).sum() / n, rounds


@pytest.mark.parametrize("n", [1, 2, 17, 96])
def test_auction_algorithm_model_matches_hungarian(n):
    r = np.random.RandomState(n)
    for x, y in (((r.rand(n, 3) - 0.5), (r.rand(n, 3) - 0.5)),                       # uniform vs uniform
                 ((r.randn(n, 3) * 0.01), (r.rand(n, 3) - 0.5)),                      # collapsed vs spread
                 (np.round((r.rand(n, 3) - 0.5) * 4) / 4, np.round((r.rand(n, 3) - 0.5) * 4) / 4),   # duplicates / ties
                 (np.zeros((n, 3)), (r.rand(n, 3) - 0.5) * 100.0)):                   # identical persons, large extent
        x, y = x.astype(np.float32), y.astype(np.float32)
        got, rounds = _auction_model(x, y)
        ref = EO.emd(x, y)
        assert abs(got - ref) <= 1e-6 * max(ref, 1e-12) + 1e-9, (n, got, ref)
        assert rounds < (1 << 20)
