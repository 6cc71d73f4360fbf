"""ContactFormer temporal attention layer (reference contact_former/transformer.py:44-207).

CPU: the oracle restatement against golden outputs of the live reference classes (tests/golden/make_golden_cf.py).
GPU: the CUDA layer (C ABI lsdm_cf_*) through the reference-shaped modules against the golden values and the oracle.

Tolerances (relative L2): fp32 / 3xTF32 builds 2e-5; single-pass TF32 projections 2e-3 (2^-11 operand rounding through
q.k, softmax and two more projections).
"""
import numpy as np
import pytest
import torch

import cf_oracle as FO
from golden.make_golden_cf import CASES, cf_case, cf_state_dict
from util import golden, rel_l2


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    g = golden("cf_layer")
    sd = cf_state_dict(5)
    x, causal, rnd = cf_case(*CASES[name])
    with torch.no_grad():
        assert rel_l2(FO.mha(sd, x, None, "self_attn."), g[f"{name}_mha"]) < 2e-6
        assert rel_l2(FO.mha(sd, x, causal, "self_attn."), g[f"{name}_mha_causal"]) < 2e-6
        assert rel_l2(FO.mha(sd, x, rnd, "self_attn."), g[f"{name}_mha_rnd"]) < 2e-6
        assert rel_l2(FO.mha(sd, x, torch.zeros_like(rnd), "self_attn."), g[f"{name}_mha_allmasked"]) < 2e-6
        assert rel_l2(FO.ffn(sd, x, "pos_wise_ffnn."), g[f"{name}_ffn"]) < 2e-6
        assert rel_l2(FO.encoder_layer(sd, x), g[f"{name}_layer"]) < 2e-6
        assert rel_l2(FO.encoder_layer(sd, x, causal), g[f"{name}_layer_causal"]) < 2e-6


def test_mirror_modules_keep_the_reference_state_dict_keys():
    from lsdm_b200.contact_former.transformer import EncoderLayer

    layer = EncoderLayer(8, 64, 64, 64)
    assert set(layer.state_dict().keys()) == set(cf_state_dict(5).keys())
    layer.load_state_dict(cf_state_dict(5))  # strict
    with pytest.raises(NotImplementedError):
        layer(torch.zeros(1, 4, 2, 64))      # training mode: only the eval forward is implemented
    layer.eval()
    with pytest.raises(Exception):
        layer(torch.zeros(1, 4, 2, 64))      # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-5), ("fp32", 2e-5), ("tf32", 2e-3)])
def test_layer_vs_reference_golden(precision, tol):
    from lsdm_b200.contact_former.transformer import EncoderLayer

    g = golden("cf_layer")
    layer = EncoderLayer(8, 64, 64, 64, precision=precision)
    layer.load_state_dict(cf_state_dict(5))
    layer = layer.cuda().eval()
    for name in CASES:
        x, causal, rnd = cf_case(*CASES[name])
        xc = x.cuda()
        with torch.no_grad():
            assert rel_l2(layer.self_attn(xc).cpu(), g[f"{name}_mha"]) < tol
            assert rel_l2(layer.self_attn(xc, causal.cuda()).cpu(), g[f"{name}_mha_causal"]) < tol
            assert rel_l2(layer.self_attn(xc, rnd).cpu(), g[f"{name}_mha_rnd"]) < tol           # host mask is moved
            assert rel_l2(layer.self_attn(xc, torch.zeros_like(rnd)).cpu(), g[f"{name}_mha_allmasked"]) < tol
            assert rel_l2(layer.pos_wise_ffnn(xc).cpu(), g[f"{name}_ffn"]) < tol
            assert rel_l2(layer(xc).cpu(), g[f"{name}_layer"]) < tol
            assert rel_l2(layer(xc, causal.cuda()).cpu(), g[f"{name}_layer_causal"]) < tol
            assert torch.equal(xc.cpu(), x)  # the input is not modified


@pytest.mark.gpu
def test_layer_full_size_vs_oracle_and_masked_rows():
    """The reference's deployment shape: one sequence of 256 frames x 655 vertices (contact_former.py:271-275)."""
    from lsdm_b200.contact_former.transformer import EncoderLayer, MultiHeadAttention

    sd = cf_state_dict(9)
    layer = EncoderLayer(8, 64, 64, 64, precision="3xtf32")
    layer.load_state_dict(sd)
    layer = layer.cuda().eval()
    fast = EncoderLayer(8, 64, 64, 64)  # default precision: TF32 projections + tensor-core attention
    assert fast.self_attn.precision == "tf32"
    fast.load_state_dict(sd)
    fast = fast.cuda().eval()
    r = np.random.RandomState(1)
    x = torch.from_numpy(r.standard_normal((1, 256, 655, 64)).astype(np.float32))
    with torch.no_grad():
        got = layer(x.cuda()).cpu()
        got_fast = fast(x.cuda()).cpu()
        # oracle on a vertex subset (attention is independent per vertex; LayerNorm / FFN are per row)
        vs = [0, 1, 100, 333, 654]
        ref = FO.encoder_layer(sd, x[:, :, vs])
    assert torch.isfinite(got).all() and torch.isfinite(got_fast).all()
    assert rel_l2(got[:, :, vs], ref) < 2e-5
    assert rel_l2(got_fast[:, :, vs], ref) < 1e-3  # north_star's bound for what a path returns
    # a row whose keys are all masked is NaN in the reference (softmax over -inf); other rows are unaffected
    m = torch.ones(1, 256, 256)
    m[0, 7, :] = 0
    mha = layer.self_attn
    with torch.no_grad():
        gf = fast.self_attn(x[:, :, :4].cuda(), m).cpu()  # tensor-core kernel: same NaN row
    assert torch.isnan(gf[0, 7]).all() and torch.isfinite(gf[0, 8]).all()
    with torch.no_grad():
        gm = mha(x[:, :, :4].cuda(), m).cpu()
        rm = FO.mha(sd, x[:, :, :4], m, "self_attn.")
    assert torch.isnan(gm[0, 7]).all() and torch.isnan(rm[0, 7]).all()
    keep = [i for i in range(256) if i != 7]
    assert rel_l2(gm[0, keep], rm[0, keep]) < 2e-5
    # shape errors
    with pytest.raises(ValueError):
        mha(x[:, :, :4].cuda(), torch.ones(1, 255, 256))
    with pytest.raises(NotImplementedError):
        MultiHeadAttention(8, 64, 32, 64)


@pytest.mark.gpu
@pytest.mark.parametrize("S", [32, 96, 160, 256])
def test_tensor_core_attention_vs_oracle(S):
    """precision 'tf32' with seg_len % 32 == 0 takes the tcgen05 attention kernel (scores in tensor memory); same bound as
    the TF32 projections.  Checked with and without masks, against the oracle and against the CUDA-core kernel."""
    import ctypes as C
    from lsdm_b200 import _lib
    from lsdm_b200.contact_former.transformer import MultiHeadAttention

    sd = {k[len("self_attn."):]: v for k, v in cf_state_dict(7).items() if k.startswith("self_attn.")}
    mha = MultiHeadAttention(8, 64, 64, 64, precision="tf32")
    mha.load_state_dict(sd)
    mha = mha.cuda().eval()
    x, causal, rnd = cf_case(20 + S, 2, S, 7)
    lib = _lib.load()
    with torch.no_grad():
        for m in (None, causal, rnd, torch.zeros_like(rnd)):
            ref = FO.mha(sd, x, m)
            got = mha(x.cuda(), m).cpu()
            assert torch.isfinite(got).all()
            assert rel_l2(got, ref) < 2e-3
            _lib.check(lib.lsdm_cf_set_option(b"attn_tc", 0))
            try:
                simt = mha(x.cuda(), m).cpu()
            finally:
                _lib.check(lib.lsdm_cf_set_option(b"attn_tc", 1))
            assert rel_l2(got, simt) < 2e-3
            if m is None or bool((m != 0).any()):
                assert not torch.equal(got, simt)  # two different kernels really ran (all-masked: both give exactly fc(0) + x)
