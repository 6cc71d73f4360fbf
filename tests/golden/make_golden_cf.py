"""Golden vectors for the ContactFormer attention layer from the UNMODIFIED reference classes (build container only).

    python tests/golden/make_golden_cf.py      ->  tests/golden/cf_layer.npz

``contact_former/transformer.py`` is pure torch / numpy, so it is imported as is.  Weights and inputs are regenerated
from seeds by ``cf_case`` below (the fixture holds reference OUTPUTS only).
"""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("LSDM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def cf_state_dict(seed, n_head=8, d_hid=64):
    """EncoderLayer state dict (reference key names), seeded; LayerNorm affine randomised so that it matters."""
    r = np.random.RandomState(seed)

    def t(shape, std, mean=0.0):
        return torch.from_numpy((r.standard_normal(shape) * std + mean).astype(np.float32))

    sd = {}
    for n in ("w_q", "w_k", "w_v"):
        sd[f"self_attn.{n}.weight"], sd[f"self_attn.{n}.bias"] = t((n_head * 64, 64), 0.125), t((n_head * 64,), 0.05)
    sd["self_attn.fc.weight"], sd["self_attn.fc.bias"] = t((64, n_head * 64), 0.06), t((64,), 0.05)
    sd["self_attn.layer_norm.weight"], sd["self_attn.layer_norm.bias"] = t((64,), 0.1, 1.0), t((64,), 0.05)
    sd["pos_wise_ffnn.w_1.weight"], sd["pos_wise_ffnn.w_1.bias"] = t((d_hid, 64, 1), 0.15), t((d_hid,), 0.05)
    sd["pos_wise_ffnn.w_2.weight"], sd["pos_wise_ffnn.w_2.bias"] = t((64, d_hid, 1), 0.15), t((64,), 0.05)
    sd["pos_wise_ffnn.layer_norm.weight"], sd["pos_wise_ffnn.layer_norm.bias"] = t((64,), 0.1, 1.0), t((64,), 0.05)
    return sd


def cf_case(seed, bs, S, V):
    r = np.random.RandomState(seed)
    x = torch.from_numpy(r.standard_normal((bs, S, V, 64)).astype(np.float32))
    causal = torch.tril(torch.ones(S, S)).unsqueeze(0).expand(bs, -1, -1).contiguous()   # keep j <= i
    rnd = torch.from_numpy((r.rand(bs, S, S) < 0.6).astype(np.float32))
    rnd[:, torch.arange(S), torch.arange(S)] = 1.0                                       # no fully masked row
    return x, causal, rnd


CASES = {"a": (11, 2, 16, 5), "b": (12, 1, 37, 3)}


def main():
    spec = importlib.util.spec_from_file_location("ref_cf_transformer", os.path.join(REF, "contact_former", "transformer.py"))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    out = {}
    sd = cf_state_dict(5)
    layer = T.EncoderLayer(8, 64, 64, 64).eval()
    layer.load_state_dict(sd)
    for name, (seed, bs, S, V) in CASES.items():
        x, causal, rnd = cf_case(seed, bs, S, V)
        with torch.no_grad():
            out[f"{name}_mha"] = layer.self_attn(x).numpy()
            out[f"{name}_mha_causal"] = layer.self_attn(x, causal).numpy()
            out[f"{name}_mha_rnd"] = layer.self_attn(x, rnd).numpy()
            out[f"{name}_mha_allmasked"] = layer.self_attn(x, torch.zeros_like(rnd)).numpy()
            out[f"{name}_ffn"] = layer.pos_wise_ffnn(x).numpy()
            out[f"{name}_layer"] = layer(x).numpy()
            out[f"{name}_layer_causal"] = layer(x, causal).numpy()
    np.savez_compressed(os.path.join(HERE, "cf_layer.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
