"""CLIP text tower (reference model/sdm.py:245-277 -> openai/CLIP ``encode_text``).

CPU: the oracle restatement (oracle/clip_oracle.py) is pinned against transformers' CLIPTextModelWithProjection -- an
independent implementation of the same published model -- with random weights mapped key by key (the ``clip`` package
itself is absent: see the oracle's header).  GPU: the CUDA tower (C ABI ``lsdm_clip_*``) against the oracle.

Tolerances (relative L2 of the [B,512] embedding): fp32 CUDA-core and 3xTF32 tensor-core builds 2e-5 (summation order /
2^-22 operand residuals over 12 blocks); single-pass TF32 5e-3 (2^-11 operand rounding through 48 dense layers; offered,
not the default).  north_star's bound for what the path returns is 1e-3.
"""
import numpy as np
import pytest
import torch

import clip_oracle as CO
from lsdm_b200 import synthetic as syn
from util import rel_l2


def _hf(layers, vocab):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection

    torch.manual_seed(0)
    cfg = CLIPTextConfig(vocab_size=vocab, num_hidden_layers=layers, eos_token_id=vocab - 1, bos_token_id=vocab - 2, pad_token_id=0)
    m = CLIPTextModelWithProjection(cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.normal_(0, 0.04)
            if "norm" in n and n.endswith("weight"):
                p.add_(1.0)
    return m


@pytest.mark.parametrize("layers,vocab", [(2, 512), (12, 2048)])
def test_oracle_matches_transformers_clip(layers, vocab):
    m = _hf(layers, vocab)
    sd = CO.from_hf(m.state_dict())
    tok = syn.make_clip_tokens(3, 6, vocab=vocab)
    with torch.no_grad():
        ref = m(input_ids=tok.long()).text_embeds
        got = CO.encode_text(sd, tok)
    assert rel_l2(got, ref) < 5e-6
    # causal attention: positions after the EOT token do not matter (the property the device path's seq_len relies on)
    with torch.no_grad():
        assert rel_l2(CO.encode_text(sd, tok[:, :22]), got) < 5e-6
        tok2 = tok.clone()
        for b in range(len(tok2)):
            e = int(tok2[b].argmax())
            tok2[b, e + 1:] = torch.randint(1, vocab - 2, (tok2.shape[1] - e - 1,), dtype=tok2.dtype)
        assert rel_l2(CO.encode_text(sd, tok2), got) < 5e-6


def test_synthetic_clip_state_dict_has_the_openai_keys():
    sd = syn.make_clip_state_dict(0, layers=2, vocab=64, prefix="clip_model.")
    assert sd["clip_model.text_projection"].shape == (512, 512) and sd["clip_model.positional_embedding"].shape == (77, 512)
    assert sd["clip_model.transformer.resblocks.1.attn.in_proj_weight"].shape == (1536, 512)
    assert sd["clip_model.transformer.resblocks.0.mlp.c_proj.weight"].shape == (512, 2048)
    assert len(sd) == 5 + 2 * 12
    tok = syn.make_clip_tokens(1, 9, vocab=64)
    assert tok.shape == (9, 77) and int(tok[0].argmax()) == 21 and (tok.argmax(1) <= 21).all() and (tok[:, 22:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-5), ("fp32", 2e-5), ("tf32", 5e-3)])
def test_tower_vs_oracle(precision, tol):
    from lsdm_b200.model.clip_text import ClipTextTower

    sd = syn.make_clip_state_dict(1, vocab=4096, prefix="clip_model.")
    tok = syn.make_clip_tokens(2, 7, vocab=4096)
    tower = ClipTextTower(precision=precision)
    assert tower.load_state_dict(sd) == len(sd)
    assert tower.dims == {"width": 512, "layers": 12, "heads": 8, "ctx": 77, "vocab": 4096, "embed": 512}
    n0 = tower.launch_count()
    got = tower.encode_text(tok)
    assert tower.launch_count() - n0 == 2 + 12 * 7  # embed+LN, (qkv, attn, out_proj, add+LN, c_fc, c_proj, add+LN) x 12, pool+project
    with torch.no_grad():
        ref = CO.encode_text({k[len("clip_model."):]: v for k, v in sd.items()}, tok)
    assert torch.isfinite(got).all()
    assert rel_l2(got.cpu(), ref) < tol
    # seq_len only bounds the computed positions: every value is bit-identical to the full 77-token context
    assert torch.equal(tower.encode_text(tok, seq_len=77), got)
    assert torch.equal(tower.encode_text(tok.cuda().long()), got)
    # a sample whose EOT lies beyond seq_len is flagged, the others are unaffected
    short = tower.encode_text(tok, seq_len=10)
    late = (tok.argmax(1) >= 10)
    assert late.any() and (~late).any()
    assert torch.isnan(short[late.cuda()]).all() and torch.equal(short[(~late).cuda()], got[(~late).cuda()])


@pytest.mark.gpu
def test_tower_vit_b32_shapes_batch64():
    from lsdm_b200.model.clip_text import ClipTextTower

    sd = syn.make_clip_state_dict(0)  # full ViT-B/32 text side: vocabulary 49408, 63 M parameters
    tok = syn.make_clip_tokens(5, 64)
    tower = ClipTextTower()
    tower.load_state_dict(sd, prefix="")
    got = tower.encode_text(tok)
    with torch.no_grad():
        ref = CO.encode_text(sd, tok[:, :22])
    assert rel_l2(got.cpu(), ref) < 2e-5
    # single sentence, shortest possible (SOT, EOT)
    one = torch.zeros(1, 77, dtype=torch.int64)
    one[0, 0], one[0, 1] = 49406, 49407
    with torch.no_grad():
        assert rel_l2(tower.encode_text(one).cpu(), CO.encode_text(sd, one)) < 2e-5


@pytest.mark.gpu
def test_tower_error_paths():
    from lsdm_b200 import _lib
    from lsdm_b200.model.clip_text import ClipTextTower

    sd = syn.make_clip_state_dict(1, layers=2, vocab=256)
    bad = {k: v for k, v in sd.items() if k != "transformer.resblocks.1.mlp.c_fc.bias"}
    with pytest.raises(_lib.LsdmError):
        ClipTextTower().load_state_dict(bad, prefix="")
    t = ClipTextTower()
    with pytest.raises(_lib.LsdmError):
        t.encode_text(torch.zeros(1, 77, dtype=torch.int64))  # no weights
    t.load_state_dict(sd, prefix="")
    with pytest.raises(ValueError):
        t.encode_text(torch.zeros(1, 76, dtype=torch.int64))
    with pytest.raises(ValueError):
        t.encode_text(torch.zeros(1, 77))
    with pytest.raises(_lib.LsdmError):
        t.encode_text(torch.zeros(1, 77, dtype=torch.int64), seq_len=78)


@pytest.mark.gpu
def test_model_accepts_reference_checkpoint_with_clip_and_token_ids():
    """A reference checkpoint carries clip_model.* (model/sdm.py:231); y may then be clip.tokenize's output."""
    from lsdm_b200.model.sdm import SceneDiffusionModel
    from lsdm_b200.util.model_util import get_default_model_proxd
    import lsdm_oracle as O
    from util import injected_rng

    B = 2
    sd = syn.make_state_dict(0, "wellcond")
    clip_sd = syn.make_clip_state_dict(3, vocab=1024, prefix="clip_model.")
    model = SceneDiffusionModel(**get_default_model_proxd())
    model.load_state_dict({**sd, **clip_sd})
    model.eval()
    assert model.clip_text.dims["layers"] == 12
    inp = syn.make_inputs(1, B)
    fps, _ = syn.make_step_randoms(2, B, 1)
    tok = syn.make_clip_tokens(4, B, vocab=1024)
    t = torch.tensor([999, 3])
    with torch.no_grad():
        emb = CO.encode_text({k[len("clip_model."):]: v for k, v in clip_sd.items()}, tok)
        xo = inp["x_T"].clone()
        oc, x0o, _ = O.forward(sd, xo, inp["mask"], t, inp["given_objs"], inp["given_cats"], emb, list(fps[0]))
    x = inp["x_T"].clone().cuda()
    with injected_rng(fps_starts=list(fps[0])), torch.no_grad():
        out_cat, x0 = model(x, inp["mask"].cuda(), t.cuda(), inp["given_objs"].cuda(), inp["given_cats"].cuda(), y=tok)
    assert rel_l2(x0.cpu(), x0o) < 1e-3 and rel_l2(out_cat.cpu(), oc) < 1e-3
    # strings: tokenizer hook (stands in for clip.tokenize) -> same path
    words = {"a": 5, "chair": 9}
    def tokenize(texts, context_length=77, truncate=True):
        out = torch.zeros(len(texts), context_length, dtype=torch.int64)
        for i, s in enumerate(texts):
            ids = [1022] + [words[w] for w in s.split()] + [1023]
            out[i, :len(ids)] = torch.tensor(ids)
        return out
    model.set_tokenizer(tokenize)
    e1 = model._encode_text(["a chair", "chair"])
    e2 = model.clip_text.encode_text(torch.cat([tokenize(["a chair", "chair"], 22), torch.zeros(2, 55, dtype=torch.int64)], 1))
    assert torch.equal(e1, e2)


@pytest.mark.gpu
def test_strings_run_natively_through_the_bpe_tokenizer_and_the_device_tower():
    """y: tuple of strings as run/test_sdm.py passes it -> lsdm_b200's own BPE tokeniser (restatement of clip.tokenize, here with a
    toy merge table: the clip vocabulary file is not available offline) -> device CLIP tower -> the [B,512] embedding the oracle
    tower computes from the same token ids."""
    from lsdm_b200.model.clip_tokenizer import ClipBpeTokenizer
    from lsdm_b200.model.sdm import SceneDiffusionModel
    from lsdm_b200.util.model_util import get_default_model_proxd

    merges = [("c", "h"), ("ch", "a"), ("i", "r</w>"), ("cha", "ir</w>"), ("t", "h"), ("th", "e</w>"), ("s", "o"), ("f", "a</w>")]
    tok = ClipBpeTokenizer(merges=merges)
    vocab = tok.eot + 1
    clip_sd = syn.make_clip_state_dict(5, vocab=vocab, prefix="clip_model.")
    model = SceneDiffusionModel(**get_default_model_proxd())
    model.load_state_dict({**syn.make_state_dict(0, "wellcond"), **clip_sd})
    model.eval()
    model._tokenizer = tok
    texts = ("the chair", "a sofa next to the chair, facing the tv")
    ids = tok.tokenize(list(texts), context_length=22, truncate=True)
    assert ids[0, 0] == tok.sot and (ids == tok.eot).sum(1).tolist() == [1, 1]
    e = model._encode_text(texts)
    with torch.no_grad():
        ref = CO.encode_text({k[len("clip_model."):]: v for k, v in clip_sd.items()}, torch.cat([ids, torch.zeros(2, 55, dtype=torch.long)], 1))
    assert rel_l2(e.cpu(), ref) < 1e-4
