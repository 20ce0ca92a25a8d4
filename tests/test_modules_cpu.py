"""CPU-only checks of the host side: drop-in surface (names, signatures, state_dict keys), the C ABI
library loads and exports every symbol the header declares, and the ops refuse to run without CUDA."""
import inspect
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


def test_state_dict_keys_match_reference(pkg):
    want = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    net = pkg.HieCoAttnHotPath(10000, 512, 1001, 1024)
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    ref = {k: v for k, v in want.items() if not k.startswith("baseline.")}
    assert got == ref
    bq = pkg.QuestionBaselineEncoder(10000, 300, 1024)
    assert {f"baseline.question_encoder.{k}": list(v.shape) for k, v in bq.state_dict().items()} == \
           {k: v for k, v in want.items() if k.startswith("baseline.")}


def test_signatures_match_reference(pkg):
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(pkg.HierarchicalCoAttentionNet.__init__) == ["self", "ques_enc_params", "img_enc_params", "K", "mlp_dim"]
    assert sig(pkg.HierarchicalCoAttentionNet.forward) == ["self", "x_img", "x_ques", "x_ques_lens"]
    assert sig(pkg.QuestionCoAttentionEncoder.__init__) == ["self", "vocab_size", "word_emb_dim", "hidden_dim"]
    assert sig(pkg.QuestionCoAttentionEncoder.forward) == ["self", "x", "x_lens"]
    assert sig(pkg.PhraseConvPool.__init__) == ["self", "emb_dim"]
    assert sig(pkg.PhraseConvPool.forward)[:2] == ["self", "x_question"]
    assert sig(pkg.ParallelCoAttention.__init__) == ["self", "hidden_dim"]
    assert sig(pkg.ParallelCoAttention.forward) == ["self", "x_img", "x_ques_hierarchy"]
    assert sig(pkg.MLPClassifier.__init__) == ["self", "hidden_dim", "mlp_dim", "K"]
    assert sig(pkg.MLPClassifier.forward) == ["self", "x_img_feats", "x_ques_feats"]
    assert sig(pkg.VQABaselineNet.__init__) == ["self", "ques_enc_params", "img_enc_params", "K"]


def test_model_shim_exports_what_main_imports():
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    m = importlib.import_module("model")
    assert hasattr(m, "VQABaselineNet") and hasattr(m, "HierarchicalCoAttentionNet")       # reference main.py:15


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "hiecoattn_b200.h")).read()
    declared = set(re.findall(r"HCA_API[^;(]*?\b(hca_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(pkg._lib.SIGNATURES), declared ^ set(pkg._lib.SIGNATURES)
    lib = pkg._lib.lib()                                   # loads the .so; no CUDA call is made
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hca_abi_version() == pkg._lib.ABI_VERSION
    assert lib.hca_phrase_conv_pool_workspace(160, 26, 512) > 0     # pure host arithmetic
    assert pkg._lib.get_option("gemm") in ("tc", "ffma")


def test_no_cpu_fallback(pkg):
    net = pkg.HieCoAttnHotPath(50, 32, 7, 16)
    with pytest.raises((NotImplementedError, RuntimeError)):
        net(torch.zeros(2, 5, 32), torch.ones(2, 3, dtype=torch.long), torch.tensor([3, 2]))
    with pytest.raises((NotImplementedError, RuntimeError)):
        pkg.ops.mlp(torch.zeros(3, 2, 32), torch.zeros(3, 2, 32), *[torch.zeros(1)] * 8)


def test_product_path_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "visual-question-answering_b200")
    for f in os.listdir(pkg_dir):
        if f.endswith(".py"):
            src = open(os.path.join(pkg_dir, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f
            assert "torch_port" not in src and "hiecoattn_oracle" not in src, f


def test_synthetic_inputs_follow_the_reference_conventions(syn):
    x = syn.make_inputs(32, T=26, vocab=100, K=11, min_len=1)
    lens, tok = x["lens"], x["tokens"]
    assert (lens[:-1] >= lens[1:]).all() and lens.min() >= 1 and lens.max() <= 26        # sort_batch, utils.py:33-45
    for b in range(32):
        assert (tok[b, :lens[b]] >= 1).all() and (tok[b, lens[b]:] == 0).all()           # <PAD>=0, utils.py:18-30,106
    assert x["labels"].min() >= 0 and x["labels"].max() < 11
    p = syn.make_params(32, 100, 11, 16)
    assert (p["question_encoder.word_embedding.weight"][0] == 0).all()                  # padding_idx row, model.py:263
