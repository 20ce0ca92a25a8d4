"""The CPU-baseline port (oracle/torch_port.py) reproduces the reference's outputs and gradients (golden)."""
import os

import numpy as np
import pytest
import torch

import torch_port as TP
from conftest import GOLDEN


@pytest.mark.parametrize("name", ["small_a", "small_b", "small_c"])
def test_port_matches_reference_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    p = TP.make_params({k[2:]: v for k, v in g.items() if k.startswith("p.")}, torch.float64)
    feats = torch.tensor(g["x.feats"], dtype=torch.float64)
    loss, logits = TP.train_step(p, feats, torch.tensor(g["x.tokens"]), torch.tensor(g["x.lens"]), torch.tensor(g["x.labels"]))
    assert np.allclose(logits.detach().numpy(), g["f64.logits"], rtol=1e-10, atol=1e-12)
    assert abs(float(loss) - float(g["f64.loss"])) < 1e-12
    for k, v in p.items():
        if k.startswith("co_attention.W_b"):
            assert v.grad is None
            continue
        ref = g[f"f64.grad.{k}"]
        assert np.allclose(v.grad.numpy(), ref, rtol=1e-8, atol=1e-11), k
