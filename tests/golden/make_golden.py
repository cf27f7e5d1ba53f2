"""Regenerates the fixtures under tests/golden/.

1. rmse_known_answers.json — the ONLY known-answer vectors the reference's own tests hold for this path
   (/root/reference/hem/ops/test_losses.py:7-27: constant (1,64,64,3) tensors and the expected hem.rmse).
   Transcribed from that test (TensorFlow itself cannot run here).
2. oracle_iwgan_tiny.json — outputs of the CPU oracle on a tiny fixed IWGAN problem, committed so that a
   change to the oracle (the checker) is itself caught by the CPU suite.  These pin the oracle against
   itself, not against TensorFlow: parity with TF stays UNPINNED (oracle/__init__.py).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys
from collections import OrderedDict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import models as OM  # noqa: E402


def rmse_fixture():
    cases = [
        {"x": 1.0, "x_hat": 1.0, "rmse": 0.0},    # test_losses.py:15-17
        {"x": 1.0, "x_hat": 0.0, "rmse": 1.0},    # test_losses.py:19-21
        {"x": -1.0, "x_hat": 1.0, "rmse": 2.0},   # test_losses.py:23-25
        {"x": 1.0, "x_hat": -1.0, "rmse": 2.0},   # test_losses.py:26-27
    ]
    return {"source": "algoterranean/3dgan hem/ops/test_losses.py:7-27", "shape": [1, 64, 64, 3], "cases": cases}


def oracle_fixture():
    H, C, L, B = 32, 3, 4, 2
    gs, ds = OM.gan_param_specs("iwgan", H, C, L)
    p = OM.init_params(OrderedDict(list(gs.items()) + list(ds.items())), 0, torch.float64)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, H, H, C, generator=g, dtype=torch.float64)
    z = torch.randn(B, L, generator=g, dtype=torch.float64)
    a = torch.rand(B, 1, generator=g, dtype=torch.float64)
    r = OM.gan_grads(p, x, z, a, "iwgan", H, C, L)
    return {"config": {"H": H, "C": C, "L": L, "B": B, "seed": 0, "noise_seed": 1, "dtype": "float64"},
            "g_loss": float(r["g_loss"]), "d_loss": float(r["d_loss"]),
            "grad_norms": {k: float(v.norm()) for k, v in r["grads"].items()},
            "grad_sums": {k: float(v.sum()) for k, v in r["grads"].items()}}


if __name__ == "__main__":
    json.dump(rmse_fixture(), open(os.path.join(HERE, "rmse_known_answers.json"), "w"), indent=1)
    json.dump(oracle_fixture(), open(os.path.join(HERE, "oracle_iwgan_tiny.json"), "w"), indent=1)
    print("wrote fixtures")
