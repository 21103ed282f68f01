"""Shared helpers of the test-suite: golden fixtures, weight sets, comparisons."""
import os

import numpy as np
import torch

from b200taco import synth
from b200taco.model import Tacotron2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def weights(kind="default", bf16=False, ctx_center=None, dtype=torch.float32):
    """Rebuild the weight sets of tests/golden/make_golden.py from seeds (CPU state_dict)."""
    sd = synth.seeded_state_dict(Tacotron2)
    with torch.no_grad():
        if kind in ("stress", "planted"):
            synth.stress_init_(sd)
        if kind == "planted":
            synth.plant_stop_counter_(sd, rate=6e-3, jitter=2e-2, ctx_center=ctx_center)
        if bf16:
            synth.round_to_bf16_(sd)
    if dtype != torch.float32:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return sd


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def stack(aligns):
    return torch.stack(list(aligns), dim=0)
