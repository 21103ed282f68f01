"""oracle/train_oracle.py (train-mode restatement: dropout with explicit masks, BatchNorm batch statistics, gradients by torch
autograd) held to golden vectors of the UNMODIFIED reference run in train mode (tests/golden/train_b3.npz, written by
tests/golden/make_golden_train.py): outputs, all 62 parameter gradients (norm, a seeded random projection, 256 strided
entries each) and the BatchNorm running statistics after the step.

Tolerances: outputs 2e-6 (mel_coarse, gate) / 2e-7 (alignment) absolute as for the eval oracle, 2e-5 for mel_post; gradients 2e-5 of the tensor's largest
entry (float32 sums in a different order; the reference's cuDNN-free CPU kernels and the oracle's explicit ops differ in
accumulation order only).
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

import util
from b200taco import synth
from b200taco.model import Tacotron2
from oracle import train_oracle as TO

_spec = importlib.util.spec_from_file_location("make_golden_train", os.path.join(util.GOLDEN, "make_golden_train.py"))
_mgt = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mgt)


def oracle_step(sd, use_masks):
    ids, mel, lengths, dec, post, w = _mgt.train_case()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = {**sd, **params}
    mem, enc_stats = TO.encoder_train(full, ids)
    coarse, gate, align = TO.decoder_train(full, mem, mel, lengths, dec if use_masks else None)
    mel_post, post_stats = TO.postnet_train(full, coarse, post if use_masks else None)
    loss = (mel_post * w["post"]).sum() + (coarse * w["coarse"]).sum() + (gate * w["gate"]).sum() + (align * w["align"]).sum()
    loss.backward()
    return (mel_post, coarse, gate, align), {k: p.grad for k, p in params.items()}, enc_stats, post_stats


@pytest.mark.parametrize("case,use_masks", [("drop", True), ("nodrop", False)])
def test_train_oracle_matches_reference_golden(case, use_masks):
    g = util.golden("train_b3")
    sd = synth.seeded_state_dict(Tacotron2)
    outs, grads, enc_stats, post_stats = oracle_step(sd, use_masks)
    for got, key, tol in zip(outs, ("mel_post", "mel_coarse", "gate", "align"), (2e-5, 2e-6, 2e-6, 2e-7)):
        # mel_post: five BatchNorm layers normalise with statistics over only B * S = 21 samples, which amplifies rounding
        assert util.maxdiff(got.detach(), g[f"{case}_{key}"]) <= tol, (key, util.maxdiff(got.detach(), g[f"{case}_{key}"]))
    got = _mgt.summarise(grads)
    keys = [k[len(case) + 6:] for k in g if k.startswith(f"{case}_grad/")]
    assert len(keys) == 62 and sorted(keys) == sorted(got)
    for k in keys:
        ref = g[f"{case}_grad/{k}"].numpy()
        mine = got[k]
        scale = max(ref[2], 1e-30)                                   # largest entry of the reference gradient
        if k.endswith(".0.bias") or k.endswith("attention.v.bias"):
            # a conv bias in front of batch-statistics BatchNorm, and the scalar added to every energy before the softmax,
            # have mathematically zero gradient: both sides hold rounding noise
            assert ref[0] <= 1e-3 and mine[0] <= 1e-3, k
            continue
        assert abs(mine[0] - ref[0]) <= 2e-5 * ref[0] + 1e-12, (k, mine[0], ref[0])                  # norm
        assert np.abs(mine[3:] - ref[3:]).max() <= 2e-5 * scale, (k, np.abs(mine[3:] - ref[3:]).max(), scale)   # entries
        assert abs(mine[1] - ref[1]) <= 2e-5 * ref[0] * 8 + 1e-12, k                                # projection ~ N(0, norm)
    # BatchNorm side effect (momentum 0.1, unbiased variance) from the oracle's batch statistics
    B, T, S = _mgt.B, _mgt.T, _mgt.S
    for prefix, stats, n in (("encoder", enc_stats, B * T), ("postnet", post_stats, B * S)):
        for i, (mean, var) in enumerate(stats):
            rm, rv = g[f"{case}_buf/{prefix}.convolutions.{i}.1.running_mean"], g[f"{case}_buf/{prefix}.convolutions.{i}.1.running_var"]
            assert util.maxdiff(0.1 * mean.detach(), rm) <= 2e-6
            assert util.maxdiff(0.9 + 0.1 * var.detach() * n / (n - 1), rv) <= 2e-6 * max(1.0, float(rv.abs().max()))
            assert int(g[f"{case}_buf/{prefix}.convolutions.{i}.1.num_batches_tracked"]) == 1
