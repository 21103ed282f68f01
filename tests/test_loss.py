"""The reference's Tacotron2Loss (train.py:126-217): CPU oracle against golden vectors of the unmodified reference, and the
device version (guided-attention target kernel through the C ABI) against both.

Tolerances: losses 2e-6 relative (float32 sums in a different order), target 1e-6 absolute on values in [0, 1].
"""
import importlib.util
import os

import pytest
import torch

import util
from oracle import loss_oracle as LO

_spec = importlib.util.spec_from_file_location("make_golden_loss", os.path.join(util.GOLDEN, "make_golden_loss.py"))
_mgl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mgl)
STEPS = (0, 2500, 9000)


def rel(a, b):
    return abs(float(a) - float(b)) / max(1e-12, abs(float(b)))


@pytest.mark.parametrize("step", STEPS)
def test_loss_oracle_matches_reference_golden(step):
    g = util.golden("loss_b3")
    outs, targets, tl = _mgl.loss_case()
    total, mel, gate, kl, weight = LO.tacotron2_loss(outs, targets, tl, step)
    want = g[f"s{step}_losses"]
    for got, ref in zip((total, mel, gate, kl, weight), want):
        assert rel(got, ref) <= 2e-6, (float(got), float(ref))
    assert util.maxdiff(LO.diagonal_attention_target(tl, 12, step), g[f"s{step}_target"]) <= 1e-6


@pytest.mark.parametrize("step", STEPS)
def test_product_loss_terms_match_reference_golden(step):
    """The device-agnostic part of b200taco.loss (everything but the target kernel) on the CPU, fed the golden target."""
    from b200taco.loss import Tacotron2Loss, annealed_sigma, loss_terms
    g = util.golden("loss_b3")
    outs, targets, tl = _mgl.loss_case()
    mel, gate, kl, entropy = loss_terms(outs, targets, g[f"s{step}_target"])
    crit = Tacotron2Loss()
    weight = crit.attention_weight(entropy)
    total = mel + gate + weight * kl
    for got, ref in zip((total, mel, gate, kl, weight), g[f"s{step}_losses"]):
        assert rel(got, ref) <= 2e-6, (float(got), float(ref))
    assert util.maxdiff(annealed_sigma(tl, step, crit.sigma_warmup_steps), LO.sigma_schedule(tl, step)) == 0.0
    # without text lengths there is no guided-attention term (train.py:196)
    mel0, gate0, kl0, ent0 = loss_terms(outs, targets, None)
    assert float(kl0) == 0.0 and ent0 is None and rel(mel0, mel) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("step", STEPS)
def test_device_loss_matches_reference_golden(step):
    from b200taco.loss import Tacotron2Loss
    g = util.golden("loss_b3")
    outs, targets, tl = _mgl.loss_case()
    dev = torch.device("cuda:0")
    crit = Tacotron2Loss()
    crit.global_step = step
    outs_d = (outs[0].to(dev), outs[1].to(dev), outs[2].to(dev), [a.to(dev) for a in outs[3]])
    targets_d = tuple(t.to(dev) for t in targets)
    total, mel, gate, kl = crit(outs_d, targets_d, tl.to(dev))
    want = g[f"s{step}_losses"]
    for got, ref in zip((total, mel, gate, kl, crit.current_attention_weight), want):
        assert rel(got, ref) <= 2e-6, (float(got), float(ref))
    assert crit.global_step == step + 1
    crit.global_step = step
    tgt = crit.create_diagonal_attention_target(tl.to(dev), 12, outs_d[3])
    assert util.maxdiff(tgt.cpu(), g[f"s{step}_target"]) <= 1e-6


@pytest.mark.gpu
def test_device_target_matches_oracle_at_training_size():
    """B=64, 800 decoder steps, ragged text lengths up to 150 (BASELINE config 5 shape): kernel vs vectorised CPU oracle."""
    from b200taco.loss import Tacotron2Loss
    gen = torch.Generator().manual_seed(5)
    tl = torch.randint(1, 151, (64,), generator=gen)
    tl[0], tl[1] = 150, 1
    for step in (0, 1234, 4000):
        crit = Tacotron2Loss()
        crit.global_step = step
        tgt = crit.create_diagonal_attention_target(tl.cuda(), 800, [torch.zeros(64, 150, device="cuda")])
        ref = LO.diagonal_attention_target(tl, 800, step)
        assert tgt.shape == ref.shape == (64, 800, 150)
        assert util.maxdiff(tgt.cpu(), ref) <= 1e-6
        assert abs(crit.current_sigma - float(LO.sigma_schedule(tl, step).mean())) <= 1e-6


def test_device_loss_has_no_cpu_fallback():
    from b200taco._cabi import B200TacoError
    from b200taco.loss import guided_attention_target
    with pytest.raises(B200TacoError, match="CUDA"):
        guided_attention_target(torch.tensor([3, 2]), torch.tensor([3.0, 3.0]), 4, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("step", STEPS)
def test_fused_device_loss_values_and_gradients(step):
    """b200taco_loss_fwd / _bwd (whole Tacotron2Loss.forward as three launches, scalars left on the device) against the
    reference golden (values) and against torch autograd through the CPU oracle (gradients of all four model outputs).
    The list of alignments is the unbind(0) of ONE (S,B,T) buffer, as b200taco.model returns it."""
    from b200taco.loss import Tacotron2Loss, alignment_block
    g = util.golden("loss_b3")
    outs, targets, tl = _mgl.loss_case()
    dev = torch.device("cuda:0")
    block = torch.stack(outs[3], 0).to(dev).requires_grad_(True)
    leaves = [outs[0].to(dev).requires_grad_(True), outs[1].to(dev).requires_grad_(True), outs[2].to(dev).requires_grad_(True)]
    aligns = list(block.unbind(0))
    assert alignment_block(aligns) is block                       # no torch.stack of S views
    crit = Tacotron2Loss()
    assert crit.fused
    crit.global_step = step
    total, mel, gate, kl = crit((leaves[0], leaves[1], leaves[2], aligns), tuple(t.to(dev) for t in targets), tl.to(dev))
    assert all(t.is_cuda and t.dim() == 0 for t in (total, mel, gate, kl))
    for got, ref in zip((total, mel, gate, kl, crit.current_attention_weight), g[f"s{step}_losses"]):
        assert rel(got, ref) <= 2e-6, (float(got), float(ref))
    assert abs(crit.current_sigma - float(LO.sigma_schedule(tl, step).mean())) <= 1e-6
    total.backward()
    # oracle gradients: torch autograd through the CPU restatement
    c_leaves = [outs[0].clone().requires_grad_(True), outs[1].clone().requires_grad_(True), outs[2].clone().requires_grad_(True)]
    c_block = torch.stack(outs[3], 0).requires_grad_(True)
    ref_total = LO.tacotron2_loss((c_leaves[0], c_leaves[1], c_leaves[2], list(c_block.unbind(0))), targets, tl, step)[0]
    ref_total.backward()
    for got, ref, name in zip(leaves + [block], c_leaves + [c_block], ("mel_post", "mel_coarse", "gate", "alignments")):
        scale = float(ref.grad.abs().max())
        assert util.maxdiff(got.grad.cpu(), ref.grad) <= 2e-6 * max(scale, 1e-12) + 1e-12, name


@pytest.mark.gpu
def test_fused_device_loss_without_attention_term_and_unfused_agree():
    """No text lengths -> no guided-attention term (train.py:196); fused and tensor-expression paths give the same numbers."""
    from b200taco.loss import Tacotron2Loss
    outs, targets, tl = _mgl.loss_case()
    dev = torch.device("cuda:0")
    outs_d = (outs[0].to(dev), outs[1].to(dev), outs[2].to(dev), [a.to(dev) for a in outs[3]])
    targets_d = tuple(t.to(dev) for t in targets)
    fused, plain = Tacotron2Loss(), Tacotron2Loss()
    plain.fused = False
    for lengths in (None, tl.to(dev)):
        a, b = fused(outs_d, targets_d, lengths), plain(outs_d, targets_d, lengths)
        for x, y in zip(a, b):
            assert rel(x, y) <= 2e-6 or abs(float(x) - float(y)) <= 1e-9
        assert abs(fused.current_attention_weight - plain.current_attention_weight) <= 1e-6
    assert float(fused(outs_d, targets_d, None)[3]) == 0.0


@pytest.mark.gpu
def test_fused_device_loss_at_training_size():
    """B=64, 800 frames, T=150 (BASELINE config 5 shape): fused kernels vs the tensor-expression path on the same device."""
    from b200taco.loss import Tacotron2Loss
    gen = torch.Generator().manual_seed(11)
    B, S, T = 64, 800, 150
    dev = torch.device("cuda:0")
    tl = torch.randint(20, T + 1, (B,), generator=gen); tl[0] = T
    ml = torch.randint(100, S + 1, (B,), generator=gen); ml[0] = S
    mel_t = (torch.randn(B, 80, S, generator=gen) * 2 - 5).to(dev)
    post = (mel_t.transpose(1, 2) + 0.2 * torch.randn(B, S, 80, generator=gen).to(dev)).contiguous()
    coarse = (mel_t.transpose(1, 2) + 0.3 * torch.randn(B, S, 80, generator=gen).to(dev)).contiguous()
    gate = torch.randn(B, S, generator=gen).to(dev)
    logits = 2.0 * torch.randn(S, B, T, generator=gen)
    for b in range(B):
        logits[:, b, int(tl[b]):] = -1e9
    block = torch.softmax(logits, 2).to(dev)
    from b200taco.loss import gate_targets
    gt = gate_targets(ml.to(dev), S)
    fused, plain = Tacotron2Loss(), Tacotron2Loss()
    plain.fused = False
    for crit in (fused, plain):
        crit.global_step = 1000
    a = fused((post, coarse, gate, list(block.unbind(0))), (mel_t, gt, ml.to(dev)), tl.to(dev))
    b = plain((post, coarse, gate, list(block.unbind(0))), (mel_t, gt, ml.to(dev)), tl.to(dev))
    for x, y in zip(a, b):
        assert rel(x, y) <= 2e-5, (float(x), float(y))            # fp32 sums over 4 M elements in different orders
    assert abs(fused.current_attention_weight - plain.current_attention_weight) <= 1e-5
