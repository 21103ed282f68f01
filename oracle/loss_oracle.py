"""CPU restatement of the reference's ``Tacotron2Loss`` (train.py:126-217) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product
(``b200taco/loss.py``) never does.  Pinned against outputs of the unmodified reference run in the build container
(``tests/golden/loss_b3.npz``, written by ``tests/golden/make_golden_loss.py``).

The restatement is vectorised (the reference loops over B x S in Python, train.py:163-171) but performs the same float32
operations per element.
"""
import torch
import torch.nn.functional as F

# src/config.py:51-54
ATTENTION_INITIAL_SIGMA_FACTOR = 0.05
ATTENTION_SIGMA_WARMUP_STEPS = 4000
ATTENTION_MIN_SIGMA = 1.0
ATTENTION_MAX_SIGMA_CAP = 20.0
N_MELS = 80


def sigma_schedule(text_lengths, global_step):
    """train.py:151-156."""
    init_sigma = torch.clamp(text_lengths.float() * ATTENTION_INITIAL_SIGMA_FACTOR, min=3.0, max=ATTENTION_MAX_SIGMA_CAP)
    progress = min(1.0, global_step / float(ATTENTION_SIGMA_WARMUP_STEPS))
    return init_sigma - (init_sigma - ATTENTION_MIN_SIGMA) * progress


def diagonal_attention_target(text_lengths, num_steps, global_step):
    """train.py:148-171: (B, num_steps, max_text_len)."""
    B = text_lengths.shape[0]
    Tmax = int(text_lengths.max())
    sigma = sigma_schedule(text_lengths, global_step)
    pos = torch.arange(Tmax, dtype=torch.float32)
    t = torch.arange(num_steps, dtype=torch.int64)
    L = text_lengths.to(torch.int64)
    expected = torch.minimum((t[None, :] * L[:, None]) // num_steps, L[:, None] - 1).float()          # (B, S)
    g = torch.exp(-0.5 * ((pos[None, None, :] - expected[:, :, None]) / sigma[:, None, None]) ** 2)
    valid = pos[None, None, :] < L[:, None, None].float()
    g = torch.where(valid, g, torch.zeros_like(g))
    g = g / (g.sum(dim=2, keepdim=True) + 1e-8)
    assert g.shape == (B, num_steps, Tmax)
    return g


def tacotron2_loss(model_outputs, targets, text_lengths, global_step, attn_weight_start=1.0, min_attn_weight=0.2,
                   entropy_target=3.5):
    """train.py:174-217.  Returns (total, mel, gate, kl, attention_weight_used)."""
    mel_post, mel_out, gate_out, alignments = model_outputs
    mel_target, gate_target, mel_lengths = targets
    mel_target = mel_target.transpose(1, 2)
    ids = torch.arange(int(mel_lengths.max()))
    mask = ~(ids[None, :] < mel_lengths[:, None])
    mask = mask.unsqueeze(-1).expand(-1, -1, N_MELS)
    l1 = (mel_out - mel_target).abs().masked_fill(mask, 0.0)
    l2 = (mel_post - mel_target).abs().masked_fill(mask, 0.0)
    valid = (~mask).float().sum()
    loss_mel = l1.sum() / valid + l2.sum() / valid
    loss_gate = F.binary_cross_entropy_with_logits(gate_out, gate_target)
    kl = torch.tensor(0.0)
    weight = attn_weight_start
    if len(alignments) > 1 and text_lengths is not None:
        attn = torch.stack(list(alignments), dim=1)
        T_dec = attn.shape[1]
        target = diagonal_attention_target(text_lengths, T_dec, global_step)
        safe = attn.clamp_min(1e-8)
        logp = safe.log()
        kl = F.kl_div(logp, target, reduction="batchmean") / T_dec
        kl = torch.clamp(kl, max=150.0)
        entropy = -(safe * logp).sum(dim=2).mean()
        if entropy <= entropy_target:
            weight = max(min_attn_weight, attn_weight_start * float((entropy / entropy_target).clamp_min(0.0)))
    total = loss_mel + loss_gate + weight * kl
    return total, loss_mel, loss_gate, kl, weight
