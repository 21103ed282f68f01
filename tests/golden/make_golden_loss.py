"""Golden vectors of the reference's Tacotron2Loss (train.py:126-217), produced by RUNNING THE REFERENCE in the build container.

    python tests/golden/make_golden_loss.py        # needs /root/reference; writes tests/golden/loss_b3.npz

train.py is imported unmodified; ``librosa``, ``matplotlib`` and ``soundfile`` (imported at module level, never called on this
path) are stubbed.  Inputs come from ``loss_case`` below, which the tests import to rebuild them from the seed.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def loss_case(B=3, S=12, lengths=(9, 7, 4), mel_lengths=(12, 10, 7), seed=21):
    """Seeded inputs of one loss evaluation, shaped like a model output / batch of train.py:500-516."""
    g = torch.Generator().manual_seed(seed)
    Tmax = max(lengths)
    text_lengths = torch.tensor(lengths, dtype=torch.long)
    mel_l = torch.tensor(mel_lengths, dtype=torch.long)
    mel_target = torch.randn(B, 80, S, generator=g) * 2 - 5
    mel_out = mel_target.transpose(1, 2) + 0.3 * torch.randn(B, S, 80, generator=g)
    mel_post = mel_target.transpose(1, 2) + 0.2 * torch.randn(B, S, 80, generator=g)
    gate_out = torch.randn(B, S, generator=g)
    gate_target = torch.zeros(B, S)
    for b in range(B):
        gate_target[b, mel_lengths[b] - 1:] = 1.0                      # train.py:506-508
    logits = 3.0 * torch.randn(S, B, Tmax, generator=g)
    for b in range(B):
        logits[:, b, lengths[b]:] = -1e9                               # padded positions carry no attention
    alignments = list(torch.softmax(logits, dim=2).unbind(0))
    return (mel_post, mel_out, gate_out, alignments), (mel_target, gate_target, mel_l), text_lengths


def main():
    for m in ("librosa", "matplotlib", "matplotlib.pyplot", "soundfile"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, REF)
    os.chdir(REF)
    import train  # the unmodified reference
    out = {}
    for step in (0, 2500, 9000):
        crit = train.Tacotron2Loss()
        crit.global_step = step
        outs, targets, tl = loss_case()
        total, mel, gate, kl = crit(outs, targets, tl)
        out[f"s{step}_losses"] = np.array([float(total), float(mel), float(gate), float(kl), crit.current_attention_weight],
                                          dtype=np.float64)
        crit2 = train.Tacotron2Loss()
        crit2.global_step = step
        out[f"s{step}_target"] = crit2.create_diagonal_attention_target(tl, 12, outs[3]).numpy()
    np.savez_compressed(os.path.join(HERE, "loss_b3.npz"), **out)
    print({k: v if v.ndim == 1 else v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
