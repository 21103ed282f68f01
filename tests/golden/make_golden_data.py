"""Golden vectors of the reference's batch assembly (src/data_utils.py:45-84, train.py:496-508) and vocoder hand-off
(inference.py:71, train.py:71-79), produced by RUNNING THE REFERENCE CODE in the build container.

    python tests/golden/make_golden_data.py        # needs /root/reference; writes tests/golden/data_b5.npz

``src/data_utils.py`` is imported unmodified (``librosa`` stubbed: src/__init__ pulls in src/audio.py).  The statements of
``train.py`` and ``inference.py`` that this path consists of live inside long functions that cannot be called on their own;
they are executed here verbatim, with their line numbers.  Inputs come from ``data_case`` below, which the tests import.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def data_case(seed=33, text_lens=(17, 9, 23, 9, 14), mel_lens=(41, 25, 60, 31, 37)):
    """A ragged batch as ``TextMelDataset.__getitem__`` yields it: (text (L,) int64, mel (80, S) float32) pairs.  Two samples
    share a text length (the order of ties is the sort's), and text / mel lengths are not monotone in each other (so the
    reference's unsorted ``mel_lengths`` differs from the row order)."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.randint(0, 72, (L,), generator=g), torch.randn(80, S, generator=g) * 2 - 5) for L, S in zip(text_lens, mel_lens)]


def handoff_case(seed=34, B=4, S=29):
    g = torch.Generator().manual_seed(seed)
    mel_post = torch.randn(B, S, 80, generator=g)
    gate = torch.randn(B, S, generator=g) - 1.5
    gate[2] = -5.0                                                     # a row whose gate never fires
    return mel_post, gate, torch.tensor([20, 25, 17, 29])


def main():
    for m in ("librosa", "librosa.filters", "librosa.feature", "librosa.effects"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, REF)
    from src.data_utils import TextMelCollate  # the unmodified reference
    out = {}
    batch = data_case()
    text_padded, input_lengths, mel_padded, mel_lengths = TextMelCollate()(batch)
    out.update(text_padded=text_padded.numpy(), input_lengths=input_lengths.numpy(), mel_padded=mel_padded.numpy(),
               mel_lengths=mel_lengths.numpy())
    # train.py:499-508, verbatim (device = cpu)
    device = torch.device("cpu")
    sort_idx = torch.argsort(input_lengths, descending=True)
    text_padded = text_padded[sort_idx].to(device)
    input_lengths = input_lengths[sort_idx].to(device)
    mel_padded = mel_padded[sort_idx].to(device)
    mel_lengths = mel_lengths[sort_idx].to(device)
    gate_target = torch.zeros(mel_padded.size(0), mel_padded.size(2), device=device)
    for j, l in enumerate(mel_lengths):
        gate_target[j, l.item()-1:] = 1
    out.update(t_text_padded=text_padded.numpy(), t_input_lengths=input_lengths.numpy(), t_mel_padded=mel_padded.numpy(),
               t_mel_lengths=mel_lengths.numpy(), t_gate_target=gate_target.numpy())
    # inference.py:71 and train.py:69-79, verbatim
    mel_post, gate, mel_lengths = handoff_case()
    mel_for_vocoder = mel_post.transpose(1, 2)
    ends = []
    max_len_cap = 1
    for b in range(mel_post.size(0)):
        gate_sig = torch.sigmoid(gate[b])
        stops = (gate_sig > 0.5).nonzero(as_tuple=True)[0]
        if len(stops) == 0 and max_len_cap is not None:
            end_idx = int(mel_lengths[b].item())
        else:
            end_idx = (stops[0].item() + 1) if len(stops) > 0 else mel_post.size(1)
        ends.append(end_idx)
    trimmed = torch.zeros_like(mel_for_vocoder)
    for b, e in enumerate(ends):
        trimmed[b, :, :e] = mel_post[b, :e].transpose(0, 1)            # mel_b = mel_post[b, :end_idx]  (train.py:78)
    out.update(h_mel_for_vocoder=mel_for_vocoder.contiguous().numpy(), h_ends=np.array(ends), h_trimmed=trimmed.numpy())
    np.savez_compressed(os.path.join(HERE, "data_b5.npz"), **out)
    print({k: v.shape for k, v in out.items()}, "ends", ends, "mel_lengths", out["mel_lengths"], "input_lengths", out["input_lengths"])


if __name__ == "__main__":
    main()
