"""CPU restatement of the reference's batch assembly and vocoder hand-off -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product
(``b200taco/data.py``, ``b200taco/engine.py``) never does.  Pinned against outputs of the unmodified reference
(``tests/golden/data_b5.npz``, written by ``tests/golden/make_golden_data.py``, which imports ``src/data_utils.py`` and
executes the statements of train.py:71-79, :500-508 and inference.py:71 verbatim).
"""
import torch

N_MELS = 80


def collate(batch):
    """TextMelCollate.__call__, src/data_utils.py:45-84: sort by text length (descending), zero-pad text and mel rows in
    that order; ``mel_lengths`` stays in the ORIGINAL sample order (it is computed at :60 and never permuted)."""
    lens = torch.tensor([len(x[0]) for x in batch], dtype=torch.long)
    input_lengths, order = torch.sort(lens, dim=0, descending=True)                    # :54-56
    mel_lengths = torch.tensor([x[1].shape[1] for x in batch], dtype=torch.long)       # :60
    text_padded = torch.zeros(len(batch), int(input_lengths[0]), dtype=torch.long)     # :64, :68
    mel_padded = torch.zeros(len(batch), N_MELS, int(mel_lengths.max()))               # :65, :69
    for row, src in enumerate(order.tolist()):                                         # :72-82
        text, mel = batch[src]
        text_padded[row, :text.shape[0]] = text
        mel_padded[row, :, :mel.shape[1]] = mel
    return text_padded, input_lengths, mel_padded, mel_lengths


def prepare_training_batch(batch):
    """train.py:496-508: second sort by input length and the stop-token targets."""
    text_padded, input_lengths, mel_padded, mel_lengths = batch
    idx = torch.argsort(input_lengths, descending=True)                                # :500
    text_padded, input_lengths, mel_padded, mel_lengths = text_padded[idx], input_lengths[idx], mel_padded[idx], mel_lengths[idx]
    gate = torch.zeros(mel_padded.shape[0], mel_padded.shape[2])
    for j, l in enumerate(mel_lengths.tolist()):                                       # :506-508
        gate[j, l - 1:] = 1
    return text_padded, input_lengths, mel_padded, mel_lengths, gate


def stop_lengths(gate_logits, fallback_lengths=None):
    """Per-utterance trim of train.py:71-79: first frame with sigmoid(gate) > 0.5, plus one; without a firing gate the
    fallback length (or all frames)."""
    out = []
    for b in range(gate_logits.shape[0]):
        stops = (torch.sigmoid(gate_logits[b]) > 0.5).nonzero(as_tuple=True)[0]
        if len(stops) > 0:
            out.append(int(stops[0]) + 1)
        else:
            out.append(int(fallback_lengths[b]) if fallback_lengths is not None else gate_logits.shape[1])
    return out


def vocoder_layout(mel_postnet, lengths=None, frames=None):
    """inference.py:71: (B, S, 80) -> (B, 80, S); rows zeroed from their own length on, frame axis cut to ``frames``."""
    voc = mel_postnet.transpose(1, 2).clone()
    if lengths is not None:
        for b, n in enumerate(lengths):
            voc[b, :, int(n):] = 0
    return voc[:, :, :frames] if frames is not None else voc
