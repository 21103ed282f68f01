"""Batch assembly on the device: ``TextMelCollate`` of the reference (src/data_utils.py:45-84) and the per-batch preparation of
its training loop (train.py:500-508: second length sort, stop-token targets) without Python loops over the batch.

``TextMelCollate()(batch)`` takes the same input as the reference's collate function - a list of ``(text_sequence (L_i,)
int64, mel (80, S_i) float32)`` pairs, e.g. straight from ``TextMelDataset.__getitem__`` - and returns the same 4-tuple
``(text_padded, input_lengths, mel_padded, mel_lengths)``, as DEVICE tensors.  The ragged samples are concatenated into two
flat pinned host buffers (one host pass over the data, no padded intermediate), copied with two asynchronous transfers, and
padded / reordered by ONE kernel (``b200taco_collate``, csrc/data.cu).  CUDA only; there is no CPU fallback.

Reference quirk, mirrored by default: ``input_lengths`` comes back sorted (descending) together with the rows of
``text_padded`` / ``mel_padded``, but ``mel_lengths`` is returned in the ORIGINAL sample order (src/data_utils.py:60 builds
it before the sorted fill loop of :72-82 and never permutes it), so ``mel_lengths[i]`` does not belong to row i unless the
batch happened to be sorted already.  ``mirror_reference_mel_lengths=False`` returns them in row order instead.
"""
import ctypes as C

import torch

from . import _cabi, config
from ._cabi import B200TacoError, check
from .loss import gate_targets


class TextMelCollate:
    def __init__(self, device="cuda", mirror_reference_mel_lengths: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B200TacoError(f"b200taco.data assembles batches on a CUDA device (got {self.device}); there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mirror_reference_mel_lengths = mirror_reference_mel_lengths
        self._pinned = {}

    def _staging(self, name, n, dtype):
        """Pinned host staging buffer, grown geometrically and reused across batches (the copy of batch i must have completed
        before the buffer is rewritten for batch i + 1: __call__ synchronises on its copy event)."""
        buf = self._pinned.get(name)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(n, 1) * 2, dtype=dtype).pin_memory()
            self._pinned[name] = buf
        return buf[:n]

    def __call__(self, batch):
        B = len(batch)
        text_lens = torch.LongTensor([len(x[0]) for x in batch])
        input_lengths, order = torch.sort(text_lens, dim=0, descending=True)           # src/data_utils.py:54-56
        mel_lens = torch.LongTensor([x[1].shape[1] for x in batch])                    # :60 (original order)
        Lmax, Smax = int(input_lengths[0]), int(mel_lens.max())                        # :57, :61
        for _, mel in batch:
            if mel.shape[0] != config.n_mels:
                raise B200TacoError(f"collate: mel with {mel.shape[0]} bins, expected {config.n_mels}")
        t_off = torch.zeros(B + 1, dtype=torch.int64)
        m_off = torch.zeros(B + 1, dtype=torch.int64)
        torch.cumsum(text_lens, 0, out=t_off[1:])
        torch.cumsum(mel_lens, 0, out=m_off[1:])
        text_flat = self._staging("text", int(t_off[-1]), torch.int64)
        mel_flat = self._staging("mel", int(m_off[-1]) * config.n_mels, torch.float32)
        for i, (text, mel) in enumerate(batch):                                        # one pass over the ragged data
            text_flat[int(t_off[i]):int(t_off[i + 1])] = torch.as_tensor(text, dtype=torch.int64)
            mel_flat[int(m_off[i]) * config.n_mels:int(m_off[i + 1]) * config.n_mels] = \
                torch.as_tensor(mel, dtype=torch.float32).reshape(-1)
        meta = self._staging("meta", 3 * (B + 1), torch.int64)
        meta[:B + 1] = t_off
        meta[B + 1:2 * B + 2] = m_off
        meta[2 * B + 2:3 * B + 2] = order
        dev = self.device
        with torch.cuda.device(dev):
            text_d = text_flat.to(dev, non_blocking=True)
            mel_d = mel_flat.to(dev, non_blocking=True)
            meta_d = meta.to(dev, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record()
            order_d = meta_d[2 * B + 2:3 * B + 2].to(torch.int32)
            text_padded = torch.empty(B, Lmax, dtype=torch.int64, device=dev)
            mel_padded = torch.empty(B, config.n_mels, Smax, dtype=torch.float32, device=dev)
            check(_cabi.lib().b200taco_collate(text_d.data_ptr(), meta_d.data_ptr(), mel_d.data_ptr(), meta_d[B + 1:].data_ptr(),
                                               order_d.data_ptr(), B, Lmax, Smax, text_padded.data_ptr(), mel_padded.data_ptr(),
                                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "b200taco_collate")
            mel_lengths = mel_lens if self.mirror_reference_mel_lengths else mel_lens[order]
            out = (text_padded, input_lengths.to(dev, non_blocking=True), mel_padded, mel_lengths.to(dev, non_blocking=True))
            copied.synchronize()                     # the staging buffers may be rewritten by the next call
        return out


def prepare_training_batch(batch):
    """train.py:500-508 on device tensors: the loop's second sort by input length (descending; a no-op permutation for a batch
    that came out of ``TextMelCollate``, kept because the reference applies it - also to ``mel_lengths``) and the stop-token
    targets (ones from frame ``mel_lengths[j] - 1`` on) as one kernel launch instead of a Python loop with a ``.item()`` per row.
    Returns ``(text_padded, input_lengths, mel_padded, mel_lengths, gate_target)``."""
    text_padded, input_lengths, mel_padded, mel_lengths = batch
    # train.py:500 sorts on the host; the batch is already in descending order, where the host sort is the identity - a stable
    # sort reproduces that on the device (the device's default sort permutes equal lengths)
    sort_idx = torch.argsort(input_lengths, descending=True, stable=True)
    text_padded, input_lengths = text_padded[sort_idx], input_lengths[sort_idx]
    mel_padded, mel_lengths = mel_padded[sort_idx], mel_lengths[sort_idx]
    return text_padded, input_lengths, mel_padded, mel_lengths, gate_targets(mel_lengths, mel_padded.shape[2])
