"""``Tacotron2Loss`` of the reference (train.py:126-217) with its Python B x S loop on the device.

Same constructor, attributes (``current_attention_weight``, ``global_step``, ``current_sigma`` ...), call signature and
return value as the reference class, so ``train.py`` can use it unchanged.  The masked L1 / BCE / KL terms are the
reference's own tensor expressions; what changes is ``create_diagonal_attention_target``: the reference fills the
(B, T_dec, T_enc) target with a Python double loop (51,200 iterations and ``.item()`` syncs at B=64, 800 frames - seconds per
step), here it is one kernel launch (``b200taco_guided_attention_target``).  CUDA tensors only; there is no CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi, config
from ._cabi import B200TacoError, check


def guided_attention_target(text_lengths: torch.Tensor, sigma: torch.Tensor, num_steps: int, max_text_len: int) -> torch.Tensor:
    """(B, num_steps, max_text_len) float32 target of train.py:148-171 for per-utterance widths ``sigma`` (B,)."""
    if text_lengths.device.type != "cuda":
        raise B200TacoError(f"b200taco.loss runs on CUDA tensors only (got {text_lengths.device}); there is no CPU fallback")
    dev = text_lengths.device
    tl = text_lengths.to(dtype=torch.int64).contiguous()
    sg = sigma.to(device=dev, dtype=torch.float32).contiguous()
    B = tl.shape[0]
    out = torch.empty(B, num_steps, max_text_len, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_cabi.lib().b200taco_guided_attention_target(
            tl.data_ptr(), sg.data_ptr(), B, int(num_steps), int(max_text_len), out.data_ptr(),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "b200taco_guided_attention_target")
    return out


class Tacotron2Loss(nn.Module):
    def __init__(self):
        super().__init__()
        # train.py:129-141
        self.l1_loss = nn.L1Loss(reduction="none")
        self.bce_loss = nn.BCEWithLogitsLoss()
        self.attn_weight_start = 1.0
        self.min_attn_weight = 0.2
        self.entropy_target = 3.5
        self.current_attention_weight = self.attn_weight_start
        self.global_step = 0
        self.sigma_warmup_steps = config.attention_sigma_warmup_steps

    def get_mask_from_lengths(self, lengths):
        """train.py:143-147 (True where padding)."""
        max_len = int(torch.max(lengths).item())
        ids = torch.arange(0, max_len, device=lengths.device, dtype=torch.long)
        return ~(ids < lengths.unsqueeze(1)).bool()

    def create_diagonal_attention_target(self, text_lengths, num_steps, alignments):
        """train.py:148-171, one kernel instead of the B x S loop."""
        max_text_len = int(text_lengths.max().item())
        init_sigma = torch.clamp(text_lengths.float() * config.attention_initial_sigma_factor, min=3.0,
                                 max=config.attention_max_sigma_cap)
        progress = min(1.0, self.global_step / float(self.sigma_warmup_steps))
        sigma = init_sigma - (init_sigma - config.attention_min_sigma) * progress
        self.current_sigma = float(sigma.mean().item())
        return guided_attention_target(text_lengths.to(alignments[0].device), sigma, num_steps, max_text_len)

    def forward(self, model_outputs, targets, text_lengths=None):
        mel_out_postnet, mel_out, gate_out, alignments = model_outputs
        mel_target, gate_target, mel_lengths = targets
        mel_target = mel_target.transpose(1, 2)
        mask = self.get_mask_from_lengths(mel_lengths).unsqueeze(-1).expand(-1, -1, config.n_mels)
        mel_loss_1 = self.l1_loss(mel_out, mel_target).masked_fill(mask, 0.0)
        mel_loss_2 = self.l1_loss(mel_out_postnet, mel_target).masked_fill(mask, 0.0)
        valid = (~mask).float().sum()
        loss_mel = mel_loss_1.sum() / valid + mel_loss_2.sum() / valid
        loss_gate = self.bce_loss(gate_out, gate_target)

        attention_kl = torch.tensor(0.0, device=mel_out.device)
        if len(alignments) > 1 and text_lengths is not None:
            attn_weights = torch.stack(list(alignments), dim=1)                  # (B, T_dec, T_enc_max)
            T_dec = attn_weights.shape[1]
            diagonal_target = self.create_diagonal_attention_target(text_lengths, T_dec, alignments)
            attn_safe = attn_weights.clamp_min(1e-8)
            log_pred = attn_safe.log()
            attention_kl = F.kl_div(log_pred, diagonal_target, reduction="batchmean") / T_dec
            attention_kl = torch.clamp(attention_kl, max=150.0)
            attn_entropy = -(attn_safe * log_pred).sum(dim=2).mean()
            if attn_entropy <= self.entropy_target:
                ratio = (attn_entropy / self.entropy_target).clamp_min(0.0)
                self.current_attention_weight = max(self.min_attn_weight, self.attn_weight_start * ratio.item())
            else:
                self.current_attention_weight = self.attn_weight_start

        total_loss = loss_mel + loss_gate + self.current_attention_weight * attention_kl
        self.global_step += 1
        return total_loss, loss_mel, loss_gate, attention_kl
