"""``Tacotron2Loss`` of the reference (train.py:126-217) with its Python B x S loop on the device.

Same constructor, attributes (``current_attention_weight``, ``global_step``, ``current_sigma`` ...), call signature and
4-tuple result as the reference class, so ``train.py`` can use it unchanged (``criterion = Tacotron2Loss()``, train.py:333).
The loss terms are small tensor expressions (``loss_terms`` below, device agnostic); the expensive part of the reference is
``create_diagonal_attention_target``, a Python double loop over B x T_dec with ``.item()`` syncs (51,200 iterations at B=64
and 800 frames: seconds per step).  Here the target is ONE kernel launch (``b200taco_guided_attention_target``, csrc/loss.cu).
CUDA tensors only; there is no CPU fallback.
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


def annealed_sigma(text_lengths: torch.Tensor, global_step: int, warmup_steps: int) -> torch.Tensor:
    """Width of the guided-attention Gaussian per utterance (train.py:151-156): from clamp(0.05 L, 3, 20) down to 1.0."""
    start = (text_lengths.float() * config.attention_initial_sigma_factor).clamp(3.0, config.attention_max_sigma_cap)
    done = min(1.0, global_step / float(warmup_steps))
    return start - (start - config.attention_min_sigma) * done


def masked_l1(pred: torch.Tensor, target: torch.Tensor, frame_is_pad: torch.Tensor) -> torch.Tensor:
    """Mean absolute error over the mel bins of the real frames (train.py:182-193); pred/target (B,S,80), mask (B,S)."""
    keep = (~frame_is_pad).unsqueeze(-1).to(pred.dtype)
    n_valid = keep.sum() * pred.shape[-1]
    return ((pred - target).abs() * keep).sum() / n_valid


def guided_attention_terms(alignments, diagonal_target: torch.Tensor):
    """KL(target || attention) per decoder step and the mean attention entropy (train.py:199-204)."""
    attn = torch.stack(list(alignments), dim=1).clamp_min(1e-8)                  # (B, T_dec, T_enc)
    log_attn = attn.log()
    kl = F.kl_div(log_attn, diagonal_target, reduction="batchmean") / attn.shape[1]
    entropy = -(attn * log_attn).sum(dim=2).mean()
    return kl.clamp(max=150.0), entropy


def loss_terms(model_outputs, targets, diagonal_target=None):
    """(mel, gate, kl, entropy) of train.py:174-204; ``diagonal_target`` None = no guided-attention term."""
    mel_post, mel_coarse, gate_logits, alignments = model_outputs
    mel_target, gate_target, mel_lengths = targets
    want = mel_target.transpose(1, 2)                                            # (B,80,S) -> (B,S,80)
    frames = torch.arange(int(mel_lengths.max()), device=mel_lengths.device)
    frame_is_pad = frames[None, :] >= mel_lengths[:, None]
    mel = masked_l1(mel_coarse, want, frame_is_pad) + masked_l1(mel_post, want, frame_is_pad)
    gate = F.binary_cross_entropy_with_logits(gate_logits, gate_target)
    if diagonal_target is None:
        return mel, gate, torch.zeros((), device=mel_coarse.device), None
    kl, entropy = guided_attention_terms(alignments, diagonal_target)
    return mel, gate, kl, entropy


class Tacotron2Loss(nn.Module):
    def __init__(self):
        super().__init__()
        # attribute surface of train.py:129-141
        self.l1_loss = nn.L1Loss(reduction="none")
        self.bce_loss = nn.BCEWithLogitsLoss()
        self.attn_weight_start = 1.0
        self.min_attn_weight = 0.2
        self.entropy_target = 3.5
        self.current_attention_weight = self.attn_weight_start
        self.global_step = 0
        self.sigma_warmup_steps = config.attention_sigma_warmup_steps

    def get_mask_from_lengths(self, lengths):
        """True where padding (train.py:143-147)."""
        frames = torch.arange(int(lengths.max().item()), device=lengths.device)
        return frames[None, :] >= lengths[:, None]

    def create_diagonal_attention_target(self, text_lengths, num_steps, alignments):
        """train.py:148-171 as one kernel launch."""
        sigma = annealed_sigma(text_lengths, self.global_step, self.sigma_warmup_steps)
        self.current_sigma = float(sigma.mean().item())
        return guided_attention_target(text_lengths.to(alignments[0].device), sigma, num_steps, int(text_lengths.max().item()))

    def attention_weight(self, entropy) -> float:
        """Entropy-adaptive weight of the guided-attention term (train.py:205-210)."""
        if entropy is None or float(entropy) > self.entropy_target:
            return self.attn_weight_start
        ratio = max(0.0, float(entropy) / self.entropy_target)
        return max(self.min_attn_weight, self.attn_weight_start * ratio)

    def forward(self, model_outputs, targets, text_lengths=None):
        alignments = model_outputs[3]
        target = None
        if len(alignments) > 1 and text_lengths is not None:
            target = self.create_diagonal_attention_target(text_lengths, len(alignments), alignments)
        loss_mel, loss_gate, attention_kl, entropy = loss_terms(model_outputs, targets, target)
        if target is not None:
            self.current_attention_weight = self.attention_weight(entropy)
        total = loss_mel + loss_gate + self.current_attention_weight * attention_kl
        self.global_step += 1
        return total, loss_mel, loss_gate, attention_kl
