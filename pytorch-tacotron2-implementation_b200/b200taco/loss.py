"""``Tacotron2Loss`` of the reference (train.py:126-217) with its Python B x S loop on the device.

Same constructor, attributes (``current_attention_weight``, ``global_step``, ``current_sigma`` ...), call signature and
4-tuple result as the reference class, so ``train.py`` can use it unchanged (``criterion = Tacotron2Loss()``, train.py:333).
The loss terms are small tensor expressions (``loss_terms`` below, device agnostic); the expensive part of the reference is
``create_diagonal_attention_target``, a Python double loop over B x T_dec with ``.item()`` syncs (51,200 iterations at B=64
and 800 frames: seconds per step).  Here the target is ONE kernel launch (``b200taco_guided_attention_target``, csrc/loss.cu),
and ``Tacotron2Loss.forward`` itself is three launches (``b200taco_loss_fwd``: masked L1 x 2 + BCE, KL + entropy with the
target recomputed on the fly, a deterministic finishing reduction that also applies the entropy-adaptive weight) that leave
every scalar ON THE DEVICE: no ``.item()`` / ``float()`` host round trip anywhere in the step (``current_attention_weight``
and ``current_sigma`` are read back lazily, only when somebody looks at them).  Gradients: ``b200taco_loss_bwd``.
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


def gate_targets(mel_lengths: torch.Tensor, num_frames: int) -> torch.Tensor:
    """(B, num_frames) stop-token targets of the training loop (train.py:505-508: ones from frame ``mel_lengths[b] - 1`` on)
    as one launch instead of a Python loop with a ``.item()`` per utterance."""
    if mel_lengths.device.type != "cuda":
        raise B200TacoError(f"b200taco.loss runs on CUDA tensors only (got {mel_lengths.device}); there is no CPU fallback")
    ml = mel_lengths.to(dtype=torch.int64).contiguous()
    out = torch.empty(ml.shape[0], int(num_frames), dtype=torch.float32, device=ml.device)
    with torch.cuda.device(ml.device):
        check(_cabi.lib().b200taco_gate_targets(ml.data_ptr(), ml.shape[0], int(num_frames), out.data_ptr(),
                                                C.c_void_p(torch.cuda.current_stream(ml.device).cuda_stream)), "b200taco_gate_targets")
    return out


def alignment_block(alignments) -> torch.Tensor:
    """The list of per-step (B, T) attention weights as ONE contiguous (S, B, T) tensor: the buffer the decoder wrote them into
    when the list is its ``unbind(0)`` (b200taco.model returns exactly that; gradients then flow to the buffer directly), a
    ``torch.stack`` otherwise."""
    if torch.is_tensor(alignments):
        return alignments.contiguous()
    a0 = alignments[0]
    base = a0._base
    if (base is not None and base.dim() == 3 and base.shape[0] == len(alignments) and base.shape[1:] == a0.shape
            and base.is_contiguous() and all(a._base is base for a in alignments)
            and alignments[-1].data_ptr() == base.data_ptr() + (len(alignments) - 1) * base.stride(0) * base.element_size()):
        return base
    return torch.stack(list(alignments), dim=0)


class _FusedLoss(torch.autograd.Function):
    """b200taco_loss_fwd / b200taco_loss_bwd (csrc/loss.cu).  Result: 8 device scalars
    [mel, gate, kl, entropy, attention weight, total, kl before the clamp, valid mel bins]."""

    @staticmethod
    def forward(ctx, mel_post, mel_coarse, gate_logits, align, mel_target, gate_target, mel_lengths, text_lengths, sigma, consts):
        dev = mel_coarse.device
        f32 = dict(dtype=torch.float32, device=dev)
        mel_post, mel_coarse = mel_post.to(**f32).contiguous(), mel_coarse.to(**f32).contiguous()
        gate_logits, gate_target = gate_logits.to(**f32).contiguous(), gate_target.to(**f32).contiguous()
        mel_target = mel_target.to(**f32).contiguous()
        mel_lengths = mel_lengths.to(device=dev, dtype=torch.int64).contiguous()
        B, S = gate_logits.shape
        if mel_target.shape != (B, config.n_mels, S) or mel_coarse.shape != (B, S, config.n_mels) or mel_post.shape != mel_coarse.shape:
            raise B200TacoError(f"loss: mel shapes {tuple(mel_post.shape)}, {tuple(mel_coarse.shape)}, target {tuple(mel_target.shape)} "
                                f"do not match gate {tuple(gate_logits.shape)}")
        T = 0
        if align is not None:
            align = align.to(**f32).contiguous()
            T = align.shape[2]
            if align.shape[:2] != (S, B):
                raise B200TacoError(f"loss: alignments {tuple(align.shape)} do not match (S, B) = ({S}, {B})")
            text_lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
            sigma = sigma.to(**f32).contiguous()
        lib = _cabi.lib()
        scratch = torch.empty(max(1, int(lib.b200taco_loss_scratch_bytes(B, S)) // 4), **f32)
        out = torch.empty(8, **f32)
        w_start, w_min, ent_target, w_current = consts
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            check(lib.b200taco_loss_fwd(mel_post.data_ptr(), mel_coarse.data_ptr(), mel_target.data_ptr(), gate_logits.data_ptr(),
                                        gate_target.data_ptr(), mel_lengths.data_ptr(), ptr(align), ptr(text_lengths) if align is not None else None,
                                        ptr(sigma) if align is not None else None, B, S, T, w_start, w_min, ent_target, w_current,
                                        scratch.data_ptr(), out.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                  "b200taco_loss_fwd")
        ctx.save_for_backward(mel_post, mel_coarse, gate_logits, align, mel_target, gate_target, mel_lengths, text_lengths, sigma, out)
        ctx.dims = (B, S, T)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        mel_post, mel_coarse, gate_logits, align, mel_target, gate_target, mel_lengths, text_lengths, sigma, out = ctx.saved_tensors
        B, S, T = ctx.dims
        dev = mel_coarse.device
        grad_out = grad_out.to(dtype=torch.float32, device=dev).contiguous()
        g_post, g_coarse, g_gate = torch.empty_like(mel_post), torch.empty_like(mel_coarse), torch.empty_like(gate_logits)
        g_align = torch.empty_like(align) if align is not None and ctx.needs_input_grad[3] else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            check(_cabi.lib().b200taco_loss_bwd(mel_post.data_ptr(), mel_coarse.data_ptr(), mel_target.data_ptr(), gate_logits.data_ptr(),
                                                gate_target.data_ptr(), mel_lengths.data_ptr(), ptr(align),
                                                ptr(text_lengths) if align is not None else None, ptr(sigma) if align is not None else None,
                                                B, S, T, out.data_ptr(), grad_out.data_ptr(), g_post.data_ptr(), g_coarse.data_ptr(),
                                                g_gate.data_ptr(), ptr(g_align), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                  "b200taco_loss_bwd")
        return g_post, g_coarse, g_gate, g_align, None, None, None, None, None, None


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
        self.fused = True          # CUDA inputs: whole forward as three launches, no host sync (False: tensor expressions)

    # ``current_attention_weight`` / ``current_sigma`` are plain floats in the reference (train.py:137, :157, :207); here they may
    # be device scalars that are only copied back when somebody reads them (train.py:526 does, once per logging step)
    @property
    def current_attention_weight(self):
        w = self._attention_weight
        return float(w) if torch.is_tensor(w) else w

    @current_attention_weight.setter
    def current_attention_weight(self, value):
        self._attention_weight = value

    @property
    def current_sigma(self):
        s = self.__dict__.get("_sigma")
        return float(s) if torch.is_tensor(s) else s

    @current_sigma.setter
    def current_sigma(self, value):
        self._sigma = value

    def get_mask_from_lengths(self, lengths):
        """True where padding (train.py:143-147)."""
        frames = torch.arange(int(lengths.max().item()), device=lengths.device)
        return frames[None, :] >= lengths[:, None]

    def create_diagonal_attention_target(self, text_lengths, num_steps, alignments):
        """train.py:148-171 as one kernel launch."""
        sigma = annealed_sigma(text_lengths, self.global_step, self.sigma_warmup_steps)
        self.current_sigma = sigma.mean()
        return guided_attention_target(text_lengths.to(alignments[0].device), sigma, num_steps, int(text_lengths.max().item()))

    def attention_weight(self, entropy) -> float:
        """Entropy-adaptive weight of the guided-attention term (train.py:205-210)."""
        if entropy is None or float(entropy) > self.entropy_target:
            return self.attn_weight_start
        ratio = max(0.0, float(entropy) / self.entropy_target)
        return max(self.min_attn_weight, self.attn_weight_start * ratio)

    def _forward_fused(self, model_outputs, targets, text_lengths):
        mel_post, mel_coarse, gate_logits, alignments = model_outputs
        mel_target, gate_target, mel_lengths = targets
        align = sigma = None
        if len(alignments) > 1 and text_lengths is not None:
            align = alignment_block(alignments)
            sigma = annealed_sigma(text_lengths.to(mel_coarse.device), self.global_step, self.sigma_warmup_steps)
            self.current_sigma = sigma.mean()
        w = self._attention_weight
        if torch.is_tensor(w):      # the weight of a step WITHOUT attention term is the previous step's (train.py:195-214): rare, sync
            w = float(w)
        out = _FusedLoss.apply(mel_post, mel_coarse, gate_logits, align, mel_target, gate_target, mel_lengths, text_lengths, sigma,
                               (self.attn_weight_start, self.min_attn_weight, self.entropy_target, w))
        if align is not None:
            self.current_attention_weight = out[4].detach()
        self.global_step += 1
        return out[5], out[0], out[1], out[2]

    def forward(self, model_outputs, targets, text_lengths=None):
        if self.fused and model_outputs[1].is_cuda:
            return self._forward_fused(model_outputs, targets, text_lengths)
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
