"""Training-mode stages as ``torch.autograd.Function``s over the C ABI (work in progress, stage by stage).

Each stage's forward keeps what its hand-written CUDA backward needs in a "saved" device buffer
allocated here by PyTorch; gradients come back in the reference's parameter layouts, so autograd
hands them to the optimiser (or to DDP's all-reduce hooks) unchanged.  Dropout masks are drawn with
PyTorch's generator on the device (0/1 floats) and passed down; BatchNorm running statistics are
updated here from the batch statistics the kernels leave in the saved buffer
(momentum 0.1, unbiased variance, as nn.BatchNorm1d does).

Built so far: postnet (src/model.py:392-407 + residual :476-477), encoder (src/model.py:50-77).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, config
from ._cabi import check
from .engine import _WEIGHT_MAP, Engine

POST_P = 0.5              # F.dropout(..., 0.5, self.training), src/model.py:402-405


def _weights_struct(tensors: dict, only_prefix: str = "") -> tuple:
    """b200taco_weights over the given {state_dict key: tensor}; missing keys stay null."""
    w = _cabi.Weights()
    keep = []
    for key, field, idx in _WEIGHT_MAP:
        if key not in tensors or not key.startswith(only_prefix):
            continue
        t = tensors[key]
        if not t.is_contiguous():
            t = t.contiguous()
        keep.append(t)
        if idx is None:
            setattr(w, field, t.data_ptr())
        else:
            getattr(w, field)[idx] = t.data_ptr()
    return w, keep


def _post_dims():
    n = config.postnet_n_convolutions
    d = config.postnet_embedding_dim
    cin = [config.n_mels] + [d] * (n - 1)
    cout = [d] * (n - 1) + [config.n_mels]
    return cin, cout


def postnet_param_keys():
    keys = []
    for i in range(config.postnet_n_convolutions):
        p = f"postnet.convolutions.{i}"
        keys += [p + ".0.weight", p + ".0.bias", p + ".1.weight", p + ".1.bias"]
    return keys


class PostnetTrainFn(torch.autograd.Function):
    """post = coarse + PostNet(coarse) in train mode.  Inputs: engine, coarse (B,S,80), masks (list of 5 (B*S,c_out)
    0/1 float tensors or None), then the 20 postnet parameters in ``postnet_param_keys()`` order."""

    @staticmethod
    def forward(ctx, eng: Engine, coarse, masks, *params):
        lib = eng._lib
        coarse = coarse.contiguous()
        B, S, _ = coarse.shape
        keys = postnet_param_keys()
        tensors = dict(zip(keys, [p.detach() for p in params]))
        w, keep = _weights_struct(tensors, "postnet.")
        saved = torch.empty(lib.b200taco_postnet_train_saved_bytes(B, S) // 4, dtype=torch.float32, device=coarse.device)
        scratch = torch.empty(lib.b200taco_postnet_train_scratch_bytes(B, S) // 4, dtype=torch.float32, device=coarse.device)
        post = torch.empty_like(coarse)
        mptr = None
        if masks is not None:
            masks = [m.contiguous() for m in masks]
            mptr = (C.c_void_p * len(masks))(*[m.data_ptr() for m in masks])
        scale = 1.0 / (1.0 - POST_P) if masks is not None else 1.0
        with torch.cuda.device(coarse.device):
            check(lib.b200taco_postnet_train_fwd(eng._h, C.byref(w), coarse.data_ptr(), B, S, mptr, C.c_float(scale),
                                                 post.data_ptr(), saved.data_ptr(), saved.numel() * 4, scratch.data_ptr(),
                                                 scratch.numel() * 4, eng._stream()), "b200taco_postnet_train_fwd")
        ctx.eng, ctx.masks, ctx.scale, ctx.keys = eng, masks, scale, keys
        ctx.save_for_backward(coarse, saved, *[p.detach() for p in params])
        ctx.bn_stats = PostnetTrainFn._bn_stats(saved, B * S)
        return post

    @staticmethod
    def _bn_stats(saved, n):
        """[(batch mean, biased batch var)] per layer, views into the saved buffer."""
        _, cout = _post_dims()
        out, off = [], 0
        for c in cout:
            off += 2 * n * c
            out.append((saved[off:off + c], saved[off + c:off + 2 * c]))
            off += 3 * c
        return out

    @staticmethod
    def backward(ctx, dpost):
        eng, lib = ctx.eng, ctx.eng._lib
        coarse, saved, *params = ctx.saved_tensors
        B, S, _ = coarse.shape
        dpost = dpost.contiguous()
        tensors = dict(zip(ctx.keys, params))
        w, keep = _weights_struct(tensors, "postnet.")
        grads = {k: torch.empty_like(t) for k, t in tensors.items()}
        g, keep2 = _weights_struct(grads, "postnet.")
        dcoarse = torch.empty_like(coarse)
        scratch = torch.empty(lib.b200taco_postnet_train_scratch_bytes(B, S) // 4, dtype=torch.float32, device=coarse.device)
        mptr = None
        if ctx.masks is not None:
            mptr = (C.c_void_p * len(ctx.masks))(*[m.data_ptr() for m in ctx.masks])
        with torch.cuda.device(coarse.device):
            check(lib.b200taco_postnet_train_bwd(eng._h, C.byref(w), coarse.data_ptr(), B, S, mptr, C.c_float(ctx.scale),
                                                 saved.data_ptr(), dpost.data_ptr(), dcoarse.data_ptr(), C.byref(g),
                                                 scratch.data_ptr(), scratch.numel() * 4, eng._stream()),
                  "b200taco_postnet_train_bwd")
        return (None, dcoarse, None) + tuple(grads[k] for k in ctx.keys)


def encoder_param_keys():
    keys = ["encoder.embedding.weight"]
    for i in range(config.encoder_n_convolutions):
        p = f"encoder.convolutions.{i}"
        keys += [p + ".0.weight", p + ".0.bias", p + ".1.weight", p + ".1.bias"]
    for s in ("", "_reverse"):
        keys += [f"encoder.lstm.weight_ih_l0{s}", f"encoder.lstm.weight_hh_l0{s}", f"encoder.lstm.bias_ih_l0{s}",
                 f"encoder.lstm.bias_hh_l0{s}"]
    return keys


class EncoderTrainFn(torch.autograd.Function):
    """memory (B,T,512) = Encoder(ids) in train mode (batch-statistics BatchNorm; the reference's encoder has no
    dropout, F3).  Inputs: engine, ids (B,T) int64, then the 21 encoder parameters in ``encoder_param_keys()`` order."""

    @staticmethod
    def forward(ctx, eng: Engine, ids, *params):
        lib = eng._lib
        ids = ids.to(dtype=torch.int64).contiguous()
        B, T = ids.shape
        dev = ids.device
        keys = encoder_param_keys()
        tensors = dict(zip(keys, [p.detach() for p in params]))
        w, keep = _weights_struct(tensors, "encoder.")
        saved = torch.empty(lib.b200taco_encoder_train_saved_bytes(B, T) // 4, dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.b200taco_encoder_train_scratch_bytes(B, T) // 4, dtype=torch.float32, device=dev)
        memory = torch.empty(B, T, config.encoder_embedding_dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.b200taco_encoder_train_fwd(eng._h, C.byref(w), ids.data_ptr(), B, T, memory.data_ptr(), saved.data_ptr(),
                                                 saved.numel() * 4, scratch.data_ptr(), scratch.numel() * 4, eng._stream()),
                  "b200taco_encoder_train_fwd")
        ctx.eng, ctx.keys = eng, keys
        ctx.save_for_backward(ids, saved, *[p.detach() for p in params])
        n, d = B * T, config.encoder_embedding_dim
        ctx.bn_stats = []
        off = n * d
        for _ in range(config.encoder_n_convolutions):
            off += 2 * n * d
            ctx.bn_stats.append((saved[off:off + d], saved[off + d:off + 2 * d]))
            off += 3 * d
        return memory

    @staticmethod
    def backward(ctx, dmemory):
        eng, lib = ctx.eng, ctx.eng._lib
        ids, saved, *params = ctx.saved_tensors
        B, T = ids.shape
        dmemory = dmemory.contiguous()
        tensors = dict(zip(ctx.keys, params))
        w, keep = _weights_struct(tensors, "encoder.")
        grads = {k: torch.empty_like(t) for k, t in tensors.items()}
        g, keep2 = _weights_struct(grads, "encoder.")
        scratch = torch.empty(lib.b200taco_encoder_train_scratch_bytes(B, T) // 4, dtype=torch.float32, device=ids.device)
        with torch.cuda.device(ids.device):
            check(lib.b200taco_encoder_train_bwd(eng._h, C.byref(w), ids.data_ptr(), B, T, saved.data_ptr(), dmemory.data_ptr(),
                                                 C.byref(g), scratch.data_ptr(), scratch.numel() * 4, eng._stream()),
                  "b200taco_encoder_train_bwd")
        return (None, None) + tuple(grads[k] for k in ctx.keys)


def draw_postnet_masks(B, S, device, generator=None):
    _, cout = _post_dims()
    return [(torch.rand(B * S, c, device=device, generator=generator) >= POST_P).float() for c in cout]


@torch.no_grad()
def update_running_stats(bn: torch.nn.BatchNorm1d, mean, var_biased, n: int):
    """nn.BatchNorm1d train-mode side effect: momentum update with the UNBIASED batch variance."""
    m = bn.momentum if bn.momentum is not None else 0.1
    bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
    bn.running_var.mul_(1 - m).add_(var_biased * (n / max(1, n - 1)), alpha=m)
    bn.num_batches_tracked += 1
