"""Training-mode stages as ``torch.autograd.Function``s over the C ABI (work in progress, stage by stage).

Each stage's forward keeps what its hand-written CUDA backward needs in a "saved" device buffer
allocated here by PyTorch; gradients come back in the reference's parameter layouts, so autograd
hands them to the optimiser (or to DDP's all-reduce hooks) unchanged.  Dropout masks are drawn with
PyTorch's generator on the device (0/1 floats) and passed down; BatchNorm running statistics are
updated here from the batch statistics the kernels leave in the saved buffer
(momentum 0.1, unbiased variance, as nn.BatchNorm1d does).

Built so far: postnet (src/model.py:392-407 + residual :476-477), encoder (src/model.py:50-77), teacher-forced
decoder (src/model.py:246-297).
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
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)     # train.py runs the model under autocast
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
        # clones: the engine must not keep the multi-hundred-MB `saved` buffer alive after the backward
        eng._post_bn_stats = [(m.clone(), v.clone()) for m, v in ctx.bn_stats]
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
    @torch.amp.custom_bwd(device_type="cuda")
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
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
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
        eng._enc_bn_stats = [(m.clone(), v.clone()) for m, v in ctx.bn_stats]
        return memory

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
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


PRENET_P = 0.5            # PreNet dropout, src/model.py:170
LSTM_P = config.p_attention_dropout if hasattr(config, "p_attention_dropout") else 0.1   # src/model.py:262, :270


def decoder_param_keys():
    d = "decoder."
    keys = [d + "prenet.layers.0.weight", d + "prenet.layers.1.weight"]
    keys += [d + "attention." + k for k in ("query_layer.weight", "memory_layer.weight", "location_conv.weight",
                                            "location_dense.weight", "v.weight", "v.bias", "energy_scale")]
    for m in ("attention_lstm", "decoder_lstm"):
        keys += [d + f"{m}.{k}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    keys += [d + "linear_projection.weight", d + "linear_projection.bias", d + "gate_layer.weight", d + "gate_layer.bias"]
    return keys


class DecoderTrainFn(torch.autograd.Function):
    """(coarse (B,S,80), gate (B,S), align (S,B,T)) = Decoder.forward in train mode.  Inputs: engine (weights packed
    from the CURRENT parameters), memory (B,T,512), mel_targets (B,80,S), text_lengths (B,) or None, masks dict
    (p1, p2: (S*B,256); a, d: (S,B,1024)) or None, then the 21 decoder parameters in ``decoder_param_keys()`` order."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, eng: Engine, memory, mel_targets, text_lengths, masks, *params):
        lib = eng._lib
        memory = memory.contiguous()
        mel_targets = mel_targets.to(dtype=torch.float32).contiguous()
        B, T, _ = memory.shape
        S = mel_targets.shape[2]
        dev = memory.device
        tl = text_lengths.to(device=dev, dtype=torch.int64).contiguous() if text_lengths is not None else None
        saved = torch.empty(lib.b200taco_decoder_train_saved_bytes(B, T, S) // 4, dtype=torch.float32, device=dev)
        mel = torch.empty(B, S, config.n_mels, dtype=torch.float32, device=dev)
        gate = torch.empty(B, S, dtype=torch.float32, device=dev)
        align = torch.empty(S, B, T, dtype=torch.float32, device=dev)
        ws, wsz = eng._workspace(B, T, S)
        if masks is not None:
            masks = {k: v.contiguous() for k, v in masks.items()}
        mp = lambda k: masks[k].data_ptr() if masks is not None else None
        ps = 1.0 / (1.0 - PRENET_P) if masks is not None else 1.0
        ls = 1.0 / (1.0 - LSTM_P) if masks is not None else 1.0
        with torch.cuda.device(dev):
            check(lib.b200taco_decoder_train_fwd(eng._h, memory.data_ptr(), mel_targets.data_ptr(),
                                                 tl.data_ptr() if tl is not None else None, B, T, S,
                                                 mp("p1"), mp("p2"), C.c_float(ps), mp("a"), mp("d"), C.c_float(ls),
                                                 mel.data_ptr(), gate.data_ptr(), align.data_ptr(), saved.data_ptr(),
                                                 saved.numel() * 4, ws, wsz, eng._stream()), "b200taco_decoder_train_fwd")
        ctx.eng, ctx.masks, ctx.scales, ctx.tl, ctx.keys = eng, masks, (ps, ls), tl, decoder_param_keys()
        ctx.save_for_backward(memory, saved, align, *[p.detach() for p in params])
        return mel, gate, align

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dmel, dgate, dalign):
        eng, lib = ctx.eng, ctx.eng._lib
        memory, saved, align, *params = ctx.saved_tensors
        B, T, _ = memory.shape
        S = align.shape[0]
        dev = memory.device
        dmel = (dmel if dmel is not None else torch.zeros(B, S, config.n_mels, device=dev)).contiguous()
        dgate = (dgate if dgate is not None else torch.zeros(B, S, device=dev)).contiguous()
        dalign = dalign.contiguous() if dalign is not None else None
        tensors = dict(zip(ctx.keys, params))
        w, keep = _weights_struct(tensors, "decoder.")
        grads = {k: torch.empty_like(t) for k, t in tensors.items()}
        g, keep2 = _weights_struct(grads, "decoder.")
        dmemory = torch.empty_like(memory)
        scratch = torch.empty(lib.b200taco_decoder_train_scratch_bytes(B, T, S) // 4, dtype=torch.float32, device=dev)
        masks, (ps, ls), tl = ctx.masks, ctx.scales, ctx.tl
        mp = lambda k: masks[k].data_ptr() if masks is not None else None
        with torch.cuda.device(dev):
            check(lib.b200taco_decoder_train_bwd(eng._h, C.byref(w), memory.data_ptr(), tl.data_ptr() if tl is not None else None,
                                                 B, T, S, mp("p1"), mp("p2"), C.c_float(ps), mp("a"), mp("d"), C.c_float(ls),
                                                 saved.data_ptr(), align.data_ptr(), dmel.data_ptr(), dgate.data_ptr(),
                                                 dalign.data_ptr() if dalign is not None else None, dmemory.data_ptr(),
                                                 C.byref(g), scratch.data_ptr(), scratch.numel() * 4, eng._stream()),
                  "b200taco_decoder_train_bwd")
        return (None, dmemory, None, None, None) + tuple(grads[k] for k in ctx.keys)


def draw_decoder_masks(B, S, device, generator=None):
    r = lambda *shape, p: (torch.rand(*shape, device=device, generator=generator) >= p).float()
    return {"p1": r(S * B, config.prenet_dim, p=PRENET_P), "p2": r(S * B, config.prenet_dim, p=PRENET_P),
            "a": r(S, B, config.attention_rnn_dim, p=LSTM_P), "d": r(S, B, config.decoder_rnn_dim, p=LSTM_P)}


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
