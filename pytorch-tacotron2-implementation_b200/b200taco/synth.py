"""Synthetic inputs and weight surgeries for benchmarks and parity tests.

Everything here works on a plain ``state_dict`` (name -> tensor), so the same surgery can
be applied to the reference model, to the CPU oracle and to :class:`b200taco.model.Tacotron2`.
There is no network for datasets or checkpoints: phoneme ids are seeded random draws and
weights are the seeded default initialisation, optionally modified so that the attention
path and the stop gate are actually exercised (F11, F12).
"""
from __future__ import annotations

import math

import torch

from . import config

WEIGHT_SEED = 1234          # train.py:300
INPUT_SEED = 0

# Tensors the CUDA decoder keeps in bf16 (the two LSTMs' matrices: 98 % of the bytes a
# decoder step reads).  Parity tests round exactly these in the oracle's copy.
BF16_WEIGHT_KEYS = (
    "decoder.attention_lstm.weight_ih", "decoder.attention_lstm.weight_hh",
    "decoder.decoder_lstm.weight_ih", "decoder.decoder_lstm.weight_hh",
)


def phoneme_ids(B: int, T: int, seed: int = INPUT_SEED) -> torch.Tensor:
    """(B,T) int64 ids uniform over the 72 symbols (src/config.py:14-23)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, len(config.SYMBOLS), (B, T), generator=g, dtype=torch.int64)


def mel_targets(B: int, S: int, seed: int = INPUT_SEED) -> torch.Tensor:
    """(B,80,S) log-mel-like targets: N(-5, 2^2) clipped to [ln 1e-5, 2] (src/audio.py:46-47)."""
    g = torch.Generator().manual_seed(seed + 1)
    m = torch.randn(B, config.n_mels, S, generator=g) * 2.0 - 5.0
    return m.clamp_(math.log(1e-5), 2.0)


def c1_batch():
    """BASELINE config 1: B=4, lengths [37,29,21,12] padded to 37 with id 0, 120 target frames."""
    lengths = torch.tensor([37, 29, 21, 12], dtype=torch.int64)
    ids = phoneme_ids(4, 37)
    for b, n in enumerate(lengths.tolist()):
        ids[b, n:] = 0
    return ids, lengths, mel_targets(4, 120)


def round_to_bf16_(sd: dict, keys=BF16_WEIGHT_KEYS) -> dict:
    """Make the listed tensors bf16-representable (value-preserving for the bf16 kernels)."""
    for k in keys:
        sd[k].copy_(sd[k].to(torch.bfloat16).to(sd[k].dtype))
    return sd


def stress_init_(sd: dict, scale: float = 2.5, energy_scale: float = 4.0) -> dict:
    """F12: default init barely exercises attention; scale the decoder's non-LSTM matrices.

    Gives attention peaks ~0.17 instead of ~1/T while staying non-chaotic, so that wrong
    taps / wrong channel order / a dropped query term show up at 1e-3 instead of 1e-7.
    """
    for k, v in sd.items():
        if k.startswith("decoder.") and v.dim() == 2 and "_lstm." not in k:
            v.mul_(scale)
    sd["decoder.attention.energy_scale"].fill_(energy_scale)
    return sd


def plant_stop_counter_(sd: dict, unit: int = 7, rate: float = 0.0015, jitter: float = 0.003,
                        gain: float = 8.0, seed: int = 99, ctx_center=None) -> dict:
    """F11: make the stop gate fire at an utterance-dependent step (default init never fires).

    Decoder-LSTM unit ``unit`` is turned into an integrator: its i/f/o gates are pinned
    open (bias +20), its g pre-activation reads only the context vector through a small
    random probe plus ``rate``, so c_j grows by ~``rate`` + probe.ctx per step.  The gate
    layer reads only that unit: logit = gain * tanh(c_j) - 3, which crosses 0 at
    c_j = atanh(3/gain).  The slope per step (~gain*rate) is far above fp32 noise, so stop
    indices are exact-comparable as long as that unit's path stays fp32-accurate.

    ``ctx_center`` (512,), e.g. ``memory.mean((0, 1))``: the context vectors of all utterances
    share a large common component; subtracting its probe response from the bias leaves
    only the utterance-specific part, which spreads the stop steps (rate=8e-4, jitter=6e-3
    gives ~[300, 900] frames at B=32, T=150).
    """
    H = config.decoder_rnn_dim
    w_ih, w_hh = sd["decoder.decoder_lstm.weight_ih"], sd["decoder.decoder_lstm.weight_hh"]
    b_ih, b_hh = sd["decoder.decoder_lstm.bias_ih"], sd["decoder.decoder_lstm.bias_hh"]
    rows = [g * H + unit for g in range(4)]            # i, f, g, o
    for r in rows:
        w_ih[r].zero_(); w_hh[r].zero_(); b_hh[r] = 0.0; b_ih[r] = 0.0
    for r in (rows[0], rows[1], rows[3]):
        b_ih[r] = 20.0
    gen = torch.Generator().manual_seed(seed)
    probe = torch.randn(config.encoder_embedding_dim, generator=gen) * jitter
    # decoder_lstm input = [attention_hidden (1024), context (512)]  (src/model.py:265)
    w_ih[rows[2], config.attention_rnn_dim:] = probe.to(w_ih.dtype)
    b_ih[rows[2]] = rate
    if ctx_center is not None:
        b_ih[rows[2]] -= float(probe.double() @ ctx_center.detach().double().cpu())
    gw = sd["decoder.gate_layer.weight"]
    gw.zero_()
    gw[0, unit] = gain
    sd["decoder.gate_layer.bias"].fill_(-3.0)
    return sd


def seeded_state_dict(model_ctor, seed: int = WEIGHT_SEED) -> dict:
    """``torch.manual_seed(seed); model_ctor().state_dict()`` detached onto the CPU in fp32."""
    torch.manual_seed(seed)
    m = model_ctor()
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
