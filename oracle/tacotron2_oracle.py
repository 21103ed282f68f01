"""CPU oracle for the Tacotron 2 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  Nothing under
``pytorch-tacotron2-implementation_b200/`` imports it, and the product path raises if
the CUDA library is missing instead of coming here.

What it is: a functional (state_dict in, tensors out) restatement of the reference's
eval-mode arithmetic for ``Tacotron2.forward`` / ``Tacotron2.inference``
(/root/reference/src/model.py:454, :482), written against plain torch CPU tensor ops.
The reference's own arithmetic lives in a third-party dependency that is not under
/root/reference: PyTorch (pinned ``torch==2.6.0`` in requirements.txt:1; this image has
2.11.0).  The restatement therefore spells the published semantics of ``nn.Conv1d``,
``nn.BatchNorm1d`` (eval), ``nn.LSTM``/``nn.LSTMCell`` (gate order i,f,g,o),
``nn.Linear``, ``F.softmax`` and ``torch.bmm`` out by hand, and is anchored on the
reference's call sites, cited per function below.

Parity status: the reference ships NO tests, golden vectors or fixtures (SURVEY.md
section 4), so the reference itself does not pin results.  The oracle is pinned instead
against outputs of the reference *run in the build container*: ``tests/golden/*.npz``
were produced by ``tests/golden/make_golden.py`` importing /root/reference/src/model.py,
and ``tests/test_oracle_golden.py`` holds the oracle to those vectors (fp32: <= 2e-6).

Works in float32 (timing / parity with the reference) and float64 (truth for
tolerance analysis): the dtype follows the state_dict.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm1d default, used at src/model.py:36 and :357
MASK_FILL = -1e9       # src/model.py:132


# --------------------------------------------------------------------------- helpers
def _bn_eval(x, sd, prefix):
    """Eval-mode BatchNorm1d over (B, C, T): running statistics (src/model.py:36)."""
    mean = sd[prefix + ".running_mean"][None, :, None]
    var = sd[prefix + ".running_var"][None, :, None]
    w = sd[prefix + ".weight"][None, :, None]
    b = sd[prefix + ".bias"][None, :, None]
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def _lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTMCell: z = W_ih x + b_ih + W_hh h + b_hh, gates i,f,g,o (src/model.py:195-203)."""
    z = torch.addmm(b_ih, x, w_ih.t()) + torch.addmm(b_hh, h, w_hh.t())
    i, f, g, o = z.chunk(4, dim=1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


# --------------------------------------------------------------------------- encoder
def encoder_conv_stack(sd, ids):
    """Embedding + 3x relu(BN(conv1d k5 p2)); returns (B, 512, T).  src/model.py:59-67."""
    x = sd["encoder.embedding.weight"][ids]            # (B, T, 512)
    x = x.transpose(1, 2)
    n_conv = 0
    while f"encoder.convolutions.{n_conv}.0.weight" in sd:
        n_conv += 1
    for i in range(n_conv):
        p = f"encoder.convolutions.{i}"
        w = sd[p + ".0.weight"]
        x = F.conv1d(x, w, sd[p + ".0.bias"], padding=(w.shape[2] - 1) // 2)
        x = torch.relu(_bn_eval(x, sd, p + ".1"))
    return x


def encoder_bilstm(sd, x):
    """1-layer BiLSTM over the full padded length, zero initial state, no packing.

    x: (B, T, 512) -> (B, T, 512) = cat[forward, backward].  src/model.py:42-48, :75.
    """
    B, T, _ = x.shape
    outs = []
    for suffix, order in (("", range(T)), ("_reverse", range(T - 1, -1, -1))):
        w_ih = sd["encoder.lstm.weight_ih_l0" + suffix]
        w_hh = sd["encoder.lstm.weight_hh_l0" + suffix]
        b = sd["encoder.lstm.bias_ih_l0" + suffix] + sd["encoder.lstm.bias_hh_l0" + suffix]
        xin = torch.addmm(b, x.reshape(B * T, -1), w_ih.t()).view(B, T, -1)
        H = w_hh.shape[1]
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        out = x.new_zeros(B, T, H)
        for t in order:
            z = xin[:, t] + h @ w_hh.t()
            i, f, g, o = z.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            out[:, t] = h
        outs.append(out)
    return torch.cat(outs, dim=2)


def encoder(sd, ids):
    """Encoder.forward, src/model.py:50-77: ids (B,T) int64 -> memory (B,T,512)."""
    return encoder_bilstm(sd, encoder_conv_stack(sd, ids).transpose(1, 2))


def make_pad_mask(lengths, max_len):
    """True where PAD.  src/model.py:443-450."""
    return torch.arange(max_len)[None, :] >= lengths[:, None]


# --------------------------------------------------------------------------- decoder
class DecoderState:
    """All recurrent state of one decode (src/model.py:219-234 and :104-115)."""

    def __init__(self, sd, memory, mask=None):
        B, T, _ = memory.shape
        z = lambda n: memory.new_zeros(B, n)
        self.memory = memory
        self.mask = mask
        # LocationSensitiveAttention.init_states, src/model.py:104-115
        self.processed_memory = memory @ sd["decoder.attention.memory_layer.weight"].t()
        self.prev_attn = z(T)
        self.cum_attn = z(T)
        # Decoder._initialize_decoder_states, src/model.py:219-234
        self.x = z(sd["decoder.prenet.layers.0.weight"].shape[1])      # go frame
        self.ah, self.ac = z(1024), z(1024)
        self.dh, self.dc = z(1024), z(1024)
        self.ctx = z(memory.shape[2])
        H = sd["decoder.attention_lstm.weight_hh"].shape[1]
        if H != 1024:
            self.ah, self.ac, self.dh, self.dc = z(H), z(H), z(H), z(H)


def prenet(sd, x):
    """PreNet.forward in eval mode (dropout off, src/model.py:166-171)."""
    i = 0
    while f"decoder.prenet.layers.{i}.weight" in sd:
        x = torch.relu(x @ sd[f"decoder.prenet.layers.{i}.weight"].t())
        i += 1
    return x


def attention_energies(sd, st, query):
    """LocationSensitiveAttention.get_alignment_energies, src/model.py:117-133."""
    a = "decoder.attention."
    pq = query @ sd[a + "query_layer.weight"].t()                      # (B, 128)
    loc_in = torch.stack([st.prev_attn, st.cum_attn], dim=1)           # (B, 2, T)
    w_loc = sd[a + "location_conv.weight"]
    loc = F.conv1d(loc_in, w_loc, None, padding=(w_loc.shape[2] - 1) // 2)   # (B, 32, T)
    lf = loc.transpose(1, 2) @ sd[a + "location_dense.weight"].t()     # (B, T, 128)
    t = torch.tanh(pq[:, None, :] + st.processed_memory + lf)
    e = (t @ sd[a + "v.weight"].t()).squeeze(-1) + sd[a + "v.bias"]
    e = e * sd[a + "energy_scale"]
    if st.mask is not None:
        e = e.masked_fill(st.mask, MASK_FILL)
    return e


def decode_step(sd, st):
    """Decoder._decode_step in eval mode, src/model.py:246-275.  Mutates ``st``."""
    d = "decoder."
    p = prenet(sd, st.x)
    st.ah, st.ac = _lstm_cell(torch.cat([p, st.ctx], dim=1), st.ah, st.ac,
                              sd[d + "attention_lstm.weight_ih"], sd[d + "attention_lstm.weight_hh"],
                              sd[d + "attention_lstm.bias_ih"], sd[d + "attention_lstm.bias_hh"])
    # LocationSensitiveAttention.forward, src/model.py:135-148
    e = attention_energies(sd, st, st.ah)
    w = torch.softmax(e, dim=1)
    st.prev_attn = w
    st.cum_attn = st.cum_attn + w
    st.ctx = torch.bmm(w[:, None, :], st.memory).squeeze(1)
    st.dh, st.dc = _lstm_cell(torch.cat([st.ah, st.ctx], dim=1), st.dh, st.dc,
                              sd[d + "decoder_lstm.weight_ih"], sd[d + "decoder_lstm.weight_hh"],
                              sd[d + "decoder_lstm.bias_ih"], sd[d + "decoder_lstm.bias_hh"])
    proj_in = torch.cat([st.dh, st.ctx], dim=1)
    mel = torch.addmm(sd[d + "linear_projection.bias"], proj_in, sd[d + "linear_projection.weight"].t())
    gate = torch.addmm(sd[d + "gate_layer.bias"], proj_in, sd[d + "gate_layer.weight"].t())
    return mel, gate, w


def decoder_teacher_forced(sd, memory, mel_targets, mask):
    """Decoder.forward, src/model.py:277-297.  mel_targets (B,80,S) -> mel (B,S,80), gate (B,S), [S x (B,T)]."""
    frames = mel_targets.transpose(1, 2)
    frames = torch.cat([torch.zeros_like(frames[:, :1]), frames[:, :-1]], dim=1)
    st = DecoderState(sd, memory, mask)
    mels, gates, aligns = [], [], []
    for i in range(frames.shape[1]):
        st.x = frames[:, i]
        mel, gate, w = decode_step(sd, st)
        mels.append(mel); gates.append(gate); aligns.append(w)
    return torch.stack(mels, dim=1), torch.stack(gates, dim=1).squeeze(-1), aligns


def decoder_free_running(sd, memory, max_len_cap=None, gate_threshold=0.5, max_decoder_steps=1000):
    """Decoder.inference, src/model.py:299-333 (batch-wide stop rule, discarded first step)."""
    st = DecoderState(sd, memory, None)
    cap = min(max_decoder_steps, max_len_cap) if max_len_cap else max_decoder_steps
    mel0, _, _ = decode_step(sd, st)        # "diagnostic" step: outputs dropped, state kept
    st.x = mel0
    mels, gates, aligns = [], [], []
    steps = 0
    while True:
        mel, gate, w = decode_step(sd, st)
        mels.append(mel); gates.append(gate); aligns.append(w)
        steps += 1
        if steps > 1 and bool(torch.sigmoid(gate).max() > gate_threshold):
            break
        if steps >= cap:
            break
        st.x = mel
    return torch.stack(mels, dim=1), torch.stack(gates, dim=1).squeeze(-1), aligns


# --------------------------------------------------------------------------- postnet
def postnet(sd, mel_bct):
    """PostNet.forward in eval mode, src/model.py:392-407: (B,80,S) -> residual (B,80,S)."""
    n = 0
    while f"postnet.convolutions.{n}.0.weight" in sd:
        n += 1
    x = mel_bct
    for i in range(n):
        p = f"postnet.convolutions.{i}"
        w = sd[p + ".0.weight"]
        x = _bn_eval(F.conv1d(x, w, sd[p + ".0.bias"], padding=(w.shape[2] - 1) // 2), sd, p + ".1")
        if i < n - 1:
            x = torch.tanh(x)
    return x


# --------------------------------------------------------------------------- wrappers
def tacotron2_forward(sd, ids, mel_targets, text_lengths=None, use_postnet=True):
    """Tacotron2.forward, src/model.py:454-480 (without the one-off F10 bias side effect)."""
    memory = encoder(sd, ids)
    if text_lengths is None:
        text_lengths = torch.full((ids.shape[0],), ids.shape[1], dtype=torch.long)
    mask = make_pad_mask(text_lengths, memory.shape[1])
    coarse, gate, aligns = decoder_teacher_forced(sd, memory, mel_targets, mask)
    if use_postnet:
        post = coarse + postnet(sd, coarse.transpose(1, 2)).transpose(1, 2)
    else:
        post = coarse
    return post, coarse, gate, aligns


def tacotron2_inference(sd, ids, max_len_cap=None, gate_threshold=0.5, max_decoder_steps=1000):
    """Tacotron2.inference, src/model.py:482-501 (text_lengths is ignored there, F8)."""
    memory = encoder(sd, ids)
    coarse, gate, aligns = decoder_free_running(sd, memory, max_len_cap, gate_threshold, max_decoder_steps)
    post = coarse + postnet(sd, coarse.transpose(1, 2)).transpose(1, 2)
    return post, coarse, gate, aligns


def per_utterance_lengths(gate_logits, threshold=0.5):
    """Caller-side trimming rule, train.py:71-79, with the steps>1 guard of src/model.py:327.

    Length of row b = (first output index t >= 1 with sigmoid(gate[b,t]) > thr) + 1,
    or the full S if the row never fires.
    """
    fired = torch.sigmoid(gate_logits) > threshold
    fired[:, 0] = False
    S = gate_logits.shape[1]
    idx = torch.where(fired.any(dim=1), fired.float().argmax(dim=1) + 1, torch.full((gate_logits.shape[0],), S))
    return idx.to(torch.int64)
