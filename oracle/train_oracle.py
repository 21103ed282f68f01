"""Differentiable CPU restatement of the reference's TRAIN-mode arithmetic (test infrastructure only).

The eval-mode oracle (tacotron2_oracle.py) is pinned against the reference's outputs; this file adds what changes
in ``model.train()``: BatchNorm uses batch statistics (nn.BatchNorm1d, src/model.py:36, :357) and dropout is active
(F.dropout, src/model.py:170, :262, :270, :402-405).  Dropout masks are explicit inputs (0/1 float tensors, scaled by
1/(1-p) here) so that the CUDA path and this restatement see the same draws; gradients come from torch autograd.
"""
import torch
import torch.nn.functional as F


def _bn_train(x, sd, prefix):
    """x (B,C,T) -> batch-stat normalisation; returns y, batch mean, biased batch var."""
    mean = x.mean(dim=(0, 2))
    var = x.var(dim=(0, 2), unbiased=False)
    y = (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + 1e-5)
    return y * sd[prefix + ".weight"][None, :, None] + sd[prefix + ".bias"][None, :, None], mean, var


def postnet_train(sd, coarse, masks=None, p=0.5):
    """coarse (B,S,80) -> post (B,S,80) = coarse + PostNet(coarse) in train mode; masks[i]: (B*S, c_out_i) or None."""
    n = 0
    while f"postnet.convolutions.{n}.0.weight" in sd:
        n += 1
    B, S, _ = coarse.shape
    x = coarse.transpose(1, 2)
    stats = []
    for i in range(n):
        pre = f"postnet.convolutions.{i}"
        w = sd[pre + ".0.weight"]
        x, mean, var = _bn_train(F.conv1d(x, w, sd[pre + ".0.bias"], padding=(w.shape[2] - 1) // 2), sd, pre + ".1")
        stats.append((mean, var))
        if i < n - 1:
            x = torch.tanh(x)
        if masks is not None:
            m = masks[i].reshape(B, S, -1).transpose(1, 2)
            x = x * m / (1.0 - p)
    return coarse + x.transpose(1, 2), stats


def encoder_train(sd, ids):
    """Encoder.forward in train mode (src/model.py:50-77): embedding, 3 x relu(BN_batch(conv5)), BiLSTM over the full
    padded length with zero initial state.  Returns memory (B,T,512) and the batch statistics per BatchNorm."""
    x = F.embedding(ids, sd["encoder.embedding.weight"]).transpose(1, 2)
    stats = []
    i = 0
    while f"encoder.convolutions.{i}.0.weight" in sd:
        pre = f"encoder.convolutions.{i}"
        w = sd[pre + ".0.weight"]
        x, mean, var = _bn_train(F.conv1d(x, w, sd[pre + ".0.bias"], padding=(w.shape[2] - 1) // 2), sd, pre + ".1")
        x = torch.relu(x)
        stats.append((mean, var))
        i += 1
    x = x.transpose(1, 2)                                   # (B,T,512)
    B, T, _ = x.shape
    outs = []
    for sfx, rev in (("", False), ("_reverse", True)):
        w_ih, w_hh = sd["encoder.lstm.weight_ih_l0" + sfx], sd["encoder.lstm.weight_hh_l0" + sfx]
        b = sd["encoder.lstm.bias_ih_l0" + sfx] + sd["encoder.lstm.bias_hh_l0" + sfx]
        H = w_hh.shape[1]
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        hs = [None] * T
        for t in (range(T - 1, -1, -1) if rev else range(T)):
            z = x[:, t] @ w_ih.t() + h @ w_hh.t() + b
            i_, f_, g_, o_ = z.chunk(4, dim=1)
            c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
            h = torch.sigmoid(o_) * torch.tanh(c)
            hs[t] = h
        outs.append(torch.stack(hs, dim=1))
    return torch.cat(outs, dim=2), stats


def decoder_train(sd, memory, mel_targets, text_lengths=None, masks=None, p_prenet=0.5, p_lstm=0.1):
    """Decoder.forward in train mode, src/model.py:277-297 with _decode_step :246-275: teacher forcing, dropout on both
    prenet layers (p=0.5, :170) and on the CARRIED attention / decoder hidden states (p=0.1, :262, :270).
    masks = dict(p1, p2: (S*B,256) rows ordered [step][b]; a, d: (S,B,1024)) of 0/1 floats, or None (no dropout).
    Returns mel (B,S,80), gate (B,S), alignments (S,B,T)."""
    from oracle import tacotron2_oracle as O
    B, T, _ = memory.shape
    S = mel_targets.shape[2]
    mask = O.make_pad_mask(text_lengths, T) if text_lengths is not None else None
    frames = mel_targets.transpose(1, 2)
    frames = torch.cat([torch.zeros_like(frames[:, :1]), frames[:, :-1]], dim=1)
    st = O.DecoderState(sd, memory, mask)
    d = "decoder."
    mels, gates, aligns = [], [], []
    for i in range(S):
        x = frames[:, i]
        x = torch.relu(x @ sd[d + "prenet.layers.0.weight"].t())
        if masks is not None:
            x = x * masks["p1"][i * B:(i + 1) * B] / (1 - p_prenet)
        x = torch.relu(x @ sd[d + "prenet.layers.1.weight"].t())
        if masks is not None:
            x = x * masks["p2"][i * B:(i + 1) * B] / (1 - p_prenet)
        st.ah, st.ac = O._lstm_cell(torch.cat([x, st.ctx], dim=1), st.ah, st.ac,
                                    sd[d + "attention_lstm.weight_ih"], sd[d + "attention_lstm.weight_hh"],
                                    sd[d + "attention_lstm.bias_ih"], sd[d + "attention_lstm.bias_hh"])
        if masks is not None:
            st.ah = st.ah * masks["a"][i] / (1 - p_lstm)
        e = O.attention_energies(sd, st, st.ah)
        w = torch.softmax(e, dim=1)
        st.prev_attn = w
        st.cum_attn = st.cum_attn + w
        st.ctx = torch.bmm(w[:, None, :], st.memory).squeeze(1)
        st.dh, st.dc = O._lstm_cell(torch.cat([st.ah, st.ctx], dim=1), st.dh, st.dc,
                                    sd[d + "decoder_lstm.weight_ih"], sd[d + "decoder_lstm.weight_hh"],
                                    sd[d + "decoder_lstm.bias_ih"], sd[d + "decoder_lstm.bias_hh"])
        if masks is not None:
            st.dh = st.dh * masks["d"][i] / (1 - p_lstm)
        proj_in = torch.cat([st.dh, st.ctx], dim=1)
        mels.append(torch.addmm(sd[d + "linear_projection.bias"], proj_in, sd[d + "linear_projection.weight"].t()))
        gates.append(torch.addmm(sd[d + "gate_layer.bias"], proj_in, sd[d + "gate_layer.weight"].t()))
        aligns.append(w)
    return torch.stack(mels, dim=1), torch.stack(gates, dim=1).squeeze(-1), torch.stack(aligns, dim=0)
