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
