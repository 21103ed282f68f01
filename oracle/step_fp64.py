"""Independent float64 numpy restatement of ONE decoder step -- TEST INFRASTRUCTURE.

Exists so that a bug in ``tacotron2_oracle.py`` and a bug in the CUDA kernels can be told
apart: it shares no code with either.  Follows the step formulas of SURVEY.md section
8(a), i.e. Decoder._decode_step (/root/reference/src/model.py:246-275) with
LocationSensitiveAttention.forward (:117-148), eval mode.  Pure loops over the location
convolution taps; only meant for small T.
"""
import numpy as np


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _cell(W_ih, W_hh, b_ih, b_hh, x, h, c):
    z = W_ih @ x + b_ih + W_hh @ h + b_hh
    H = h.shape[0]
    i, f, g, o = z[:H], z[H:2 * H], z[2 * H:3 * H], z[3 * H:]
    c2 = _sig(f) * c + _sig(i) * np.tanh(g)
    return _sig(o) * np.tanh(c2), c2


def step(w, s, memory, pm, mask=None):
    """One step for ONE utterance.

    w: dict of float64 numpy weights keyed like the reference state_dict.
    s: dict with x(80) ah ac dh dc(1024) ctx(512) prev cum(T).  Returns (mel, gate, attn, new_s).
    memory (T,512), pm (T,128) = memory @ W_mem^T.
    """
    d, a = "decoder.", "decoder.attention."
    p = np.maximum(w[d + "prenet.layers.0.weight"] @ s["x"], 0.0)
    p = np.maximum(w[d + "prenet.layers.1.weight"] @ p, 0.0)
    ah, ac = _cell(w[d + "attention_lstm.weight_ih"], w[d + "attention_lstm.weight_hh"],
                   w[d + "attention_lstm.bias_ih"], w[d + "attention_lstm.bias_hh"],
                   np.concatenate([p, s["ctx"]]), s["ah"], s["ac"])
    T = memory.shape[0]
    Wloc = w[a + "location_conv.weight"]          # (32, 2, 31)
    K = Wloc.shape[2]
    half = (K - 1) // 2
    chans = (s["prev"], s["cum"])                 # channel order: prev, cum (src/model.py:125)
    loc = np.zeros((Wloc.shape[0], T))
    for f in range(Wloc.shape[0]):
        for c in range(2):
            for k in range(K):
                for t in range(T):
                    u = t + k - half
                    if 0 <= u < T:
                        loc[f, t] += Wloc[f, c, k] * chans[c][u]
    pq = w[a + "query_layer.weight"] @ ah
    e = np.empty(T)
    for t in range(T):
        u = np.tanh(pq + pm[t] + w[a + "location_dense.weight"] @ loc[:, t])
        e[t] = (w[a + "v.weight"][0] @ u + w[a + "v.bias"][0]) * w[a + "energy_scale"]
    if mask is not None:
        e = np.where(mask, -1e9, e)
    ex = np.exp(e - e.max())
    attn = ex / ex.sum()
    ctx = attn @ memory
    dh, dc = _cell(w[d + "decoder_lstm.weight_ih"], w[d + "decoder_lstm.weight_hh"],
                   w[d + "decoder_lstm.bias_ih"], w[d + "decoder_lstm.bias_hh"],
                   np.concatenate([ah, ctx]), s["dh"], s["dc"])
    pin = np.concatenate([dh, ctx])
    mel = w[d + "linear_projection.weight"] @ pin + w[d + "linear_projection.bias"]
    gate = w[d + "gate_layer.weight"] @ pin + w[d + "gate_layer.bias"]
    new = dict(x=mel, ah=ah, ac=ac, dh=dh, dc=dc, ctx=ctx, prev=attn, cum=s["cum"] + attn)
    return mel, gate, attn, new
