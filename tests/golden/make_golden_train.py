"""Golden vectors of the reference in TRAIN mode (dropout, BatchNorm batch statistics, autograd gradients), produced by RUNNING
THE UNMODIFIED REFERENCE in the build container.

    python tests/golden/make_golden_train.py        # needs /root/reference; writes tests/golden/train_b3.npz

``src/model.py`` is imported unmodified (``librosa`` stubbed).  ``torch.nn.functional.dropout`` - the only source of randomness
in ``Tacotron2.forward`` (src/model.py:170, :262, :270, :403-405) - is replaced for the duration of the call by a function that
replays EXPLICIT 0/1 masks in call order (x * mask / (1 - p), which is what F.dropout computes for the mask it draws), so that
the oracle and the CUDA path can be held to the same draws.  Call order inside one forward: per decode step prenet layer 0,
prenet layer 1, attention hidden state, decoder hidden state; then the five postnet layers.  A second case runs with dropout
as the identity (train-mode BatchNorm only).

The fixture holds the four outputs, the BatchNorm running statistics after the step and, for each of the 62 parameter
gradients, its L2 norm, a seeded random projection and 256 strided entries (the full gradients are 112 MB).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))

from b200taco import synth  # noqa: E402

REF = "/root/reference"
B, T, S = 3, 11, 7
LENGTHS = (11, 8, 5)
P_PRENET, P_LSTM, P_POST = 0.5, 0.1, 0.5


def train_case(seed=77):
    """Seeded inputs: ids, mel targets, text lengths, dropout masks in the oracle's layout (oracle/train_oracle.py), and the
    weights of the scalar loss  sum(w_post * post) + sum(w_coarse * coarse) + sum(w_gate * gate) + sum(w_align * align)."""
    g = torch.Generator().manual_seed(seed)
    ids = synth.phoneme_ids(B, T, seed=seed)
    mel = torch.randn(B, 80, S, generator=g) * 2 - 5
    lengths = torch.tensor(LENGTHS)
    r = lambda *shape, p: (torch.rand(*shape, generator=g) >= p).float()
    dec = {"p1": r(S * B, 256, p=P_PRENET), "p2": r(S * B, 256, p=P_PRENET), "a": r(S, B, 1024, p=P_LSTM), "d": r(S, B, 1024, p=P_LSTM)}
    post = [r(B * S, c, p=P_POST) for c in (512, 512, 512, 512, 80)]
    w = dict(post=torch.randn(B, S, 80, generator=g), coarse=torch.randn(B, S, 80, generator=g), gate=torch.randn(B, S, generator=g),
             align=torch.randn(S, B, T, generator=g))
    return ids, mel, lengths, dec, post, w


def name_seed(name):
    """Deterministic across processes (hash() is salted): seed of the projection vector of a parameter."""
    return sum((i + 1) * ord(c) for i, c in enumerate(name)) % (2 ** 31)


def summarise(named_grads):
    out = {}
    for name, grad in named_grads.items():
        g = grad.detach().double().flatten()
        v = torch.randn(g.numel(), generator=torch.Generator().manual_seed(name_seed(name)), dtype=torch.float64)
        stride = max(1, g.numel() // 256)
        sample = g[::stride][:256]
        out[name] = np.concatenate([[float(g.norm()), float((g * v).sum()), float(g.abs().max())],
                                    torch.nn.functional.pad(sample, (0, 256 - sample.numel())).numpy()])
    return out


def main():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    sys.path.insert(0, REF)
    from src.model import Tacotron2  # the unmodified reference
    import torch.nn.functional as F
    ids, mel, lengths, dec, post, w = train_case()
    out = {}
    for case, use_masks in (("drop", True), ("nodrop", False)):
        torch.manual_seed(synth.WEIGHT_SEED)
        m = Tacotron2()
        m._projection_bias_initialized = True
        m.train()
        queue = []
        if use_masks:
            for i in range(S):
                queue += [(dec["p1"][i * B:(i + 1) * B], P_PRENET), (dec["p2"][i * B:(i + 1) * B], P_PRENET), (dec["a"][i], P_LSTM),
                          (dec["d"][i], P_LSTM)]
            queue += [(mk.reshape(B, S, -1).transpose(1, 2), P_POST) for mk in post]
        calls = {"n": 0}
        real = F.dropout

        def replay(x, p=0.5, training=True, inplace=False):
            assert training
            calls["n"] += 1
            if not use_masks:
                return x
            mask, want_p = queue.pop(0)
            assert mask.shape == x.shape and abs(want_p - p) < 1e-12, (mask.shape, x.shape, p)
            return x * mask / (1.0 - p)

        F.dropout = replay
        try:
            mel_post, mel_coarse, gate, aligns = m(ids, mel, lengths)
        finally:
            F.dropout = real
        assert calls["n"] == 4 * S + 5 and not queue
        al = torch.stack(aligns, 0)
        loss = (mel_post * w["post"]).sum() + (mel_coarse * w["coarse"]).sum() + (gate * w["gate"]).sum() + (al * w["align"]).sum()
        loss.backward()
        out[f"{case}_mel_post"] = mel_post.detach().numpy()
        out[f"{case}_mel_coarse"] = mel_coarse.detach().numpy()
        out[f"{case}_gate"] = gate.detach().numpy()
        out[f"{case}_align"] = al.detach().numpy()
        for k, v in summarise({k: p.grad for k, p in m.named_parameters()}).items():
            out[f"{case}_grad/{k}"] = v
        for k, v in m.state_dict().items():
            if "running" in k or "num_batches" in k:
                out[f"{case}_buf/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "train_b3.npz"), **out)
    print(len(out), "arrays;", os.path.getsize(os.path.join(HERE, "train_b3.npz")), "bytes")


if __name__ == "__main__":
    main()
