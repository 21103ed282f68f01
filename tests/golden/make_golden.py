"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE in the build container.

    python tests/golden/make_golden.py

Needs /root/reference (read-only mount; it does not exist on the GPU box, which is why the
outputs are committed).  The reference is imported unmodified; only ``librosa`` (imported
by src/audio.py at package import, never called on this path -- SURVEY.md F1) is stubbed.

Every case is fully determined by seeds, so the fixtures hold only the reference's OUTPUTS
plus a fingerprint of the weights; the inputs are rebuilt by ``b200taco.synth`` in the
tests.  Weight sets: "default" = torch.manual_seed(1234) default init,
"stress" = synth.stress_init_, "planted" = synth.plant_stop_counter_; "+bf16" = the four
decoder LSTM matrices rounded to bf16-representable values (synth.round_to_bf16_).
"""
import contextlib
import io
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


def load_reference():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.model import Tacotron2  # the unmodified reference
    return Tacotron2


def ref_model(Tacotron2, surgery=(), bf16=False, memory_for_center=None):
    torch.manual_seed(synth.WEIGHT_SEED)
    m = Tacotron2().eval()
    m._projection_bias_initialized = True          # F10 side effect is tested separately
    sd = m.state_dict()
    with torch.no_grad():
        if "stress" in surgery:
            synth.stress_init_(sd)
        if "planted" in surgery:
            synth.plant_stop_counter_(sd, **(surgery["planted"] if isinstance(surgery, dict) else {}))
        if bf16:
            synth.round_to_bf16_(sd)
    return m


def fingerprint(sd):
    """Per-tensor (sum, abs-sum) in float64: lets a test prove it rebuilt the same weights."""
    keys = sorted(sd.keys())
    return np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys]), keys


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def pack(out):
    post, coarse, gate, aligns = out
    return dict(mel_post=post.numpy(), mel_coarse=coarse.numpy(), gate=gate.numpy(),
                align=torch.stack(aligns, dim=0).numpy())


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.set_grad_enabled(False)
    Tacotron2 = load_reference()
    cases = {}

    # ---- C1 (BASELINE config 1): teacher-forced forward, B=4, lengths [37,29,21,12], S=120
    ids, lengths, mel = synth.c1_batch()
    for name, kw in (("c1_default", {}), ("c1_default_bf16", dict(bf16=True)),
                     ("c1_stress_bf16", dict(surgery=("stress",), bf16=True))):
        m = ref_model(Tacotron2, **kw)
        cases[name] = pack(m(ids, mel, lengths))
    m = ref_model(Tacotron2)
    cases["c1_default_nopostnet"] = pack(m(ids, mel, lengths, use_postnet=False))
    cases["c1_default_nolengths"] = pack(m(ids, mel, None))

    # ---- F10: first forward overwrites linear_projection.bias with the target mean
    torch.manual_seed(synth.WEIGHT_SEED)
    m = Tacotron2().eval()
    out = pack(m(ids, mel, lengths))
    out["proj_bias_after"] = m.decoder.linear_projection.bias.numpy().copy()
    cases["c1_first_call_bias_init"] = out

    # ---- encoder only (B=3, T=19) and memory for a decoder-only test
    ids_e = synth.phoneme_ids(3, 19, seed=5)
    m = ref_model(Tacotron2)
    cases["encoder_b3_t19"] = dict(memory=m.encoder(ids_e).numpy())

    # ---- free-running, default init (gate never fires -> cap), B=2 T=33 cap 40; B=1 T=50 cap 25
    m = ref_model(Tacotron2, bf16=True)
    cases["infer_default_b2_t33_cap40"] = pack(quiet(m.inference, synth.phoneme_ids(2, 33, seed=2), max_len_cap=40))
    cases["infer_default_b1_t50_cap25"] = pack(quiet(m.inference, synth.phoneme_ids(1, 50, seed=3), max_len_cap=25))
    m = ref_model(Tacotron2)
    cases["infer_default_fp32_b2_t33_cap40"] = pack(quiet(m.inference, synth.phoneme_ids(2, 33, seed=2), max_len_cap=40))

    # ---- free-running, stress + planted stop counter (batch-wide stop, F7), B=5 T=41
    ids_p = synth.phoneme_ids(5, 41, seed=7)
    torch.manual_seed(synth.WEIGHT_SEED)
    m = Tacotron2().eval()
    m._projection_bias_initialized = True
    sd = m.state_dict()
    synth.stress_init_(sd)
    center = m.encoder(ids_p).mean((0, 1))
    synth.plant_stop_counter_(sd, rate=6e-3, jitter=2e-2, ctx_center=center)
    synth.round_to_bf16_(sd)
    out = pack(quiet(m.inference, ids_p))
    cases["infer_planted_b5_t41"] = out
    # solo runs: the per-utterance stop rule's oracle (SURVEY.md 8c "stop semantics")
    solo_len, solo_mel = [], []
    for b in range(ids_p.shape[0]):
        o = quiet(m.inference, ids_p[b:b + 1])
        solo_len.append(o[1].shape[1])
        solo_mel.append(o[1][0].numpy())
    cases["infer_planted_b5_t41"]["solo_lengths"] = np.array(solo_len, dtype=np.int64)
    cases["infer_planted_b5_t41"]["ctx_center"] = center.numpy()
    S = max(solo_len)
    sm = np.zeros((len(solo_len), S, 80), np.float32)
    for b, x in enumerate(solo_mel):
        sm[b, :len(x)] = x
    cases["infer_planted_b5_t41"]["solo_mel_coarse"] = sm
    # threshold override and cap interplay
    cases["infer_planted_b5_t41_thr0.3"] = pack(quiet(m.inference, ids_p, gate_threshold=0.3))
    cases["infer_planted_b5_t41_cap9"] = pack(quiet(m.inference, ids_p, max_len_cap=9))

    # ---- postnet only
    g = torch.Generator().manual_seed(11)
    coarse = torch.randn(2, 80, 37, generator=g)
    m = ref_model(Tacotron2)
    cases["postnet_b2_s37"] = dict(residual=m.postnet(coarse).numpy())

    # ---- weight fingerprints
    m = ref_model(Tacotron2)
    fp, keys = fingerprint(m.state_dict())
    np.savez_compressed(os.path.join(HERE, "weights_fingerprint.npz"), fingerprint=fp, keys=np.array(keys))

    for name, arrays in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})
    print("torch", torch.__version__)


if __name__ == "__main__":
    main()
