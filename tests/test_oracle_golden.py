"""The CPU oracle is held to outputs of the UNMODIFIED reference (tests/golden/*.npz).

These fixtures were produced by tests/golden/make_golden.py importing
/root/reference/src/model.py in the build container; this file never touches the reference.
Tolerance: both sides are fp32 torch-CPU arithmetic with different op grouping.
"""
import numpy as np
import pytest
import torch

from oracle import tacotron2_oracle as O
from oracle import step_fp64
from b200taco import synth
from b200taco.model import Tacotron2

import util

TOL = 2e-6      # fp32 restatement vs fp32 reference on |mel| <= ~10
TOL_ALIGN = 2e-7


def test_weight_fingerprint_matches_reference_init():
    """b200taco.model.Tacotron2 draws the same default initialisation as the reference."""
    import os
    with np.load(os.path.join(util.GOLDEN, "weights_fingerprint.npz")) as z:
        fp, keys = z["fingerprint"], [str(k) for k in z["keys"]]
    sd = util.weights()
    assert sorted(sd.keys()) == keys
    mine = np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])
    np.testing.assert_allclose(mine, fp, rtol=0, atol=0)


def _check(out, gold, tol=TOL, tol_align=TOL_ALIGN):
    post, coarse, gate, aligns = out
    assert coarse.shape == gold["mel_coarse"].shape
    assert util.maxdiff(coarse, gold["mel_coarse"]) <= tol
    assert util.maxdiff(post, gold["mel_post"]) <= tol * 5
    assert util.maxdiff(gate, gold["gate"]) <= tol
    assert util.maxdiff(util.stack(aligns), gold["align"]) <= tol_align


@pytest.mark.parametrize("name,kind,bf16", [("c1_default", "default", False),
                                            ("c1_default_bf16", "default", True),
                                            ("c1_stress_bf16", "stress", True)])
def test_c1_teacher_forced(name, kind, bf16):
    ids, lengths, mel = synth.c1_batch()
    with torch.no_grad():
        out = O.tacotron2_forward(util.weights(kind, bf16), ids, mel, lengths)
    _check(out, util.golden(name), tol=1e-5 if kind == "stress" else TOL, tol_align=2e-6 if kind == "stress" else TOL_ALIGN)


def test_c1_variants():
    ids, lengths, mel = synth.c1_batch()
    sd = util.weights()
    with torch.no_grad():
        _check(O.tacotron2_forward(sd, ids, mel, lengths, use_postnet=False), util.golden("c1_default_nopostnet"))
        _check(O.tacotron2_forward(sd, ids, mel, None), util.golden("c1_default_nolengths"))


def test_c1_first_call_bias_init():
    """F10: the first forward replaces linear_projection.bias by the per-channel target mean."""
    ids, lengths, mel = synth.c1_batch()
    gold = util.golden("c1_first_call_bias_init")
    sd = util.weights()
    sd["decoder.linear_projection.bias"] = mel.mean(dim=(0, 2))
    assert util.maxdiff(sd["decoder.linear_projection.bias"], gold["proj_bias_after"]) == 0.0
    with torch.no_grad():
        _check(O.tacotron2_forward(sd, ids, mel, lengths), gold, tol=1e-5)


def test_encoder():
    ids = synth.phoneme_ids(3, 19, seed=5)
    with torch.no_grad():
        mem = O.encoder(util.weights(), ids)
    assert util.maxdiff(mem, util.golden("encoder_b3_t19")["memory"]) <= 1e-6


def test_postnet():
    g = torch.Generator().manual_seed(11)
    coarse = torch.randn(2, 80, 37, generator=g)
    with torch.no_grad():
        r = O.postnet(util.weights(), coarse)
    assert util.maxdiff(r, util.golden("postnet_b2_s37")["residual"]) <= 5e-6


@pytest.mark.parametrize("name,B,T,seed,cap,bf16", [("infer_default_b2_t33_cap40", 2, 33, 2, 40, True),
                                                    ("infer_default_b1_t50_cap25", 1, 50, 3, 25, True),
                                                    ("infer_default_fp32_b2_t33_cap40", 2, 33, 2, 40, False)])
def test_free_running_default(name, B, T, seed, cap, bf16):
    with torch.no_grad():
        out = O.tacotron2_inference(util.weights("default", bf16), synth.phoneme_ids(B, T, seed=seed), max_len_cap=cap)
    _check(out, util.golden(name))


def test_free_running_planted_stop():
    """Batch-wide stop rule (F7), threshold override, cap; and the per-utterance rule vs solo runs."""
    gold = util.golden("infer_planted_b5_t41")
    ids = synth.phoneme_ids(5, 41, seed=7)
    sd = util.weights("planted", True, ctx_center=gold["ctx_center"])
    with torch.no_grad():
        out = O.tacotron2_inference(sd, ids)
        assert out[1].shape[1] == gold["mel_coarse"].shape[1] == int(gold["solo_lengths"].min())
        _check(out, gold, tol=2e-5, tol_align=2e-6)
        _check(O.tacotron2_inference(sd, ids, gate_threshold=0.3), util.golden("infer_planted_b5_t41_thr0.3"),
               tol=2e-5, tol_align=2e-6)
        _check(O.tacotron2_inference(sd, ids, max_len_cap=9), util.golden("infer_planted_b5_t41_cap9"),
               tol=2e-5, tol_align=2e-6)
        # per-utterance rule: run the batch with the stop disabled, trim post-hoc, compare with solo runs
        long = O.tacotron2_inference(sd, ids, gate_threshold=2.0, max_len_cap=int(gold["solo_lengths"].max()) + 5)
    lens = O.per_utterance_lengths(long[2], 0.5)
    assert lens.tolist() == gold["solo_lengths"].tolist()
    for b, n in enumerate(lens.tolist()):
        assert util.maxdiff(long[1][b, :n], gold["solo_mel_coarse"][b, :n]) <= 2e-5


def test_fp64_step_agrees_with_oracle_step():
    """Independent numpy float64 restatement of one decode step vs the torch oracle (float64)."""
    sd = util.weights("stress", dtype=torch.float64)
    B, T = 2, 23
    g = torch.Generator().manual_seed(21)
    mem = torch.randn(B, T, 512, generator=g, dtype=torch.float64) * 0.3
    lengths = torch.tensor([23, 17])
    mask = O.make_pad_mask(lengths, T)
    st = O.DecoderState(sd, mem, mask)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64) * 0.5
    st.x, st.ah, st.ac, st.dh, st.dc, st.ctx = rnd(B, 80), rnd(B, 1024), rnd(B, 1024), rnd(B, 1024), rnd(B, 1024), rnd(B, 512)
    st.prev_attn = torch.softmax(rnd(B, T), 1)
    st.cum_attn = torch.softmax(rnd(B, T), 1) * 3
    w = {k: v.numpy() for k, v in sd.items()}
    for b in range(B):
        s = dict(x=st.x[b].numpy(), ah=st.ah[b].numpy(), ac=st.ac[b].numpy(), dh=st.dh[b].numpy(),
                 dc=st.dc[b].numpy(), ctx=st.ctx[b].numpy(), prev=st.prev_attn[b].numpy(), cum=st.cum_attn[b].numpy())
        mel, gate, attn, new = step_fp64.step(w, s, mem[b].numpy(), st.processed_memory[b].numpy(), mask[b].numpy())
        if b == 0:
            import copy
            st2 = copy.copy(st)
            melo, gateo, wo = O.decode_step(sd, st2)
        assert np.abs(mel - melo[b].numpy()).max() < 1e-12
        assert np.abs(gate - gateo[b].numpy()).max() < 1e-12
        assert np.abs(attn - wo[b].numpy()).max() < 1e-13
        assert np.abs(new["dc"] - st2.dc[b].numpy()).max() < 1e-12
        assert np.abs(new["cum"] - st2.cum_attn[b].numpy()).max() < 1e-13
