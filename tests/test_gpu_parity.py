"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors produced by
the unmodified reference, and the CPU oracle on the same seeded inputs.

Tolerances (fp32 state and accumulation everywhere; only summation order and the
exp/tanh implementations differ from the reference's CPU kernels):
  mel / gate   : 2e-5 absolute on |mel| <= ~10     (measured ~2e-6)
  alignments   : 2e-6 absolute on weights in [0,1]
  stop indices / decoded lengths : exact
"bf16" = decoder LSTM matrices stored as bf16 in the packed cache; the reference side then
uses the same bf16-representable values, so the tolerance is unchanged.
"""
import os

import pytest
import torch

import util
from b200taco import synth
from b200taco._cabi import B200TacoError
from b200taco.model import Tacotron2
from oracle import tacotron2_oracle as O

pytestmark = pytest.mark.gpu

TOL_MEL = 2e-5
TOL_ALIGN = 2e-6
# shape sweeps run on the STRESS weights (non-LSTM decoder matrices x 2.5, energy_scale 4.0) at the same
# tolerances (largest errors measured over all shapes: mel 1.2e-7, gate 5e-6, postnet 2e-7, alignment 4e-7)
SWEEP_TOL = TOL_MEL
SWEEP_TOL_ALIGN = TOL_ALIGN


def gpu_model(sd, weight_dtype="bf16"):
    m = Tacotron2().eval()
    m.load_state_dict(sd)
    m._projection_bias_initialized = True
    m.lstm_weight_dtype = weight_dtype
    if os.environ.get("TACO_TEST_DENSE_PATH"):           # experiment knob: run the whole file on another dense-layer path
        m.dense_path = os.environ["TACO_TEST_DENSE_PATH"]
    return m.to("cuda:0")


def check(out, gold, tol=TOL_MEL, tol_align=TOL_ALIGN, post_factor=5):
    post, coarse, gate, aligns = out
    assert isinstance(aligns, list) and len(aligns) == coarse.shape[1]          # F9
    assert tuple(coarse.shape) == tuple(gold["mel_coarse"].shape)
    assert tuple(post.shape) == tuple(gold["mel_post"].shape) and tuple(gate.shape) == tuple(gold["gate"].shape)
    e = dict(mel=util.maxdiff(coarse.cpu(), gold["mel_coarse"]), gate=util.maxdiff(gate.cpu(), gold["gate"]),
             post=util.maxdiff(post.cpu(), gold["mel_post"]),
             align=util.maxdiff(util.stack(aligns).cpu(), gold["align"]))
    print(e)
    assert e["mel"] <= tol and e["gate"] <= tol and e["post"] <= tol * post_factor and e["align"] <= tol_align, e


# ------------------------------------------------------------------ golden vectors (reference outputs)
def test_encoder_golden():
    m = gpu_model(util.weights())
    ids = synth.phoneme_ids(3, 19, seed=5).cuda()
    mem = m._get_engine().encoder(ids)
    assert util.maxdiff(mem.cpu(), util.golden("encoder_b3_t19")["memory"]) <= 2e-6


@pytest.mark.parametrize("n_utts,lo,hi", [(5, 3, 30), (19, 12, 48), (64, 100, 150)])
def test_ragged_encoder_equals_exact_length_solo_runs(n_utts, lo, hi):
    """b200taco_encoder_fwd_ragged: utterances of different lengths padded into one batch; every row must be BIT FOR BIT what the
    encoder returns for that utterance alone at its exact length (the reference's inference encodes one unpadded utterance at a
    time), and zero behind its length - whatever the padded ids are.  The solo run itself is held to the oracle."""
    sd = util.weights()
    eng = gpu_model(sd)._get_engine()
    g = torch.Generator().manual_seed(n_utts)
    lens = torch.randint(lo, hi + 1, (n_utts,), generator=g)
    lens[0] = hi
    T = int(lens.max())
    utts = [synth.phoneme_ids(1, int(L), seed=700 + i)[0] for i, L in enumerate(lens)]
    ids = torch.randint(0, 70, (n_utts, T), generator=g)                      # garbage behind every utterance's length
    for i, u in enumerate(utts):
        ids[i, :len(u)] = u
    mem = eng.encoder(ids.cuda(), lengths=lens.cuda())
    for i, u in enumerate(utts):
        solo = eng.encoder(u[None].cuda())
        L = len(u)
        assert torch.equal(mem[i, :L], solo[0]), (i, L, util.maxdiff(mem[i, :L].cpu(), solo[0].cpu()))
        assert not mem[i, L:].any()
    ref = O.encoder(sd, utts[1][None])
    assert util.maxdiff(mem[1, :len(utts[1])].cpu(), ref[0]) <= 2e-6


def test_postnet_golden():
    m = gpu_model(util.weights())
    g = torch.Generator().manual_seed(11)
    coarse = torch.randn(2, 80, 37, generator=g)
    x = coarse.transpose(1, 2).contiguous().cuda()
    post = m._get_engine().postnet(x)
    resid = (post - x).transpose(1, 2).cpu()
    assert util.maxdiff(resid, util.golden("postnet_b2_s37")["residual"]) <= 1e-5


@pytest.mark.parametrize("dense_path", ["auto", "simt", "tc"])
def test_dense_paths_golden_and_large(dense_path):
    """Both dense-layer paths (tcgen05 3xTF32 / fp32 SIMT) against the reference fixtures, then at BASELINE
    C2 size against the CPU oracle (encoder B=32 x T=150; postnet 32 x 600 frames)."""
    sd = util.weights()
    m = Tacotron2().eval()
    m.load_state_dict(sd)
    m.dense_path = dense_path
    m = m.to("cuda:0")
    eng = m._get_engine()
    ids = synth.phoneme_ids(3, 19, seed=5).cuda()
    # tensor-core accumulation truncates: ~1e-6 relative instead of fp32 round-to-nearest (include/b200taco.h)
    tol_mem = 4e-6 if dense_path == "tc" else 2e-6
    assert util.maxdiff(eng.encoder(ids).cpu(), util.golden("encoder_b3_t19")["memory"]) <= tol_mem
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 80, 37, generator=g).transpose(1, 2).contiguous().cuda()
    resid = (eng.postnet(x) - x).transpose(1, 2).cpu()
    assert util.maxdiff(resid, util.golden("postnet_b2_s37")["residual"]) <= 1e-5
    ids = synth.phoneme_ids(32, 150, seed=9)
    assert util.maxdiff(eng.encoder(ids.cuda()).cpu(), O.encoder(sd, ids)) <= tol_mem
    g = torch.Generator().manual_seed(12)
    mel = torch.randn(32, 600, 80, generator=g) * 2 - 5
    ref = mel + O.postnet(sd, mel.transpose(1, 2)).transpose(1, 2)
    assert util.maxdiff(eng.postnet(mel.cuda()).cpu(), ref) <= 2e-5          # on |mel| up to ~12


@pytest.mark.parametrize("name,kind,bf16,wdt", [("c1_default", "default", False, "fp32"),
                                                ("c1_default_bf16", "default", True, "bf16"),
                                                ("c1_stress_bf16", "stress", True, "bf16"),
                                                ("c1_stress_bf16", "stress", True, "fp32")])
def test_c1_teacher_forced_golden(name, kind, bf16, wdt):
    """BASELINE config 1: B=4, lengths [37,29,21,12], 120 target frames, masked attention."""
    ids, lengths, mel = synth.c1_batch()
    m = gpu_model(util.weights(kind, bf16), wdt)
    check(m(ids.cuda(), mel.cuda(), lengths.cuda()), util.golden(name))


def test_c1_variants_golden():
    ids, lengths, mel = synth.c1_batch()
    m = gpu_model(util.weights(), "fp32")
    out = m(ids.cuda(), mel.cuda(), lengths.cuda(), use_postnet=False)
    assert out[0] is out[1]                                       # src/model.py:479-480
    check(out, util.golden("c1_default_nopostnet"))
    check(m(ids.cuda(), mel.cuda(), None), util.golden("c1_default_nolengths"))


def test_c1_first_call_bias_init_golden():
    """F10: first forward overwrites linear_projection.bias with the target mean, once."""
    ids, lengths, mel = synth.c1_batch()
    m = gpu_model(util.weights(), "fp32")
    m._projection_bias_initialized = False
    gold = util.golden("c1_first_call_bias_init")
    check(m(ids.cuda(), mel.cuda(), lengths.cuda()), gold)
    assert m._projection_bias_initialized
    assert util.maxdiff(m.decoder.linear_projection.bias.cpu(), gold["proj_bias_after"]) <= 1e-6
    before = m.decoder.linear_projection.bias.clone()
    m(ids.cuda(), (mel + 1).cuda(), lengths.cuda())
    assert torch.equal(before, m.decoder.linear_projection.bias)


@pytest.mark.parametrize("name,B,T,seed,cap,bf16", [("infer_default_b2_t33_cap40", 2, 33, 2, 40, True),
                                                    ("infer_default_b1_t50_cap25", 1, 50, 3, 25, True),
                                                    ("infer_default_fp32_b2_t33_cap40", 2, 33, 2, 40, False)])
def test_free_running_default_golden(name, B, T, seed, cap, bf16, capsys):
    m = gpu_model(util.weights("default", bf16), "bf16" if bf16 else "fp32")
    out = m.inference(synth.phoneme_ids(B, T, seed=seed).cuda(), max_len_cap=cap)
    assert "Reached decode cap" in capsys.readouterr().out        # src/model.py:330
    check(out, util.golden(name))


def test_free_running_planted_stop_golden():
    """Stop indices exact: batch-wide rule (F7), threshold override, cap, and the per-utterance rule vs solo runs."""
    gold = util.golden("infer_planted_b5_t41")
    ids = synth.phoneme_ids(5, 41, seed=7).cuda()
    m = gpu_model(util.weights("planted", True, ctx_center=gold["ctx_center"]))
    check(m.inference(ids), gold)
    check(m.inference(ids, text_lengths=torch.full((5,), 20).cuda()), gold)          # F8: lengths ignored
    check(m.inference(ids, gate_threshold=0.3), util.golden("infer_planted_b5_t41_thr0.3"))
    check(m.inference(ids, max_len_cap=9), util.golden("infer_planted_b5_t41_cap9"))
    post, coarse, gate, aligns, lengths = m.inference_per_utterance(ids)
    assert lengths.tolist() == gold["solo_lengths"].tolist()
    assert coarse.shape[1] == int(gold["solo_lengths"].max())
    for b, n in enumerate(lengths.tolist()):
        assert util.maxdiff(coarse[b, :n].cpu(), gold["solo_mel_coarse"][b, :n]) <= TOL_MEL


@pytest.mark.parametrize("B", [1, 2])
def test_resident_small_batch_stop_rules_vs_oracle(B):
    """B <= 2 free-running decodes take the weight-resident, barrier-free kernel (decoder_res.cu): its stop rules - batch wide
    (F7, src/model.py:327), threshold override, frame cap, per utterance (train.py:71-79) - against the oracle, indices exact."""
    gold = util.golden("infer_planted_b5_t41")
    sd = util.weights("planted", True, ctx_center=gold["ctx_center"])
    ids = synth.phoneme_ids(5, 41, seed=7)[:B].contiguous()
    m = gpu_model(sd)
    for kw in ({}, {"gate_threshold": 0.3}, {"max_len_cap": 9}):
        with torch.no_grad():
            ref = O.tacotron2_inference(sd, ids, **kw)
        check(m.inference(ids.cuda(), **kw), dict(mel_post=ref[0], mel_coarse=ref[1], gate=ref[2], align=util.stack(ref[3])))
    post, coarse, gate, aligns, lengths = m.inference_per_utterance(ids.cuda())
    assert lengths.tolist() == gold["solo_lengths"][:B].tolist()
    assert coarse.shape[1] == int(gold["solo_lengths"][:B].max())
    for b, n in enumerate(lengths.tolist()):
        assert util.maxdiff(coarse[b, :n].cpu(), gold["solo_mel_coarse"][b, :n]) <= TOL_MEL


# ------------------------------------------------------------------ oracle on the same seeded inputs
@pytest.mark.parametrize("B,T", [(1, 1), (1, 200), (1, 1024), (2, 17), (2, 200), (2, 600), (3, 257), (7, 64), (8, 150), (13, 40), (16, 150),
                                 (32, 150), (33, 90), (64, 150), (64, 7)])
def test_shape_sweep_free_running_vs_oracle(B, T):
    """Every batch-tile specialisation (1,2,4,...,64) and slice count, stress weights (F12), 12 steps."""
    sd = util.weights("stress", True)
    ids = synth.phoneme_ids(B, T, seed=100 + B)
    m = gpu_model(sd)
    out = m.inference(ids.cuda(), max_len_cap=12)
    with torch.no_grad():
        ref = O.tacotron2_inference(sd, ids, max_len_cap=12)
    gold = dict(mel_post=ref[0], mel_coarse=ref[1], gate=ref[2], align=util.stack(ref[3]))
    check(out, gold, tol=SWEEP_TOL, tol_align=SWEEP_TOL_ALIGN)


@pytest.mark.parametrize("B,T,S", [(1, 23, 9), (5, 77, 30), (32, 150, 20), (64, 33, 6)])
def test_shape_sweep_teacher_forced_vs_oracle(B, T, S):
    sd = util.weights("stress", True)
    ids = synth.phoneme_ids(B, T, seed=B)
    g = torch.Generator().manual_seed(B)
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    lengths[0] = T
    mel = synth.mel_targets(B, S, seed=B)
    m = gpu_model(sd)
    out = m(ids.cuda(), mel.cuda(), lengths.cuda())
    with torch.no_grad():
        ref = O.tacotron2_forward(sd, ids, mel, lengths)
    gold = dict(mel_post=ref[0], mel_coarse=ref[1], gate=ref[2], align=util.stack(ref[3]))
    check(out, gold, tol=SWEEP_TOL, tol_align=SWEEP_TOL_ALIGN)
    al = util.stack(out[3]).cpu()
    for b in range(B):
        assert float(al[:, b, int(lengths[b]):].abs().max() if lengths[b] < T else 0.0) == 0.0   # masked positions


def test_c2_full_size_stop_lengths_exact():
    """BASELINE config 2 at full size: B=32, T=150, planted stop counter, cap 1000.

    Per-utterance decoded lengths must equal the oracle's exactly; mel within tolerance up to
    each utterance's own length.  Margins of the gate logit around the threshold are reported.
    """
    B, T = 32, 150
    ids = synth.phoneme_ids(B, T, seed=0)
    sd0 = util.weights("default")
    with torch.no_grad():
        center = O.encoder(sd0, ids).mean((0, 1))
    sd = util.weights("default")
    with torch.no_grad():
        synth.plant_stop_counter_(sd, rate=2e-3, jitter=1.2e-2, ctx_center=center)
        synth.round_to_bf16_(sd)
        ref = O.tacotron2_inference(sd, ids, gate_threshold=2.0, max_len_cap=400)
    ref_len = O.per_utterance_lengths(ref[2], 0.5)
    assert int(ref_len.min()) >= 20 and len(set(ref_len.tolist())) > 4, ref_len.tolist()
    m = gpu_model(sd)
    post, coarse, gate, aligns, lengths = m.inference_per_utterance(ids.cuda(), max_len_cap=400)
    idx = (ref_len - 1).clamp(max=ref[2].shape[1] - 1)
    margin = ref[2].gather(1, idx[:, None]).abs().min()
    print("ref lengths", ref_len.tolist(), "min |logit| at the firing step", float(margin))
    assert lengths.tolist() == ref_len.tolist()
    assert coarse.shape[1] == int(ref_len.max())
    for b, n in enumerate(ref_len.tolist()):
        assert util.maxdiff(coarse[b, :n].cpu(), ref[1][b, :n]) <= 5e-5
    # batch-wide rule on the same weights: S = min over rows
    out = m.inference(ids.cuda(), max_len_cap=400)
    assert out[1].shape[1] == int(ref_len.min())


def test_sharded_exact_length_batches_match_solo_runs():
    """BASELINE config 4 (reduced): exact-length batches reproduce every utterance's solo (B=1) reference run,
    decoded lengths exact (F8: padding would not)."""
    from b200taco import shard
    lens = [23, 17, 23, 20, 17, 23, 20, 17, 23, 20, 23, 17]
    corpus = [synth.phoneme_ids(1, t, seed=100 + i)[0] for i, t in enumerate(lens)]
    center = O.encoder(util.weights("stress", True), torch.stack([corpus[0], corpus[2]])).mean((0, 1))
    sd = util.weights("planted", True, ctx_center=center)
    m = gpu_model(sd)
    plan = shard.make_plan(lens, world_size=1, batch_size=3)
    assert all(lens[i] == b.text_len for b in plan.batches for i in b.indices)

    def decode(ids):
        return m.inference_per_utterance(ids.cuda())[4].tolist()

    got = shard.gather_lengths(shard.run_shard(plan, 0, corpus, decode), len(lens)).tolist()
    want = []
    for c in corpus:
        gate = O.tacotron2_inference(sd, c[None])[2]
        want.append(int(O.per_utterance_lengths(gate)[0]))
    assert got == want, (got, want)
    assert len(set(want)) > 1                       # the stops really differ between utterances


def test_batches_beyond_one_launch_are_chunked():
    """B > 64 rows: consecutive launches over row chunks reproduce the reference on the whole batch
    (teacher-forced, and free-running with the batch-wide stop = earliest stop of any chunk)."""
    B, T = 70, 21
    ids = synth.phoneme_ids(B, T, seed=21)
    center = O.encoder(util.weights("stress", True), ids[:2]).mean((0, 1))
    sd = util.weights("planted", True, ctx_center=center)
    m = gpu_model(sd)
    out = m.inference(ids.cuda())
    ref = O.tacotron2_inference(sd, ids)
    assert out[1].shape == ref[1].shape and len(out[3]) == len(ref[3])
    assert util.maxdiff(out[1].cpu(), ref[1]) <= TOL_MEL and util.maxdiff(out[0].cpu(), ref[0]) <= 5 * TOL_MEL
    assert util.maxdiff(util.stack(out[3]).cpu(), util.stack(ref[3])) <= TOL_ALIGN
    g = torch.Generator().manual_seed(3)
    mel = (torch.randn(B, 80, 7, generator=g) * 2 - 5)
    lens = torch.randint(5, T + 1, (B,), generator=g)
    out = m(ids.cuda(), mel.cuda(), lens.cuda())
    ref = O.tacotron2_forward(sd, ids, mel, lens)
    assert util.maxdiff(out[1].cpu(), ref[1]) <= TOL_MEL and util.maxdiff(out[2].cpu(), ref[2]) <= TOL_MEL
    assert util.maxdiff(util.stack(out[3]).cpu(), util.stack(ref[3])) <= TOL_ALIGN


def test_pipelined_batches_equal_one_by_one():
    """inference_batches overlaps the stages of consecutive batches on three streams; every batch must come out
    bit-identical to decoding it alone (different shapes per batch, host and device inputs)."""
    sd = util.weights("planted", True, ctx_center=O.encoder(util.weights("stress", True), synth.phoneme_ids(2, 30, seed=50)).mean((0, 1)))
    m = gpu_model(sd)
    batches = [synth.phoneme_ids(b, t, seed=60 + i) for i, (b, t) in enumerate([(5, 30), (32, 41), (3, 30), (8, 77), (1, 12)])]
    solo = [m.inference_per_utterance(x.cuda()) for x in batches]
    feed = [x.pin_memory() if i % 2 else x.cuda() for i, x in enumerate(batches)]
    outs = list(m.inference_batches(feed, to_host=True))
    assert len(outs) == len(batches)
    for (post, coarse, gate, aligns, lengths), o in zip(solo, outs):
        assert o[5] == coarse.shape[1] and torch.equal(o[4].cpu(), lengths.cpu())
        assert torch.equal(o[0], post) and torch.equal(o[1], coarse) and torch.equal(o[2], gate)
        assert torch.equal(o[3], torch.stack(aligns))
        assert torch.equal(o[6], post.cpu()) and torch.equal(o[7], lengths.cpu())


@pytest.mark.parametrize("rows,T,nb,cap", [(8, 41, 4, None), (32, 30, 3, None), (5, 41, 3, 9), (64, 23, 2, None)])
def test_continuous_batching_equals_batch_by_batch(rows, T, nb, cap):
    """Utterance queue (b200taco_decoder_infer_stream): rows of one persistent launch are refilled as utterances stop.
    Every utterance must come out bit-identical to its batch-by-batch decode under the per-utterance rule: same length,
    same frames / gate logits / alignments up to that length, zeros behind it."""
    sd = util.weights("planted", True, ctx_center=O.encoder(util.weights("stress", True), synth.phoneme_ids(2, 30, seed=50)).mean((0, 1)))
    m = gpu_model(sd)
    batches = [synth.phoneme_ids(rows, T, seed=80 + i).cuda() for i in range(nb)]
    solo = [m.inference_per_utterance(x, max_len_cap=cap) for x in batches]
    outs = list(m.inference_batches(batches, max_len_cap=cap, continuous=True, to_host=True))
    assert len(outs) == nb
    spread = set()
    for (post, coarse, gate, aligns, lengths), o in zip(solo, outs):
        L = lengths.cpu()
        assert torch.equal(o[4].cpu(), L) and torch.equal(o[7], L) and o[5] == int(L.max())
        al = torch.stack(aligns)
        for r in range(rows):
            n = int(L[r])
            spread.add(n)
            assert torch.equal(o[1][r, :n], coarse[r, :n]) and torch.equal(o[2][r, :n], gate[r, :n])
            assert torch.equal(o[0][r, :n], post[r, :n])                  # postnet saw zeros behind the row's length in both paths
            assert not coarse[r, n:].any()
            assert torch.equal(o[3][:n, r], al[:n, r])
            assert not o[1][r, n:].any() and not o[2][r, n:].any() and not o[3][n:, r].any()
        assert torch.equal(o[6], o[0].cpu()) and o[0].shape == o[1].shape
    assert cap is not None or len(spread) > 1, "the planted stop counter should give different lengths"
    assert m._get_engine().last_stream_steps < sum(s[1].shape[1] + 1 for s in solo) or cap is not None


def test_continuous_batching_rejects_what_it_cannot_run():
    """The utterance queue lives in the tensor-core decoder kernel: other configurations fail loudly, nothing falls back."""
    m = gpu_model(util.weights("default", True))
    eng = m._get_engine()
    mem = eng.encoder(synth.phoneme_ids(8, 20, seed=1).cuda())
    with pytest.raises(B200TacoError, match="multiple"):
        eng.decoder_infer_stream(mem[:7].contiguous(), 4, 10, 0.5)          # 7 utterances, 4 rows
    with pytest.raises(B200TacoError, match="tensor-core"):
        eng.decoder_infer_stream(mem[:4].contiguous(), 2, 10, 0.5)          # 2 rows: SIMT kernel territory
    with pytest.raises(B200TacoError, match="one shape"):
        list(eng.stream_infer([synth.phoneme_ids(8, 20, seed=1), synth.phoneme_ids(8, 21, seed=2)], 10, 0.5))
    m32 = gpu_model(util.weights("default", False), weight_dtype="fp32")
    with pytest.raises(B200TacoError, match="one shape|tensor-core"):
        list(m32.inference_batches([synth.phoneme_ids(8, 20, seed=1).cuda()], max_len_cap=5, continuous=True))
    with pytest.raises(ValueError):
        list(m.inference_batches([synth.phoneme_ids(8, 20, seed=1).cuda()], per_utterance=False, continuous=True))


def test_long_horizon_default_init_1000_steps():
    """Default init never fires (F11): 1000 frames, error must not grow with the horizon."""
    ids = synth.phoneme_ids(2, 150, seed=1)
    sd = util.weights("default", True)
    m = gpu_model(sd)
    out = m.inference(ids.cuda())
    with torch.no_grad():
        ref = O.tacotron2_inference(sd, ids)
    assert out[1].shape == ref[1].shape == (2, 1000, 80)
    gold = dict(mel_post=ref[0], mel_coarse=ref[1], gate=ref[2], align=util.stack(ref[3]))
    check(out, gold)


# ------------------------------------------------------------------ behaviour at the boundary
def test_repack_after_inplace_weight_update():
    """An optimizer-style in-place update bumps param._version; the packed cache must follow."""
    sd = util.weights("stress", True)
    m = gpu_model(sd)
    ids, lengths, mel = synth.c1_batch()
    args = (ids.cuda(), mel[:, :, :4].contiguous().cuda(), lengths.cuda())
    a = m(*args)[1].clone()
    with torch.no_grad():
        m.decoder.linear_projection.bias.add_(1.0)
    b = m(*args)[1]
    # frame 0 is decoded from the zero go-frame: only the bias changed
    assert float((b[:, 0] - a[:, 0] - 1.0).abs().max()) < 1e-5
    with torch.no_grad():
        m.decoder.decoder_lstm.weight_hh.zero_()
    c = m(*args)[1]
    assert float((c[:, 1] - b[:, 1]).abs().max()) > 1e-4


def test_errors():
    m = gpu_model(util.weights())
    ids = synth.phoneme_ids(2, 9).cuda()
    m.train()
    with pytest.raises(B200TacoError):
        m.inference(ids)
    m.eval()
    with pytest.raises(B200TacoError):                     # one launch carries at most b200taco_max_batch rows ...
        m._get_engine().encoder(synth.phoneme_ids(65, 9).cuda())
    assert m.inference(synth.phoneme_ids(65, 9).cuda(), max_len_cap=3)[1].shape == (65, 3, 80)   # ... the module chunks
    with pytest.raises(B200TacoError):
        m.inference(synth.phoneme_ids(1, 2000).cuda())     # beyond b200taco_max_text_len
    assert m._engine.launch_count() > 0


def test_stream_infer_c2_vs_oracle():
    """The path bench.py times, at its own configuration, against the ORACLE (not against another CUDA path): continuous
    batching (Engine.stream_infer) of K = 2 batches of 32 x 150 phonemes, the bench's planted stop counter (rate 8e-4, jitter
    6e-3), cap 1000.  Per-utterance decoded lengths exact; mel_coarse / gate <= 2e-5, alignments <= 2e-6 and mel_postnet
    <= 1e-4 up to each utterance's own length, zeros behind it.  The oracle decodes each batch with the batch-wide stop rule
    disabled (threshold above 1) for as many frames as the longest utterance needs, and is trimmed per row by
    train.py:71-79's rule; rows do not interact, so this is every utterance's solo decode."""
    import bench
    B, T, K = bench.B, bench.T, 2
    sd0 = util.weights("default")
    batches = [synth.phoneme_ids(B, T, seed=s) for s in (0, 17)]
    with torch.no_grad():
        center = O.encoder(sd0, batches[0]).mean((0, 1))
    sd = util.weights("default")
    with torch.no_grad():
        synth.plant_stop_counter_(sd, ctx_center=center, **bench.PLANT)
        synth.round_to_bf16_(sd)
    m = gpu_model(sd)
    outs = list(m.inference_batches([x.cuda() for x in batches], max_len_cap=bench.CAP, continuous=True))
    assert len(outs) == K
    for ids, (post, coarse, gate, align, lengths, S) in zip(batches, outs):
        L = lengths.cpu().tolist()
        with torch.no_grad():
            ref = O.tacotron2_inference(sd, ids, max_len_cap=max(L), gate_threshold=2.0)       # never stops early
        rpost, rcoarse, rgate, ralign = ref[0], ref[1], ref[2], util.stack(ref[3])
        want = O.per_utterance_lengths(rgate).tolist()                                         # train.py:71-79 with the steps > 1 guard
        assert L == want, [(a, b) for a, b in zip(L, want) if a != b]
        assert S == max(L) and min(L) >= 200 and max(L) - min(L) >= 100       # a real spread of stop frames
        for b, n in enumerate(L):
            with torch.no_grad():                            # the utterance's solo postnet: its own frames, zero padding behind
                solo = rcoarse[b:b + 1, :n]
                rpost_b = (solo + O.postnet(sd, solo.transpose(1, 2)).transpose(1, 2))[0]
            assert util.maxdiff(coarse[b, :n].cpu(), rcoarse[b, :n]) <= TOL_MEL
            assert util.maxdiff(gate[b, :n].cpu(), rgate[b, :n]) <= TOL_MEL
            assert util.maxdiff(align[:n, b].cpu(), ralign[:n, b]) <= TOL_ALIGN
            assert util.maxdiff(post[b, :n].cpu(), rpost_b) <= 5 * TOL_MEL
            assert not coarse[b, n:].any() and not gate[b, n:].any()


@pytest.mark.parametrize("rows,n_utts,cap", [(8, 37, 300), (32, 70, 300), (8, 19, 40)])
def test_ragged_continuous_batching_matches_solo_runs(rows, n_utts, cap):
    """Utterances of DIFFERENT text lengths in one persistent launch (Tacotron2.inference_corpus, BASELINE config 4): every
    utterance must decode as it does alone at its exact length - decoded length exact, frames / gate / alignments within the
    usual tolerances (the attention slices of a padded launch partition the positions differently, so not bit for bit).
    Several launches (length ratio), launches filled up to a multiple of the rows, cap hits (two of the utterances never fire
    their gate; the cap of 300 frames keeps the comparison of two float32 recurrences inside the horizon the tolerances are
    stated for)."""
    sd = util.weights("planted", True, ctx_center=O.encoder(util.weights("stress", True), synth.phoneme_ids(2, 30, seed=50)).mean((0, 1)))
    m = gpu_model(sd)
    g = torch.Generator().manual_seed(rows + n_utts)
    lens = torch.randint(12, 48, (n_utts,), generator=g).tolist()
    utts = [synth.phoneme_ids(1, L, seed=300 + i)[0] for i, L in enumerate(lens)]
    got = {i: (post, coarse, gate, al) for i, post, coarse, gate, al in m.inference_corpus(utts, max_len_cap=cap, rows=rows)}
    assert sorted(got) == list(range(n_utts))
    spread = set()
    for i, u in enumerate(utts):
        post, coarse, gate, aligns, lengths = m.inference_per_utterance(u[None].cuda(), max_len_cap=cap)
        n = int(lengths[0])
        spread.add(n)
        gp, gc, gg, ga = got[i]
        assert gc.shape == (n, 80) and ga.shape == (n, lens[i]), (i, gc.shape, n)
        # two CUDA paths, each within TOL of the oracle: 2 x TOL between them (the planted counter integrates over up to 1000 frames)
        assert util.maxdiff(gc.cpu(), coarse[0, :n].cpu()) <= TOL_MEL and util.maxdiff(gg.cpu(), gate[0, :n].cpu()) <= 2 * TOL_MEL
        assert util.maxdiff(ga.cpu(), torch.stack(aligns)[:n, 0].cpu()) <= 2 * TOL_ALIGN
        assert util.maxdiff(gp.cpu(), post[0, :n].cpu()) <= 5 * TOL_MEL
    assert cap < 100 or len(spread) > 3


def test_continuous_batching_whole_tuple_to_host():
    """to_host="all": the reference's whole 4-tuple of every batch lands in pinned host memory, equal to the device tensors."""
    m = gpu_model(util.weights("default", True))
    batches = [synth.phoneme_ids(8, 21, seed=90 + i).cuda() for i in range(2)]
    outs = list(m._get_engine().stream_infer(batches, 9, 0.5, to_host="all", live_rows=16))
    assert len(outs) == 2
    for o in outs:
        post, coarse, gate, align, lengths, S = o[:6]
        post_h, len_h, coarse_h, gate_h, align_h = o[6:]
        assert all(t.is_pinned() for t in (post_h, coarse_h, gate_h, align_h))
        assert torch.equal(post_h, post.cpu()) and torch.equal(coarse_h, coarse.cpu())
        assert torch.equal(gate_h, gate.cpu()) and torch.equal(align_h, align.cpu()) and len_h.tolist() == lengths.tolist()
        assert align_h.shape == (S, 8, 21)
