"""Batch assembly (src/data_utils.py:45-84, train.py:496-508) and vocoder hand-off (inference.py:71, train.py:71-79):
CPU oracle against golden vectors of the unmodified reference, device kernels (through the C ABI) against both.

Everything here is bit-exact: the path only moves, pads and reorders values.
"""
import importlib.util
import os

import pytest
import torch

import util
from oracle import data_oracle as DO

_spec = importlib.util.spec_from_file_location("make_golden_data", os.path.join(util.GOLDEN, "make_golden_data.py"))
_mgd = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mgd)


def test_collate_oracle_matches_reference_golden():
    g = util.golden("data_b5")
    out = DO.collate(_mgd.data_case())
    for got, key in zip(out, ("text_padded", "input_lengths", "mel_padded", "mel_lengths")):
        assert torch.equal(got, g[key]), key
    # the reference's quirk: mel_lengths stays in sample order while the rows are sorted by text length
    assert g["mel_lengths"].tolist() == [41, 25, 60, 31, 37] and g["input_lengths"].tolist() == [23, 17, 14, 9, 9]
    prep = DO.prepare_training_batch(out)
    for got, key in zip(prep, ("t_text_padded", "t_input_lengths", "t_mel_padded", "t_mel_lengths", "t_gate_target")):
        assert torch.equal(got, g[key]), key


def test_handoff_oracle_matches_reference_golden():
    g = util.golden("data_b5")
    mel_post, gate, fallback = _mgd.handoff_case()
    ends = DO.stop_lengths(gate, fallback)
    assert ends == g["h_ends"].tolist()
    assert torch.equal(DO.vocoder_layout(mel_post), g["h_mel_for_vocoder"])
    assert torch.equal(DO.vocoder_layout(mel_post, ends), g["h_trimmed"])


def test_collate_has_no_cpu_fallback():
    from b200taco._cabi import B200TacoError
    from b200taco.data import TextMelCollate
    with pytest.raises(B200TacoError, match="CUDA"):
        TextMelCollate(device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("mirror", [True, False])
def test_device_collate_matches_reference_golden(mirror):
    from b200taco.data import TextMelCollate, prepare_training_batch
    g = util.golden("data_b5")
    collate = TextMelCollate(mirror_reference_mel_lengths=mirror)
    for _ in range(2):                                  # the second call reuses the pinned staging buffers
        out = collate(_mgd.data_case())
    assert all(t.is_cuda for t in out)
    for got, key in zip(out[:3], ("text_padded", "input_lengths", "mel_padded")):
        assert torch.equal(got.cpu(), g[key]), key
    if mirror:
        assert torch.equal(out[3].cpu(), g["mel_lengths"])
        prep = prepare_training_batch(out)
        for got, key in zip(prep, ("t_text_padded", "t_input_lengths", "t_mel_padded", "t_mel_lengths", "t_gate_target")):
            assert torch.equal(got.cpu(), g[key]), key
    else:                                               # row order: every row's own mel length
        rows = [int((g["mel_padded"][i].abs().sum(0) > 0).nonzero().max()) + 1 for i in range(5)]
        assert out[3].cpu().tolist() == rows


@pytest.mark.gpu
def test_device_collate_ragged_edge_cases():
    """Single sample, length-1 text, a mel of one frame, equal lengths, a larger batch: device vs oracle, bit exact."""
    from b200taco.data import TextMelCollate
    collate = TextMelCollate()
    g = torch.Generator().manual_seed(5)
    cases = [[(torch.randint(0, 72, (1,), generator=g), torch.randn(80, 1, generator=g))],
             [(torch.randint(0, 72, (7,), generator=g), torch.randn(80, 13, generator=g)) for _ in range(3)],
             [(torch.randint(0, 72, (int(L),), generator=g), torch.randn(80, int(S), generator=g))
              for L, S in zip(torch.randint(1, 200, (64,), generator=g), torch.randint(1, 900, (64,), generator=g))]]
    for batch in cases:
        got, want = collate(batch), DO.collate(batch)
        for a, b in zip(got, want):
            assert torch.equal(a.cpu(), b)


@pytest.mark.gpu
def test_gate_targets_kernel():
    from b200taco.loss import gate_targets
    ml = torch.tensor([1, 5, 12, 12, 3])
    got = gate_targets(ml.cuda(), 12).cpu()
    want = torch.zeros(5, 12)
    for j, l in enumerate(ml.tolist()):
        want[j, l - 1:] = 1                                            # train.py:506-508
    assert torch.equal(got, want)


@pytest.mark.gpu
def test_vocoder_handoff_matches_reference_golden():
    """Stand-alone layout kernel and the postnet epilogue variant (tensor-core and CUDA-core paths) against the golden hand-off."""
    from b200taco import synth
    g = util.golden("data_b5")
    mel_post, gate, fallback = _mgd.handoff_case()
    ends = DO.stop_lengths(gate, fallback)
    m = util.Tacotron2().eval()
    m.load_state_dict(util.weights("default"))
    m = m.cuda()
    eng = m._get_engine()
    assert torch.equal(eng.vocoder_layout(mel_post.cuda()).cpu(), g["h_mel_for_vocoder"])
    lengths = torch.tensor(ends, dtype=torch.int32).cuda()
    assert torch.equal(eng.vocoder_layout(mel_post.cuda(), lengths).cpu(), g["h_trimmed"])
    assert torch.equal(eng.vocoder_layout(mel_post.cuda(), lengths, frames=max(ends)).cpu(), g["h_trimmed"][:, :, :max(ends)])
    # the postnet's own epilogue: channel-major copy == transpose of its (B,S,80) result, zeroed behind each length
    coarse = synth.mel_targets(3, 45, seed=9).transpose(1, 2).contiguous().cuda()
    lens = torch.tensor([45, 17, 30], dtype=torch.int32).cuda()
    for path in ("auto", "simt"):
        m.dense_path = path
        eng = m._get_engine()
        from b200taco.engine import zero_beyond_lengths_
        zc = zero_beyond_lengths_(coarse.clone(), lens)
        post = eng.postnet(zc, lens)                      # every layer zero behind a row's own length
        post2, voc = eng.postnet_for_vocoder(zc, lens, frames=45)
        assert torch.equal(post, post2)
        assert torch.equal(voc.cpu(), DO.vocoder_layout(post.cpu(), lens.tolist()))
        for b, n in enumerate(lens.tolist()):             # = the solo run of row b at its own length
            solo = eng.postnet(zc[b:b + 1, :n].contiguous())
            assert util.maxdiff(post[b, :n].cpu(), solo[0].cpu()) <= 1e-6 and not post[b, n:].any()
        _, voc_all = eng.postnet_for_vocoder(coarse, None, frames=45)          # no lengths: plain transpose of the plain postnet
        assert torch.equal(voc_all.cpu(), DO.vocoder_layout(eng.postnet(coarse).cpu()))


@pytest.mark.gpu
def test_inference_for_vocoder_matches_per_utterance_inference():
    from b200taco import synth
    gold = util.golden("infer_planted_b5_t41")
    m = util.Tacotron2().eval()
    m.load_state_dict(util.weights("planted", True, ctx_center=gold["ctx_center"]))
    m = m.cuda()
    ids = synth.phoneme_ids(5, 41, seed=7).cuda()
    voc, lengths = m.inference_for_vocoder(ids)
    post, _, _, _, lengths2 = m.inference_per_utterance(ids)
    assert lengths.tolist() == lengths2.tolist() == gold["solo_lengths"].tolist()
    assert voc.shape == (5, 80, int(lengths.max()))
    assert torch.equal(voc.cpu(), DO.vocoder_layout(post.cpu(), lengths.tolist()))
