"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/b200taco.h declares, the Python mirror keeps the reference's surface (state_dict
layout, signatures, error behaviour) and fails loudly without CUDA (no fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import util
from b200taco import _cabi, config, synth
from b200taco._cabi import B200TacoError
from b200taco.model import Tacotron2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200taco.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.lib()


DEBUG_HEADER = os.path.join(ROOT, "include", "b200taco_debug.h")     # measurement / bring-up aids, outside the product ABI


def declared_symbols(with_debug=True):
    src = open(HEADER).read() + (open(DEBUG_HEADER).read() if with_debug else "")
    return sorted(set(re.findall(r"B200TACO_API\s+[\w\s\*]+?\b(b200taco_\w+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200taco.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names          # the ctypes table covers exactly the two headers
    product = declared_symbols(with_debug=False)      # the product ABI carries no measurement / debug entry points
    assert not [n for n in product if "measure" in n or "debug" in n or "trace" in n or "profile" in n], product
    assert lib.b200taco_abi_version() == _cabi.ABI_VERSION
    assert lib.b200taco_max_batch() == 64 and lib.b200taco_max_text_len() >= 512


def test_weights_struct_matches_header():
    src = open(HEADER).read()
    body = src[src.index("typedef struct {", src.index("b200taco_hparams;")):src.index("} b200taco_weights;")]
    fields = []
    for name, arr in re.findall(r"const float\*\s+(\w+)(?:\[(\d)\])?;", body):
        fields.append((name, int(arr) if arr else None))
    mine = [(n, (t._length_ if hasattr(t, "_length_") else None)) for n, t in _cabi.Weights._fields_]
    assert mine == fields
    n_ptr = sum(a or 1 for _, a in fields)
    assert ctypes.sizeof(_cabi.Weights) == 8 * n_ptr


def test_create_rejects_wrong_hparams_and_reports(lib):
    hp = _cabi.HParams(**{**config.as_hparams(), "rnn_dim": 512})
    h = ctypes.c_void_p()
    st = lib.b200taco_create(ctypes.byref(hp), 0, ctypes.byref(h))
    assert st == -2 and b"hparams" in lib.b200taco_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_without_gpu_fails_loudly(lib):
    hp = _cabi.HParams(**config.as_hparams())
    h = ctypes.c_void_p()
    st = lib.b200taco_create(ctypes.byref(hp), 0, ctypes.byref(h))
    assert st == -4 and not h.value


def test_state_dict_layout_matches_reference_contract():
    m = Tacotron2()
    sd = m.state_dict()
    assert len(sd) == 86
    assert sum(p.numel() for p in m.parameters()) == 28_154_243
    shapes = {
        "encoder.embedding.weight": (72, 512), "encoder.convolutions.2.0.weight": (512, 512, 5),
        "encoder.convolutions.0.1.running_var": (512,), "encoder.lstm.weight_hh_l0_reverse": (1024, 256),
        "decoder.prenet.layers.0.weight": (256, 80), "decoder.attention.energy_scale": (),
        "decoder.attention.location_conv.weight": (32, 2, 31), "decoder.attention.v.bias": (1,),
        "decoder.attention_lstm.weight_ih": (4096, 768), "decoder.decoder_lstm.weight_ih": (4096, 1536),
        "decoder.linear_projection.weight": (80, 1536), "decoder.gate_layer.bias": (1,),
        "postnet.convolutions.0.0.weight": (512, 80, 5), "postnet.convolutions.4.0.weight": (80, 512, 5),
        "postnet.convolutions.4.1.num_batches_tracked": (),
    }
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == s, k
    assert float(sd["decoder.gate_layer.bias"]) == -3.0 and abs(float(sd["decoder.attention.energy_scale"]) - 1.2) < 1e-7
    # attribute tree used by train.py:345,358 and src/model.py:497
    assert len(list(m.decoder.attention.parameters())) == 7
    assert m.decoder.gate_threshold == 0.5 and m.n_mels == 80 and m.decoder.max_decoder_steps == 1000


def test_signatures_match_reference():
    f = inspect.signature(Tacotron2.forward)
    assert list(f.parameters) == ["self", "text_inputs", "mel_targets", "text_lengths", "use_postnet"]
    assert f.parameters["text_lengths"].default is None and f.parameters["use_postnet"].default is True
    i = inspect.signature(Tacotron2.inference)
    assert list(i.parameters) == ["self", "text_inputs", "text_lengths", "max_len_cap", "gate_threshold"]
    assert inspect.signature(Tacotron2.__init__).parameters.keys() == {"self"}


def test_no_cpu_fallback():
    m = Tacotron2().eval()
    ids, lengths, mel = synth.c1_batch()
    with pytest.raises(B200TacoError):
        m.inference(ids)
    with pytest.raises(B200TacoError):
        m(ids, mel, lengths)
    with pytest.raises(B200TacoError):
        m.encoder(ids)               # parameter holders are not callable


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch-tacotron2-implementation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|oracle/|tacotron2_oracle|step_fp64", src, re.M), \
                    f"{f} reaches into oracle/"


def test_synth_inputs_are_deterministic():
    a, b = synth.phoneme_ids(4, 9), synth.phoneme_ids(4, 9)
    assert torch.equal(a, b) and int(a.max()) < 72 and int(a.min()) >= 0
    ids, lengths, mel = synth.c1_batch()
    assert ids.shape == (4, 37) and mel.shape == (4, 80, 120) and lengths.tolist() == [37, 29, 21, 12]
    assert float(mel.max()) <= 2.0 and float(mel.min()) >= -11.52
