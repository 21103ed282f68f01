"""Drop-in ``Tacotron2`` whose arithmetic runs in libb200taco.so (CUDA, sm_100a).

Boundary preserved (reference src/model.py:410-501):

* ``Tacotron2()`` takes no arguments; hparams come from :mod:`b200taco.config`.
* ``forward(text_inputs, mel_targets, text_lengths=None, use_postnet=True)`` and
  ``inference(text_inputs, text_lengths=None, max_len_cap=None, gate_threshold=None)``
  return ``(mel_postnet (B,S,80), mel_coarse (B,S,80), gate_logits (B,S), alignments)``
  with ``alignments`` a *list* of S tensors (B,T) as the reference's callers expect
  (train.py:198, :227).
* identical parameter tree, so reference checkpoints load unchanged and
  ``model.decoder.attention.parameters()`` (train.py:345) resolves.

The nn.Module tree below only *holds* parameters (created in the reference's order so
that ``torch.manual_seed(s); Tacotron2()`` yields the same random initialisation); none
of the sub-modules is ever called.  All compute goes through :class:`b200taco.engine.Engine`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import config
from ._cabi import B200TacoError
from .engine import Engine, STOP_BATCH_WIDE, STOP_PER_UTTERANCE, zero_beyond_lengths_


def _conv_bn(c_in: int, c_out: int, k: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv1d(c_in, c_out, k, padding=(k - 1) // 2), nn.BatchNorm1d(c_out))


class _Holder(nn.Module):
    """Parameter container; calling it is a bug (compute lives in the CUDA library)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise B200TacoError(f"{type(self).__name__} holds parameters only; call Tacotron2.forward/inference")


class Encoder(_Holder):
    """Parameters of src/model.py:15-48."""

    def __init__(self):
        super().__init__()
        d = config.encoder_embedding_dim
        self.embedding = nn.Embedding(len(config.SYMBOLS), config.symbols_embedding_dim)
        self.convolutions = nn.ModuleList(
            _conv_bn(d, d, config.encoder_kernel_size) for _ in range(config.encoder_n_convolutions))
        self.lstm = nn.LSTM(d, d // 2, 1, batch_first=True, bidirectional=True)


class LocationSensitiveAttention(_Holder):
    """Parameters of src/model.py:86-102."""

    def __init__(self):
        super().__init__()
        a = config.attention_dim
        self.query_layer = nn.Linear(config.attention_rnn_dim, a, bias=False)
        self.memory_layer = nn.Linear(config.encoder_embedding_dim, a, bias=False)
        k = config.location_kernel_size
        self.location_conv = nn.Conv1d(2, config.location_n_filters, k, padding=(k - 1) // 2, bias=False)
        self.location_dense = nn.Linear(config.location_n_filters, a, bias=False)
        self.v = nn.Linear(a, 1, bias=True)
        self.energy_scale = nn.Parameter(torch.tensor(1.2))


class PreNet(_Holder):
    """Parameters of src/model.py:159-164."""

    def __init__(self, in_dim, sizes):
        super().__init__()
        dims = [in_dim] + list(sizes)
        self.layers = nn.ModuleList(nn.Linear(i, o, bias=False) for i, o in zip(dims[:-1], dims[1:]))


class Decoder(_Holder):
    """Parameters and scalar attributes of src/model.py:179-217."""

    def __init__(self):
        super().__init__()
        self.n_mels = config.n_mels
        self.max_decoder_steps = config.max_decoder_steps
        self.gate_threshold = config.gate_threshold
        h, e = config.decoder_rnn_dim, config.encoder_embedding_dim
        self.prenet = PreNet(self.n_mels, [config.prenet_dim, config.prenet_dim])
        self.attention = LocationSensitiveAttention()
        self.attention_lstm = nn.LSTMCell(config.prenet_dim + e, h)
        self.decoder_lstm = nn.LSTMCell(h + e, h)
        self.linear_projection = nn.Linear(h + e, self.n_mels)
        self.gate_layer = nn.Linear(h + e, 1)
        with torch.no_grad():
            self.gate_layer.bias.fill_(-3.0)          # src/model.py:215-217


class PostNet(_Holder):
    """Parameters of src/model.py:341-390."""

    def __init__(self):
        super().__init__()
        d, k, n = config.postnet_embedding_dim, config.postnet_kernel_size, config.postnet_n_convolutions
        chans = [config.n_mels] + [d] * (n - 1) + [config.n_mels]
        self.convolutions = nn.ModuleList(_conv_bn(chans[i], chans[i + 1], k) for i in range(n))


class Tacotron2(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.postnet = PostNet()
        self.n_mels = config.n_mels
        self._projection_bias_initialized = False      # src/model.py:422 (plain attribute, F10)
        self._engine = None
        # storage of the decoder LSTM matrices in the packed cache: "bf16" (default, half the
        # L2 traffic per step) or "fp32" (the reference's values bit for bit)
        self.lstm_weight_dtype = "bf16"
        # train-mode forward: None = the same packed cache as inference, i.e. MIXED PRECISION with the default "bf16" - the
        # forward multiplies bf16-rounded decoder LSTM matrices (fp32 state and accumulation) while the hand-written backward
        # differentiates with the fp32 master weights, as the reference's own loop does under torch.autocast (train.py:510);
        # "fp32" = a second engine that packs the master weights unrounded (forward and backward see identical numbers)
        self.train_lstm_weight_dtype = None
        self._train_engine = None
        # "auto": tcgen05 tensor-core gate contractions when B >= 4 and bf16 weights; "simt": never
        self.decoder_path = "auto"
        # dense layers: "auto" = postnet / teacher-forced prenet on tcgen05 (3xTF32), encoder on fp32 CUDA cores;
        # "tc" = everything on tcgen05; "simt" = everything on CUDA cores (include/b200taco.h)
        self.dense_path = "auto"

    # ------------------------------------------------------------------ plumbing
    def _tensors(self) -> dict:
        d = dict(self.named_parameters())
        d.update(dict(self.named_buffers()))
        return d

    def _get_engine(self) -> Engine:
        dev = self.decoder.gate_layer.weight.device
        if (self._engine is None or self._engine.device != dev
                or self._engine.weight_dtype != self.lstm_weight_dtype
                or self._engine.decoder_path != self.decoder_path
                or self._engine.dense_path != self.dense_path):
            self._engine = Engine(dev, self.lstm_weight_dtype, self.decoder_path, self.dense_path)
        self._engine.ensure_packed(self._tensors())
        return self._engine

    def _get_train_engine(self) -> Engine:
        want = self.train_lstm_weight_dtype or self.lstm_weight_dtype
        if want == self.lstm_weight_dtype:
            return self._get_engine()
        dev = self.decoder.gate_layer.weight.device
        if (self._train_engine is None or self._train_engine.device != dev or self._train_engine.weight_dtype != want
                or self._train_engine.decoder_path != self.decoder_path or self._train_engine.dense_path != self.dense_path):
            self._train_engine = Engine(dev, want, self.decoder_path, self.dense_path)
        self._train_engine.ensure_packed(self._tensors())
        return self._train_engine

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None                             # .to(device) / .float(): rebuild lazily
        self._train_engine = None
        return out

    def _check_mode(self, what):
        if self.training:
            raise B200TacoError(
                f"{what}: free-running inference in train() mode is not supported (the reference only calls it in "
                "eval mode, train.py:60, inference.py:66); call model.eval().")

    def _maybe_init_projection_bias(self, mel_targets):
        """F10 side effect, src/model.py:424-441: first forward sets the mel bias to the target mean."""
        if self._projection_bias_initialized:
            return
        bias = self.decoder.linear_projection.bias
        with torch.no_grad():
            means = mel_targets.mean(dim=(0, 2))
            if means.numel() == bias.numel():
                bias.copy_(means)
                self._projection_bias_initialized = True

    # ------------------------------------------------------------------ public API
    def forward(self, text_inputs, mel_targets, text_lengths=None, use_postnet=True):
        """Teacher-forced pass, src/model.py:454-480.  eval(): inference kernels, no autograd.  train(): dropout,
        BatchNorm batch statistics (running statistics updated) and autograd through the hand-written backward."""
        if not self._projection_bias_initialized:
            self._maybe_init_projection_bias(mel_targets)
        if self.training:
            return self._forward_train(text_inputs, mel_targets, text_lengths, use_postnet)
        with torch.no_grad():
            return self._forward_eval(text_inputs, mel_targets, text_lengths, use_postnet)

    def _forward_train(self, text_inputs, mel_targets, text_lengths, use_postnet):
        """Train-mode forward: three autograd stages over the C ABI (b200taco/train.py).  The dropout masks are drawn
        here with PyTorch's device generator; one launch carries at most ``engine.max_batch`` rows and BatchNorm
        statistics span the whole batch, so larger batches are not chunked."""
        from . import train as T
        eng = self._get_train_engine()                             # (re)packs the decoder weights if a parameter changed
        B, S = text_inputs.shape[0], mel_targets.shape[2]
        if B > eng.max_batch:
            raise B200TacoError(f"train-mode forward carries at most {eng.max_batch} rows per call (got {B})")
        dev = text_inputs.device
        named = dict(self.named_parameters())
        enc_p = [named[k] for k in T.encoder_param_keys()]
        dec_p = [named[k] for k in T.decoder_param_keys()]
        post_p = [named[k] for k in T.postnet_param_keys()]
        memory = T.EncoderTrainFn.apply(eng, text_inputs, *enc_p)
        n_enc = B * text_inputs.shape[1]
        for bn, (mean, var) in zip([c[1] for c in self.encoder.convolutions], eng._enc_bn_stats):
            T.update_running_stats(bn, mean, var, n_enc)
        if text_lengths is None:
            text_lengths = torch.full((B,), text_inputs.shape[1], dtype=torch.long, device=dev)     # src/model.py:462-469
        coarse, gate, align = T.DecoderTrainFn.apply(eng, memory, mel_targets, text_lengths, T.draw_decoder_masks(B, S, dev), *dec_p)
        if use_postnet:
            post = T.PostnetTrainFn.apply(eng, coarse, T.draw_postnet_masks(B, S, dev), *post_p)
            for bn, (mean, var) in zip([c[1] for c in self.postnet.convolutions], eng._post_bn_stats):
                T.update_running_stats(bn, mean, var, B * S)
        else:
            post = coarse
        return post, coarse, gate, list(align.unbind(0))

    def _forward_eval(self, text_inputs, mel_targets, text_lengths=None, use_postnet=True):
        """Rows are independent in eval mode, so batches beyond the library's per-launch limit
        (``engine.max_batch`` = 64 rows) run as consecutive launches over row chunks."""
        eng = self._get_engine()
        outs = []
        for lo in range(0, text_inputs.shape[0], eng.max_batch):
            hi = lo + eng.max_batch
            memory = eng.encoder(text_inputs[lo:hi])
            tl = text_lengths[lo:hi] if text_lengths is not None else None
            coarse, gate, align = eng.decoder_teacher_forced(memory, mel_targets[lo:hi], tl)
            post = eng.postnet(coarse) if use_postnet else coarse
            outs.append((post, coarse, gate, align))
        if len(outs) == 1:
            post, coarse, gate, align = outs[0]
        else:
            post, coarse, gate = (torch.cat([o[i] for o in outs], dim=0) for i in range(3))
            align = torch.cat([o[3] for o in outs], dim=1)              # (S, B, T)
            if not use_postnet:
                post = coarse                                           # src/model.py:479-480: the same tensor
        return post, coarse, gate, list(align.unbind(0))

    @torch.no_grad()
    def inference(self, text_inputs, text_lengths=None, max_len_cap=None, gate_threshold=None):
        """Free-running pass, src/model.py:482-501.  ``text_lengths`` is ignored, as there (F8)."""
        post, coarse, gate, align, _ = self._infer(text_inputs, max_len_cap, gate_threshold, STOP_BATCH_WIDE)
        return post, coarse, gate, list(align.unbind(0))

    @torch.no_grad()
    def inference_per_utterance(self, text_inputs, max_len_cap=None, gate_threshold=None):
        """Same decode, but every row keeps running until *its own* gate fires (train.py:71-79 rule).

        Only valid for exact-length batches (no padding; see F8).  Returns the 4-tuple plus
        ``lengths`` (B,) int32; rows are padded to the longest utterance.
        """
        post, coarse, gate, align, lengths = self._infer(text_inputs, max_len_cap, gate_threshold,
                                                         STOP_PER_UTTERANCE)
        return post, coarse, gate, list(align.unbind(0)), lengths

    @torch.no_grad()
    def inference_for_vocoder(self, text_inputs, max_len_cap=None, gate_threshold=None):
        """Free-running inference that hands the mel straight to a vocoder: ``(mel_for_vocoder (B,80,S), lengths (B,))``.

        The reference computes ``mel_outputs_postnet.transpose(1, 2)`` (inference.py:71) and trims each utterance at its own
        stop frame on the host (train.py:71-79: first frame whose sigmoid(gate) > 0.5, plus one).  Here every row decodes
        until its own gate fires, S is the longest utterance, and the last postnet layer writes the channel-major tensor
        from its epilogue with the frames behind each row's length zeroed."""
        self._check_mode("Tacotron2.inference_for_vocoder")
        thr = gate_threshold if gate_threshold is not None else self.decoder.gate_threshold
        cap = self.decoder.max_decoder_steps
        if max_len_cap:
            cap = min(cap, int(max_len_cap))
        eng = self._get_engine()
        if text_inputs.shape[0] > eng.max_batch:
            coarse, _, _, lengths, S = self._infer_chunked(eng, text_inputs, cap, thr, STOP_PER_UTTERANCE)
        else:
            coarse, _, _, lengths, S = eng.decoder_infer(eng.encoder(text_inputs), cap, thr, STOP_PER_UTTERANCE)
        _, voc = eng.postnet_for_vocoder(zero_beyond_lengths_(coarse.contiguous(), lengths), lengths, S)
        return voc, lengths

    @staticmethod
    def _infer_chunked(eng, text_inputs, cap, thr, stop_mode):
        """More rows than one launch carries: rows never interact, so the chunks decode one after the other.
        Batch-wide stop (src/model.py:327) = the first step at which ANY row fires = the minimum over the chunks'
        own stop steps; per-utterance stop: every chunk runs to its own longest row, results padded to the longest."""
        parts = []
        for lo in range(0, text_inputs.shape[0], eng.max_batch):
            memory = eng.encoder(text_inputs[lo:lo + eng.max_batch])
            parts.append(eng.decoder_infer(memory, cap, thr, stop_mode))
        if stop_mode == STOP_BATCH_WIDE:
            S = min(p[4] for p in parts)
            lengths = torch.cat([torch.clamp(p[3], max=S) for p in parts])
        else:
            S = max(p[4] for p in parts)
            lengths = torch.cat([p[3] for p in parts])

        def fit(t, dim):          # truncate / zero-pad along the frame axis
            n = t.shape[dim]
            if n >= S:
                return t.narrow(dim, 0, S)
            pad = list(t.shape)
            pad[dim] = S - n
            return torch.cat([t, t.new_zeros(pad)], dim=dim)

        coarse = torch.cat([fit(p[0], 1) for p in parts], dim=0)
        gate = torch.cat([fit(p[1], 1) for p in parts], dim=0)
        align = torch.cat([fit(p[2], 0) for p in parts], dim=1)
        return coarse, gate, align, lengths, S

    @torch.no_grad()
    def inference_batches(self, batches, max_len_cap=None, gate_threshold=None, per_utterance=True, to_host=False,
                          continuous=False, live_rows=64):
        """Corpus synthesis: free-running inference over an iterable of (B,T) id batches (device or pinned host
        tensors), with the encoder / decoder / postnet of consecutive batches overlapped on separate CUDA streams
        (``Engine.pipeline_infer``).  Yields per batch ``(mel_postnet, mel_coarse, gate_logits, alignments (S,B,T),
        lengths, S)`` (+ pinned host copies of mel_postnet and lengths with ``to_host``).  Each batch is decoded
        exactly as ``inference_per_utterance`` / ``inference`` would decode it alone.

        ``continuous=True`` (per-utterance stop rule, batches of one shape, >= 4 rows): continuous batching - the B
        decoder rows of one persistent kernel are refilled from the following batches as their utterances stop
        (``Engine.stream_infer``).  Every utterance's own frames ``[0, length)`` are bit-identical to the batch-by-batch
        path; frames of a row beyond its own length are zero there instead of that row's continued decode."""
        self._check_mode("Tacotron2.inference_batches")
        thr = gate_threshold if gate_threshold is not None else self.decoder.gate_threshold
        cap = self.decoder.max_decoder_steps
        if max_len_cap:
            cap = min(cap, int(max_len_cap))
        eng = self._get_engine()
        if continuous:
            if not per_utterance:
                raise ValueError("continuous batching decodes under the per-utterance stop rule")
            yield from eng.stream_infer(batches, cap, thr, to_host, live_rows=live_rows)
            return
        yield from eng.pipeline_infer(batches, cap, thr, STOP_PER_UTTERANCE if per_utterance else STOP_BATCH_WIDE, to_host)

    @torch.no_grad()
    def inference_corpus(self, utterances, max_len_cap=None, gate_threshold=None, rows=64, to_host=False, **launch_args):
        """Corpus synthesis over utterances of DIFFERENT lengths (BASELINE config 4): ``utterances`` is a sequence of 1-D id
        tensors; every utterance is encoded and decoded at its own exact length (the reference's ``inference`` has no pad
        mask, F8) under the per-utterance stop rule, ``rows`` of them at a time inside persistent decoder launches whose rows
        are refilled as utterances stop (``Engine.stream_infer_ragged``).  Yields ``(index, mel_postnet (S_i,80), mel_coarse
        (S_i,80), gate_logits (S_i,), alignments (S_i,T_i))`` per utterance, each equal to that utterance's solo
        ``inference_per_utterance`` result."""
        self._check_mode("Tacotron2.inference_corpus")
        thr = gate_threshold if gate_threshold is not None else self.decoder.gate_threshold
        cap = self.decoder.max_decoder_steps
        if max_len_cap:
            cap = min(cap, int(max_len_cap))
        yield from self._get_engine().stream_infer_ragged(utterances, cap, thr, rows=rows, to_host=to_host, **launch_args)

    def _infer(self, text_inputs, max_len_cap, gate_threshold, stop_mode):
        self._check_mode("Tacotron2.inference")
        thr = gate_threshold if gate_threshold is not None else self.decoder.gate_threshold
        cap = self.decoder.max_decoder_steps
        if max_len_cap:
            cap = min(cap, int(max_len_cap))
        eng = self._get_engine()
        if text_inputs.shape[0] > eng.max_batch:
            coarse, gate, align, lengths, S = self._infer_chunked(eng, text_inputs, cap, thr, stop_mode)
        else:
            memory = eng.encoder(text_inputs)
            coarse, gate, align, lengths, S = eng.decoder_infer(memory, cap, thr, stop_mode)
        coarse = coarse.contiguous()
        if stop_mode == STOP_PER_UTTERANCE:
            zero_beyond_lengths_(coarse, lengths)          # engine.zero_beyond_lengths_: every row as if decoded alone
        if S >= cap:
            print(f"Warning! Reached decode cap ({cap}).")                  # src/model.py:330
        if S < 3:
            print(f"[WARN] Very short mel length ({S}) - possible premature stop. "
                  f"Gate threshold={self.decoder.gate_threshold}")          # src/model.py:496-497
        post = eng.postnet(coarse, lengths if stop_mode == STOP_PER_UTTERANCE else None)
        return post, coarse, gate.contiguous(), align, lengths
