"""Host-side owner of one ``b200taco_handle``: weight packing, workspace, launches.

PyTorch is used here for device memory (the caching allocator owns every buffer the
kernels touch) and for the current CUDA stream; all arithmetic happens inside
``libb200taco.so`` behind the C ABI of include/b200taco.h.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, config
from ._cabi import B200TacoError, check

STOP_BATCH_WIDE = 0      # Decoder.inference rule, src/model.py:327
STOP_PER_UTTERANCE = 1   # caller-side trim rule, train.py:71-79, evaluated on device
WEIGHT_DTYPES = {"bf16": 0, "fp32": 1}   # storage of the four decoder LSTM matrices (include/b200taco.h)

# state_dict key -> field of b200taco_weights (index for array fields)
_WEIGHT_MAP = [("encoder.embedding.weight", "enc_embedding", None)]
for _i in range(3):
    _p = f"encoder.convolutions.{_i}"
    _WEIGHT_MAP += [(_p + ".0.weight", "enc_conv_w", _i), (_p + ".0.bias", "enc_conv_b", _i),
                    (_p + ".1.weight", "enc_bn_w", _i), (_p + ".1.bias", "enc_bn_b", _i),
                    (_p + ".1.running_mean", "enc_bn_mean", _i), (_p + ".1.running_var", "enc_bn_var", _i)]
for _i, _s in enumerate(("", "_reverse")):
    _WEIGHT_MAP += [(f"encoder.lstm.weight_ih_l0{_s}", "enc_lstm_w_ih", _i),
                    (f"encoder.lstm.weight_hh_l0{_s}", "enc_lstm_w_hh", _i),
                    (f"encoder.lstm.bias_ih_l0{_s}", "enc_lstm_b_ih", _i),
                    (f"encoder.lstm.bias_hh_l0{_s}", "enc_lstm_b_hh", _i)]
_WEIGHT_MAP += [
    ("decoder.prenet.layers.0.weight", "prenet_w", 0), ("decoder.prenet.layers.1.weight", "prenet_w", 1),
    ("decoder.attention.query_layer.weight", "attn_query_w", None),
    ("decoder.attention.memory_layer.weight", "attn_memory_w", None),
    ("decoder.attention.location_conv.weight", "attn_loc_conv_w", None),
    ("decoder.attention.location_dense.weight", "attn_loc_dense_w", None),
    ("decoder.attention.v.weight", "attn_v_w", None), ("decoder.attention.v.bias", "attn_v_b", None),
    ("decoder.attention.energy_scale", "attn_energy_scale", None),
    ("decoder.attention_lstm.weight_ih", "attn_lstm_w_ih", None),
    ("decoder.attention_lstm.weight_hh", "attn_lstm_w_hh", None),
    ("decoder.attention_lstm.bias_ih", "attn_lstm_b_ih", None),
    ("decoder.attention_lstm.bias_hh", "attn_lstm_b_hh", None),
    ("decoder.decoder_lstm.weight_ih", "dec_lstm_w_ih", None),
    ("decoder.decoder_lstm.weight_hh", "dec_lstm_w_hh", None),
    ("decoder.decoder_lstm.bias_ih", "dec_lstm_b_ih", None),
    ("decoder.decoder_lstm.bias_hh", "dec_lstm_b_hh", None),
    ("decoder.linear_projection.weight", "proj_w", None), ("decoder.linear_projection.bias", "proj_b", None),
    ("decoder.gate_layer.weight", "gate_w", None), ("decoder.gate_layer.bias", "gate_b", None),
]
for _i in range(5):
    _p = f"postnet.convolutions.{_i}"
    _WEIGHT_MAP += [(_p + ".0.weight", "post_conv_w", _i), (_p + ".0.bias", "post_conv_b", _i),
                    (_p + ".1.weight", "post_bn_w", _i), (_p + ".1.bias", "post_bn_b", _i),
                    (_p + ".1.running_mean", "post_bn_mean", _i), (_p + ".1.running_var", "post_bn_var", _i)]
WEIGHT_KEYS = [k for k, _, _ in _WEIGHT_MAP]


def _hparams() -> _cabi.HParams:
    return _cabi.HParams(**config.as_hparams())


def zero_beyond_lengths_(mel_coarse: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """Per-utterance stop rule: the rows of a batch keep decoding until the LAST one stops, so row b holds frames of a
    continued decode behind its own length.  The postnet's five k=5 layers see 10 frames to either side; with those frames
    zeroed, row b's ``mel_postnet[:length_b]`` is exactly what a solo run of utterance b (which ends at length_b, zero padding
    behind it) produces - together with ``Engine.postnet(coarse, lengths)``, which also zeroes every intermediate layer behind
    the row's length.  In place, asynchronous."""
    S = mel_coarse.shape[1]
    keep = torch.arange(S, device=mel_coarse.device)[None, :] < lengths.to(mel_coarse.device)[:, None]
    return mel_coarse.mul_(keep[:, :, None].to(mel_coarse.dtype))


class Engine:
    """One per model instance per device; not thread-safe."""

    def __init__(self, device: torch.device, weight_dtype: str = "bf16", decoder_path: str = "auto",
                 dense_path: str = "auto"):
        if weight_dtype not in WEIGHT_DTYPES:
            raise B200TacoError(f"weight_dtype must be one of {sorted(WEIGHT_DTYPES)}, got {weight_dtype!r}")
        if device.type != "cuda":
            raise B200TacoError(
                f"b200taco runs on CUDA devices only (got {device}); there is no CPU fallback")
        self.device = device
        self._lib = _cabi.lib()
        self._h = C.c_void_p()
        hp = _hparams()
        check(self._lib.b200taco_create(C.byref(hp), device.index if device.index is not None else torch.cuda.current_device(), C.byref(self._h)), "b200taco_create")
        self.weight_dtype = weight_dtype
        check(self._lib.b200taco_set_weight_dtype(self._h, WEIGHT_DTYPES[weight_dtype]), "b200taco_set_weight_dtype")
        self.decoder_path = decoder_path
        check(self._lib.b200taco_set_decoder_path(self._h, {"auto": 0, "simt": 1}[decoder_path]),
              "b200taco_set_decoder_path")
        self.dense_path = dense_path
        check(self._lib.b200taco_set_dense_path(self._h, {"auto": 0, "simt": 1, "tc": 2}[dense_path]), "b200taco_set_dense_path")
        self._packed_sig = None
        self._ws = {}            # one workspace per stage: stages of different batches overlap on different streams
        self.max_batch = self._lib.b200taco_max_batch()
        self.max_text_len = self._lib.b200taco_max_text_len()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.b200taco_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, B, T, S, stage="decoder"):
        need = int(self._lib.b200taco_workspace_bytes(self._h, B, T, S))
        ws = self._ws.get(stage)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[stage] = ws
        return C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel())

    def ensure_packed(self, tensors: dict) -> None:
        """(Re)pack when any parameter/buffer was replaced or modified in place."""
        sig = tuple((tensors[k].data_ptr(), tensors[k]._version) for k in WEIGHT_KEYS)
        if sig == self._packed_sig:
            return
        w = _cabi.Weights()
        keep = []
        for key, field, idx in _WEIGHT_MAP:
            t = tensors[key]
            if t.device != self.device or t.dtype != torch.float32:
                raise B200TacoError(f"{key}: expected float32 on {self.device}, got {t.dtype} on {t.device}")
            if not t.is_contiguous():
                t = t.contiguous()
                keep.append(t)
            if idx is None:
                setattr(w, field, t.data_ptr())
            else:
                getattr(w, field)[idx] = t.data_ptr()
        with torch.cuda.device(self.device):
            check(self._lib.b200taco_pack_weights(self._h, C.byref(w), self._stream()), "b200taco_pack_weights")
            if keep:
                torch.cuda.current_stream(self.device).synchronize()
        self._packed_sig = sig

    # ------------------------------------------------------------------ stages
    def encoder(self, ids: torch.Tensor, out: torch.Tensor = None, lengths: torch.Tensor = None) -> torch.Tensor:
        """``lengths`` (B,) : ragged batch - row b, positions [0, lengths[b]), is the encoder output of utterance b ALONE at its
        exact length (what the reference's one-utterance-at-a-time inference computes, inference.py:66); zero behind it."""
        B, T = ids.shape
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        memory = out if out is not None else torch.empty(B, T, config.encoder_embedding_dim, dtype=torch.float32,
                                                         device=self.device)
        assert memory.is_contiguous() and memory.shape == (B, T, config.encoder_embedding_dim)
        ws, wsz = self._workspace(B, T, 1, "encoder")
        with torch.cuda.device(self.device):
            if lengths is None:
                check(self._lib.b200taco_encoder_fwd(self._h, ids.data_ptr(), B, T, memory.data_ptr(),
                                                     ws, wsz, self._stream()), "b200taco_encoder_fwd")
            else:
                ln = lengths.to(device=self.device, dtype=torch.int32).contiguous()
                assert ln.shape == (B,)
                check(self._lib.b200taco_encoder_fwd_ragged(self._h, ids.data_ptr(), ln.data_ptr(), B, T, memory.data_ptr(),
                                                            ws, wsz, self._stream()), "b200taco_encoder_fwd_ragged")
        return memory

    def decoder_infer(self, memory, max_steps: int, gate_threshold: float, stop_mode: int, sync: bool = True):
        """Free-running decode.  ``sync=False`` returns the max-length buffers and the device
        tensor holding S without waiting for the kernel (benchmarks time it with CUDA events)."""
        memory = memory.contiguous()
        B, T, _ = memory.shape
        dev = self.device
        mel = torch.empty(B, max_steps, config.n_mels, dtype=torch.float32, device=dev)
        gate = torch.empty(B, max_steps, dtype=torch.float32, device=dev)
        align = torch.empty(max_steps, B, T, dtype=torch.float32, device=dev)
        lengths = torch.empty(B, dtype=torch.int32, device=dev)
        steps = torch.empty(1, dtype=torch.int32, device=dev)
        ws, wsz = self._workspace(B, T, max_steps)
        with torch.cuda.device(dev):
            check(self._lib.b200taco_decoder_infer(
                self._h, memory.data_ptr(), B, T, max_steps, float(gate_threshold), int(stop_mode),
                mel.data_ptr(), gate.data_ptr(), align.data_ptr(), lengths.data_ptr(), steps.data_ptr(),
                ws, wsz, self._stream()), "b200taco_decoder_infer")
        if not sync:
            return mel, gate, align, lengths, steps
        S = int(steps.item())            # the one host sync of an inference call
        return mel[:, :S], gate[:, :S], align[:S], lengths, S

    def stream_supported(self, rows: int) -> bool:
        """The utterance queue lives in the tensor-core decoder kernel (bf16 LSTM weights, >= 4 rows)."""
        return self.weight_dtype == "bf16" and self.decoder_path == "auto" and 4 <= rows <= self.max_batch

    def decoder_infer_stream(self, memory, rows: int, max_steps: int, gate_threshold: float, text_lengths=None):
        """Continuous batching (b200taco_decoder_infer_stream): the N utterances of ``memory`` (N,T,512), N a multiple of
        ``rows``, stream through ``rows`` decoder rows of one persistent kernel.  Returns, without synchronising,
        ``mel (N,cap,80), gate (N,cap), align (N/rows,cap,rows,T), len_h (N+1)``; frames beyond an utterance's length
        are zero.  ``len_h`` is PINNED HOST memory the kernel writes directly: ``len_h[u]`` becomes the length of
        utterance u (> 0) the moment it stops, ``len_h[N]`` the decode steps of the launch when it ends.
        ``text_lengths`` (N,) int64: ragged launch - ``memory`` rows are exact-length encoder outputs zero padded to T and every
        utterance attends to its own length only (b200taco_decoder_infer_stream_ragged)."""
        assert memory.is_contiguous()
        N, T, _ = memory.shape
        dev = self.device
        mel = torch.zeros(N, max_steps, config.n_mels, dtype=torch.float32, device=dev)
        gate = torch.zeros(N, max_steps, dtype=torch.float32, device=dev)
        align = torch.zeros(N // rows, max_steps, rows, T, dtype=torch.float32, device=dev)
        len_h = torch.zeros(N + 1, dtype=torch.int32).pin_memory()
        need = int(self._lib.b200taco_stream_workspace_bytes(self._h, N, rows, T))
        ws = self._ws.get("stream")
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._ws["stream"] = ws
        with torch.cuda.device(dev):
            if text_lengths is None:
                check(self._lib.b200taco_decoder_infer_stream(
                    self._h, memory.data_ptr(), N, rows, T, max_steps, float(gate_threshold),
                    mel.data_ptr(), gate.data_ptr(), align.data_ptr(), len_h.data_ptr(), len_h.data_ptr() + 4 * N,
                    C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), self._stream()), "b200taco_decoder_infer_stream")
            else:
                tl = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
                assert tl.shape == (N,)
                self._keep = tl                                   # alive until the next launch on this engine
                check(self._lib.b200taco_decoder_infer_stream_ragged(
                    self._h, memory.data_ptr(), tl.data_ptr(), N, rows, T, max_steps, float(gate_threshold),
                    mel.data_ptr(), gate.data_ptr(), align.data_ptr(), len_h.data_ptr(), len_h.data_ptr() + 4 * N,
                    C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), self._stream()), "b200taco_decoder_infer_stream_ragged")
        return mel, gate, align, len_h

    def decoder_teacher_forced(self, memory, mel_targets, text_lengths):
        memory = memory.contiguous()
        B, T, _ = memory.shape
        S = mel_targets.shape[2]
        dev = self.device
        mel_targets = mel_targets.to(device=dev, dtype=torch.float32).contiguous()
        tl = None
        if text_lengths is not None:
            tl = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
        mel = torch.empty(B, S, config.n_mels, dtype=torch.float32, device=dev)
        gate = torch.empty(B, S, dtype=torch.float32, device=dev)
        align = torch.empty(S, B, T, dtype=torch.float32, device=dev)
        ws, wsz = self._workspace(B, T, S)
        with torch.cuda.device(dev):
            check(self._lib.b200taco_decoder_teacher_forced(
                self._h, memory.data_ptr(), mel_targets.data_ptr(),
                tl.data_ptr() if tl is not None else None, B, T, S,
                mel.data_ptr(), gate.data_ptr(), align.data_ptr(), ws, wsz, self._stream()),
                "b200taco_decoder_teacher_forced")
        return mel, gate, align

    def postnet(self, mel_coarse: torch.Tensor, lengths=None) -> torch.Tensor:
        """mel_coarse (B,S,80) -> mel_postnet.  ``lengths`` (B,): per-utterance frame counts under the per-utterance stop rule -
        every layer then writes zeros behind an utterance's own length, so ``out[b, :lengths[b]]`` is exactly the solo result
        of utterance b (whose convolutions see zero padding behind its last frame); ``mel_coarse`` must be zero there."""
        mel_coarse = mel_coarse.contiguous()
        B, S, _ = mel_coarse.shape
        out = torch.empty_like(mel_coarse)
        ws, wsz = self._workspace(B, 1, S, "postnet")
        with torch.cuda.device(self.device):
            if lengths is None:
                check(self._lib.b200taco_postnet_fwd(self._h, mel_coarse.data_ptr(), B, S, out.data_ptr(),
                                                     ws, wsz, self._stream()), "b200taco_postnet_fwd")
            else:
                lengths = lengths.to(device=mel_coarse.device, dtype=torch.int32).contiguous()
                check(self._lib.b200taco_postnet_fwd_vocoder(self._h, mel_coarse.data_ptr(), B, S, lengths.data_ptr(), S,
                                                             out.data_ptr(), None, ws, wsz, self._stream()),
                      "b200taco_postnet_fwd_vocoder")
        return out

    def postnet_for_vocoder(self, mel_coarse: torch.Tensor, lengths=None, frames=None):
        """Postnet + vocoder hand-off in one pass: returns ``(mel_postnet (B,S,80), mel_vocoder (B,80,frames))``.  The second
        tensor is what inference.py:71 builds with ``.transpose(1, 2)`` for the vocoder, written channel major by the epilogue
        of the last postnet layer and zeroed from ``lengths[b]`` on (the per-utterance trim of train.py:71-79); ``frames``
        (default S) cuts the frame axis, e.g. to the longest utterance."""
        mel_coarse = mel_coarse.contiguous()
        B, S, _ = mel_coarse.shape
        frames = S if frames is None else int(frames)
        out = torch.empty_like(mel_coarse)
        voc = torch.empty(B, mel_coarse.shape[2], frames, dtype=torch.float32, device=mel_coarse.device)
        if lengths is not None:
            lengths = lengths.to(device=mel_coarse.device, dtype=torch.int32).contiguous()
        ws, wsz = self._workspace(B, 1, S, "postnet")
        with torch.cuda.device(self.device):
            check(self._lib.b200taco_postnet_fwd_vocoder(self._h, mel_coarse.data_ptr(), B, S,
                                                         lengths.data_ptr() if lengths is not None else None, frames,
                                                         out.data_ptr(), voc.data_ptr(), ws, wsz, self._stream()),
                  "b200taco_postnet_fwd_vocoder")
        return out, voc

    def vocoder_layout(self, mel: torch.Tensor, lengths=None, frames=None) -> torch.Tensor:
        """(B,S,80) -> (B,80,frames) channel major, zero from ``lengths[b]`` on (inference.py:71, train.py:71-79)."""
        mel = mel.contiguous()
        B, S, _ = mel.shape
        frames = S if frames is None else int(frames)
        voc = torch.empty(B, mel.shape[2], frames, dtype=torch.float32, device=mel.device)
        if lengths is not None:
            lengths = lengths.to(device=mel.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(self.device):
            check(self._lib.b200taco_vocoder_layout(mel.data_ptr(), lengths.data_ptr() if lengths is not None else None, B, S,
                                                    frames, voc.data_ptr(), self._stream()), "b200taco_vocoder_layout")
        return voc

    # ------------------------------------------------------------------ pipelined corpus synthesis
    def pipeline_infer(self, batches, max_steps: int, gate_threshold: float, stop_mode: int, to_host: bool = False,
                       pre_hook=None, timings=None):
        """Free-running inference of a sequence of id batches with the stages of consecutive batches overlapped.

        The persistent decoder kernel occupies 132 of the 148 SMs for tens of milliseconds; the dense kernels of the
        NEXT batch's encoder and of the PREVIOUS batch's postnet run meanwhile on the other SMs, on their own streams
        (the cluster BiLSTM needs whole GPCs and runs between two decoder launches).  ``batches``: iterable of (B,T)
        int64 id tensors, on the device or in (pinned) host memory.  Yields, in order, per batch
        ``(post, coarse, gate, align, lengths, S)``; with ``to_host`` the mel tensors and lengths are copied into pinned
        host buffers on the postnet stream.  Every result is complete when it is yielded.  ``pre_hook()`` runs on the
        decoder stream before each decoder launch (benchmarks flush L2 there); ``timings`` collects an (start, end) CUDA event
        pair around every decoder launch on the decoder stream.
        """
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_streams"):
            self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        sE, sD, sP = self._streams
        for s_ in (sE, sD, sP):
            s_.wait_stream(cur)
        it = iter(batches)

        def prep(ids):
            with torch.cuda.stream(sE):
                ids_dev = ids.to(device=dev, dtype=torch.int64, non_blocking=True)
                mem = self.encoder(ids_dev)
                ev = torch.cuda.Event()
                ev.record(sE)
            return mem, ev

        def launch_decoder(mem, ev_enc):
            with torch.cuda.stream(sD):
                sD.wait_event(ev_enc)
                mem.record_stream(sD)
                if pre_hook is not None:
                    pre_hook()                        # on the decoder stream: never while a decoder kernel is running
                if timings is not None:
                    t0 = torch.cuda.Event(enable_timing=True)
                    t0.record(sD)
                out = self.decoder_infer(mem, max_steps, gate_threshold, stop_mode, sync=False)
                if timings is not None:
                    t1 = torch.cuda.Event(enable_timing=True)
                    t1.record(sD)
                    timings.append((t0, t1))
                steps_h = torch.empty(1, dtype=torch.int32, pin_memory=True)
                steps_h.copy_(out[4], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(sD)
            return out, steps_h, ev

        def finish(out, steps_h, ev_dec):
            mel, gate, align, lengths, _ = out
            ev_dec.synchronize()                      # host learns S; the next decoder launch is already queued
            S = int(steps_h.item())
            with torch.cuda.stream(sP):
                sP.wait_event(ev_dec)
                for t in (mel, gate, align, lengths):
                    t.record_stream(sP)
                coarse = mel[:, :S].contiguous()
                if stop_mode == STOP_PER_UTTERANCE:
                    zero_beyond_lengths_(coarse, lengths)     # every row as if it had been decoded alone (see there)
                post = self.postnet(coarse, lengths if stop_mode == STOP_PER_UTTERANCE else None)
                res = (post, coarse, gate[:, :S], align[:S], lengths, S)
                for t_ in (post, coarse):                 # allocated on this side stream, consumed on the caller's stream
                    t_.record_stream(cur)
                if to_host:
                    post_h = torch.empty(post.shape, dtype=post.dtype, pin_memory=True)
                    len_h = torch.empty(lengths.shape, dtype=lengths.dtype, pin_memory=True)
                    post_h.copy_(post, non_blocking=True)
                    len_h.copy_(lengths, non_blocking=True)
                    res = res + (post_h, len_h)
                ev = torch.cuda.Event()
                ev.record(sP)
            return res, ev

        try:
            first = next(it)
        except StopIteration:
            return
        mem, ev_enc = prep(first)
        pending = launch_decoder(mem, ev_enc)
        done = None
        for ids in it:
            mem, ev_enc = prep(ids)                   # encoder(i+1) overlaps decoder(i)
            nxt = launch_decoder(mem, ev_enc)         # queued behind decoder(i) on the decoder stream
            res, ev_post = finish(*pending)           # postnet(i) overlaps decoder(i+1)
            if done is not None:
                done[1].synchronize()
                yield done[0]
            done, pending = (res, ev_post), nxt
        res, ev_post = finish(*pending)
        if done is not None:
            done[1].synchronize()
            yield done[0]
        ev_post.synchronize()
        cur.wait_stream(sP)
        yield res

    def stream_infer_ragged(self, utterances, max_steps: int, gate_threshold: float, rows: int = 64, launch_utterances: int = 2048,
                            length_ratio: float = 1.35, to_host: bool = False):
        """Corpus synthesis with continuous batching over utterances of DIFFERENT text lengths (BASELINE config 4).

        ``utterances``: sequence of 1-D int64 id tensors (host).  The reference's inference takes no text lengths (F8), so
        every utterance is ENCODED over its own positions only (ragged encoder batches of neighbours of the length-sorted
        queue, ``Engine.encoder(lengths=...)``: bit for bit its exact-length solo encoding) and its encoder output is placed in
        a row padded to the launch's longest utterance; the decoder launch masks every utterance to its own length, which
        reproduces its solo exact-length decode.  The launch plan (length-sorted corpus, at most ``launch_utterances``
        utterances and a length spread of at most ``length_ratio`` per launch, longest text first inside a launch, filled up to
        a multiple of ``rows``) is ``shard.plan_ragged_launches``.  Yields ``(index, mel_postnet (S,80), mel_coarse (S,80),
        gate (S,), alignments (S,T_i))`` per utterance, device tensors (pinned host tensors with ``to_host``), in completion
        order of the launches."""
        import time
        from . import shard
        dev = self.device
        if not self.stream_supported(rows):
            raise B200TacoError("stream_infer_ragged needs 4 <= rows <= 64 and the default (bf16, tensor-core) decoder")
        lens = [int(u.shape[0]) for u in utterances]
        self.last_stream_steps = 0
        for members, n_real in shard.plan_ragged_launches(lens, rows, launch_utterances, length_ratio):
            N, Tmax = len(members), max(lens[i] for i in members)
            mem = torch.zeros(N, Tmax, config.encoder_embedding_dim, dtype=torch.float32, device=dev)
            # ragged encoder batches: max_batch neighbours of the length-sorted queue at a time, padded to the longest of them, every
            # utterance encoded over its own positions only (Engine.encoder(lengths=...)) - one call per 64 utterances instead of
            # one per group of EQUAL length (8192 utterances of 201 different lengths were 804 small encoder calls)
            for lo in range(0, n_real, self.max_batch):
                part = members[lo:lo + self.max_batch]
                Tc = max(lens[i] for i in part)
                ids = torch.nn.utils.rnn.pad_sequence([utterances[i] for i in part], batch_first=True)
                assert ids.shape == (len(part), Tc)
                ln = torch.tensor([lens[i] for i in part], dtype=torch.int32)
                mem[lo:lo + len(part), :Tc] = self.encoder(ids.to(dev, non_blocking=True), lengths=ln.to(dev, non_blocking=True))
            if N > n_real:
                mem[n_real:] = mem[torch.tensor([n_real - 1 - k % n_real for k in range(N - n_real)], device=dev)]
            tl = torch.tensor([lens[i] for i in members], dtype=torch.int64)
            mel, gate, align, len_h = self.decoder_infer_stream(mem, rows, max_steps, gate_threshold, text_lengths=tl)
            done_ev = torch.cuda.Event()
            done_ev.record()
            for c in range(N // rows):                               # postnet of a chunk as soon as its utterances have stopped
                lo, hi = c * rows, (c + 1) * rows
                while int(len_h[lo:hi].min()) <= 0:
                    if done_ev.query():
                        done_ev.synchronize()
                        if int(len_h[lo:hi].min()) <= 0:
                            raise B200TacoError("decoder launch ended with unfinished utterances")
                        break
                    time.sleep(2e-5)
                S = int(len_h[lo:hi].max())
                if not hasattr(self, "_streams"):
                    self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
                sP = self._streams[2]
                with torch.cuda.stream(sP):
                    coarse = mel[lo:hi, :S].contiguous()              # zero behind every utterance's own length
                    post = self.postnet(coarse, len_h[lo:hi].to(dev, non_blocking=True))
                    if to_host:
                        post = post.to("cpu", non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(sP)
                ev.synchronize()
                for r in range(rows):
                    slot = lo + r
                    if slot >= n_real:
                        break
                    i, n = members[slot], int(len_h[slot])
                    yield i, post[r, :n], coarse[r, :n], gate[slot, :n], align[c, :n, r, :lens[i]]
            done_ev.synchronize()
            self.last_stream_steps += int(len_h[-1])

    def stream_infer(self, batches, max_steps: int, gate_threshold: float, to_host: bool = False, group_batches: int = 32,
                     pre_hook=None, timings=None, live_rows=None):
        """Corpus synthesis with continuous batching: like ``pipeline_infer`` under the per-utterance stop rule, but the
        utterances of up to ``group_batches`` consecutive batches (same shape (rows, T)) share ONE persistent decoder launch
        whose rows are refilled from the following batches as their utterances stop (``decoder_infer_stream``), so no
        row idles behind an early stop.  The kernel publishes every utterance's length in pinned host memory the moment it
        stops; the host polls it and runs the postnet (and the copies to the host) of a batch on another stream as soon as
        its last utterance is done, while the launch is still decoding the following batches.  Encoders of the next group
        overlap the decoder launch of the current one.  Yields per batch, in order, what ``pipeline_infer`` yields; frames
        of a row beyond its own length are zero (``pipeline_infer``: that row's continued decode).
        ``to_host``: True copies mel_postnet (+ lengths) of every batch into pinned host memory; ``"all"`` also mel_coarse, the gate
        logits and the alignments (the reference's whole 4-tuple), appended to the yielded tuple in that order.
        ``live_rows``: decoder rows one launch keeps busy - a multiple of the batch size, at most 64 (None: the batch size).
        A decode step of 64 rows costs ~1.4x a step of 32 (it is latency, not bandwidth), so two 32-utterance batches decoding
        side by side finish ~40 % sooner; results per utterance do not depend on it."""
        import collections
        import time
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_streams"):
            self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        sE, sD, sP = self._streams
        for s_ in (sE, sD, sP):
            s_.wait_stream(cur)
        batches = list(batches)
        if not batches:
            return
        rows, T = batches[0].shape
        if any(tuple(b.shape) != (rows, T) for b in batches) or not self.stream_supported(rows):
            raise B200TacoError("stream_infer needs batches of one shape (rows, T) with 4 <= rows <= 64 and the default "
                                "(bf16, tensor-core) decoder; use pipeline_infer otherwise")
        # launches of (nearly) equal size, an even number of batches each where possible (two batches decode side by side):
        # the drain at the end of a launch costs the same whatever its length, so neither a short last launch nor an odd one
        n_groups = (len(batches) + group_batches - 1) // group_batches
        per = (len(batches) + n_groups - 1) // n_groups
        per += per % 2 if per < len(batches) else 0
        groups = [batches[i:i + per] for i in range(0, len(batches), per)]
        self.last_stream_steps = 0                    # decode steps of all launches of this call

        def prep(group):
            with torch.cuda.stream(sE):
                mem = torch.empty(len(group) * rows, T, config.encoder_embedding_dim, dtype=torch.float32, device=dev)
                ids_dev = torch.empty(len(group) * rows, T, dtype=torch.int64, device=dev)
                for i, ids in enumerate(group):
                    ids_dev[i * rows:(i + 1) * rows].copy_(ids, non_blocking=True)
                # rows are independent in the encoder: as many utterances per launch as the library carries
                per = max(rows, self.max_batch // rows * rows)
                for lo in range(0, len(group) * rows, per):
                    self.encoder(ids_dev[lo:lo + per], out=mem[lo:lo + per])
                ev = torch.cuda.Event()
                ev.record(sE)
            return mem, ev

        def launch_decoder(mem, ev_enc):
            with torch.cuda.stream(sD):
                sD.wait_event(ev_enc)
                mem.record_stream(sD)
                if pre_hook is not None:
                    pre_hook()
                if timings is not None:
                    t0 = torch.cuda.Event(enable_timing=True)
                    t0.record(sD)
                n_b = mem.shape[0] // rows
                live = rows
                if live_rows is not None:                    # the largest multiple of the batch size that divides the launch
                    for m_ in range(max(1, int(live_rows) // rows), 0, -1):
                        if n_b % m_ == 0 and self.stream_supported(rows * m_):
                            live = rows * m_
                            break
                out = self.decoder_infer_stream(mem, live, max_steps, gate_threshold) + (live,)
                if timings is not None:
                    t1 = torch.cuda.Event(enable_timing=True)
                    t1.record(sD)
                    timings.append((t0, t1))
                ev = torch.cuda.Event()
                ev.record(sD)
            return out, ev

        def finish(out, ev_dec):
            """Generator over the batches of one launch, each post-processed as soon as its utterances have stopped."""
            mel, gate, align, len_h, live = out
            for t in (mel, gate, align):
                t.record_stream(sP)
            inflight = collections.deque()
            for i in range(mel.shape[0] // rows):
                lo, hi = i * rows, (i + 1) * rows
                while int(len_h[lo:hi].min()) <= 0:          # pinned memory written by the running kernel
                    if ev_dec.query():                        # launch over (or failed): the lengths are final
                        ev_dec.synchronize()
                        if int(len_h[lo:hi].min()) <= 0:
                            raise B200TacoError("decoder launch ended with unfinished utterances")
                        break
                    time.sleep(2e-5)
                S = int(len_h[lo:hi].max())
                with torch.cuda.stream(sP):
                    # no wait on the decoder stream: the kernel published these lengths after the frames themselves
                    coarse = mel[lo:hi, :S].contiguous()
                    lengths = len_h[lo:hi].to(dev, non_blocking=True)
                    post = self.postnet(coarse, lengths)
                    res = (post, coarse, gate[lo:hi, :S], align[lo // live, :S, lo % live:lo % live + rows], lengths, S)
                    for t_ in (post, coarse, lengths):        # allocated on this side stream, consumed on the caller's stream
                        t_.record_stream(cur)
                    if to_host:
                        post_h = torch.empty(post.shape, dtype=post.dtype, pin_memory=True)
                        post_h.copy_(post, non_blocking=True)
                        res = res + (post_h, len_h[lo:hi])
                        if to_host == "all":      # the reference's whole 4-tuple: also mel_coarse, gate logits, alignments
                            extra = []
                            for t_ in (coarse, res[2], res[3]):
                                h_ = torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True)
                                h_.copy_(t_, non_blocking=True)
                                extra.append(h_)
                            res = res + tuple(extra)
                    ev = torch.cuda.Event()
                    ev.record(sP)
                inflight.append((res, ev))
                while inflight and inflight[0][1].query():
                    yield inflight.popleft()[0]
            ev_dec.synchronize()
            self.last_stream_steps += int(len_h[-1])
            for res, ev in inflight:
                ev.synchronize()
                yield res

        launched = []                                 # every launch writes into pinned host memory it was handed: never
        try:                                          # let go of those buffers while a launch may still be running
            mem, ev_enc = prep(groups[0])
            pending = launch_decoder(mem, ev_enc)
            launched.append(pending)
            for group in groups[1:]:
                mem, ev_enc = prep(group)             # encoders of group g+1 overlap the decoder of group g
                nxt = launch_decoder(mem, ev_enc)     # queued behind the launch of group g
                launched.append(nxt)
                yield from finish(*pending)
                pending = nxt
            yield from finish(*pending)
            cur.wait_stream(sP)
        finally:                                      # also when the consumer abandons the generator or a stage raised
            for _, ev in launched:
                ev.synchronize()

    # ------------------------------------------------------------------ measurement aids
    def measure_l2_read_gbps(self, nbytes: int = 36 << 20, iters: int = 50) -> float:
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        out = C.c_float(0.0)
        with torch.cuda.device(self.device):
            check(self._lib.b200taco_measure_l2_read(buf.data_ptr(), nbytes, iters, C.byref(out),
                                                     self._stream()), "b200taco_measure_l2_read")
        return float(out.value)

    def decoder_profile(self) -> dict:
        """Phase timers of the last decoder launch (ns summed over steps) and the step count."""
        out = (C.c_int64 * 21)()
        with torch.cuda.device(self.device):
            check(self._lib.b200taco_decoder_profile(self._h, out, 21, self._stream()), "b200taco_decoder_profile")
        names = ("attention_lstm", "attention", "decoder_lstm", "projection_prenet")
        steps = max(1, int(out[4]))
        prof = {"steps": int(out[4]), "us_per_step": {n: out[i] / steps / 1e3 for i, n in enumerate(names)}}
        if any(out[5:]):
            prof["fine_us_per_step"] = [round(out[i] / steps / 1e3, 2) for i in range(5, 21)]
        return prof

    def launch_count(self, reset: bool = False) -> int:
        return int(self._lib.b200taco_launch_count(1 if reset else 0))
