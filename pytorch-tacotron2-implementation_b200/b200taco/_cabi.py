"""ctypes binding of include/b200taco.h -- the only place the shared library is loaded.

There is deliberately no fallback: if ``lib/libb200taco.so`` is missing or fails to load,
:func:`lib` raises and every model call fails loudly (north_star: "no CPU fallback").
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``python pytorch-tacotron2-implementation_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("B200TACO_LIB") or os.path.join(_PKG_ROOT, "lib", "libb200taco.so")

ABI_VERSION = 1

fp = C.POINTER(C.c_float)


class HParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_symbols", "embed_dim", "enc_convs", "enc_kernel",
        "n_mels", "rnn_dim", "prenet_dim", "attn_rnn_dim", "attn_dim",
        "loc_filters", "loc_kernel",
        "postnet_dim", "postnet_kernel", "postnet_convs",
        "max_decoder_steps")]


class Weights(C.Structure):
    """Field order and arity must match ``b200taco_weights`` in include/b200taco.h."""
    _fields_ = [
        ("enc_embedding", C.c_void_p),
        ("enc_conv_w", C.c_void_p * 3), ("enc_conv_b", C.c_void_p * 3),
        ("enc_bn_w", C.c_void_p * 3), ("enc_bn_b", C.c_void_p * 3),
        ("enc_bn_mean", C.c_void_p * 3), ("enc_bn_var", C.c_void_p * 3),
        ("enc_lstm_w_ih", C.c_void_p * 2), ("enc_lstm_w_hh", C.c_void_p * 2),
        ("enc_lstm_b_ih", C.c_void_p * 2), ("enc_lstm_b_hh", C.c_void_p * 2),
        ("prenet_w", C.c_void_p * 2),
        ("attn_query_w", C.c_void_p), ("attn_memory_w", C.c_void_p),
        ("attn_loc_conv_w", C.c_void_p), ("attn_loc_dense_w", C.c_void_p),
        ("attn_v_w", C.c_void_p), ("attn_v_b", C.c_void_p), ("attn_energy_scale", C.c_void_p),
        ("attn_lstm_w_ih", C.c_void_p), ("attn_lstm_w_hh", C.c_void_p),
        ("attn_lstm_b_ih", C.c_void_p), ("attn_lstm_b_hh", C.c_void_p),
        ("dec_lstm_w_ih", C.c_void_p), ("dec_lstm_w_hh", C.c_void_p),
        ("dec_lstm_b_ih", C.c_void_p), ("dec_lstm_b_hh", C.c_void_p),
        ("proj_w", C.c_void_p), ("proj_b", C.c_void_p),
        ("gate_w", C.c_void_p), ("gate_b", C.c_void_p),
        ("post_conv_w", C.c_void_p * 5), ("post_conv_b", C.c_void_p * 5),
        ("post_bn_w", C.c_void_p * 5), ("post_bn_b", C.c_void_p * 5),
        ("post_bn_mean", C.c_void_p * 5), ("post_bn_var", C.c_void_p * 5),
    ]


# name -> (restype, argtypes); this table is also what tests check against the header.
SIGNATURES = {
    "b200taco_abi_version": (C.c_int32, []),
    "b200taco_last_error": (C.c_char_p, []),
    "b200taco_max_batch": (C.c_int32, []),
    "b200taco_max_text_len": (C.c_int32, []),
    "b200taco_create": (C.c_int32, [C.POINTER(HParams), C.c_int32, C.POINTER(C.c_void_p)]),
    "b200taco_destroy": (C.c_int32, [C.c_void_p]),
    "b200taco_set_weight_dtype": (C.c_int32, [C.c_void_p, C.c_int32]),
    "b200taco_set_decoder_path": (C.c_int32, [C.c_void_p, C.c_int32]),
    "b200taco_set_dense_path": (C.c_int32, [C.c_void_p, C.c_int32]),
    "b200taco_pack_weights": (C.c_int32, [C.c_void_p, C.POINTER(Weights), C.c_void_p]),
    "b200taco_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "b200taco_encoder_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_encoder_fwd_ragged": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_decoder_infer": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_float, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_stream_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "b200taco_decoder_infer_stream": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                  C.c_int32, C.c_int32, C.c_float,
                                                  C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_decoder_infer_stream_ragged": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                         C.c_int32, C.c_int32, C.c_float,
                                                         C.c_void_p, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_decoder_teacher_forced": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_int32, C.c_int32, C.c_int32,
                                                    C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_postnet_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_postnet_fwd_vocoder": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_vocoder_layout": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "b200taco_collate": (C.c_int32, [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200taco_guided_attention_target": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                     C.c_void_p, C.c_void_p]),
    "b200taco_loss_scratch_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "b200taco_loss_fwd": (C.c_int32, [C.c_void_p] * 9 + [C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200taco_loss_bwd": (C.c_int32, [C.c_void_p] * 9 + [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7),
    "b200taco_gate_targets": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "b200taco_measure_l2_read": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_int32,
                                             C.POINTER(C.c_float), C.c_void_p]),
    "b200taco_measure_stream": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.POINTER(C.c_float), C.c_void_p]),
    "b200taco_measure_sync_costs": (C.c_int32, [C.POINTER(C.c_int64), C.c_void_p]),
    "b200taco_decoder_profile": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p]),
    "b200taco_decoder_trace": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p]),
    "b200taco_attn_bwd_phase_ns": (C.c_int32, [C.POINTER(C.c_uint64), C.c_int32]),
    "b200taco_debug_umma_gemm": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "b200taco_launch_count": (C.c_int64, [C.c_int32]),
    # ---- training (stage by stage, include/b200taco.h)
    "b200taco_postnet_train_saved_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b200taco_postnet_train_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b200taco_postnet_train_fwd": (C.c_int32, [C.c_void_p, C.POINTER(Weights), C.c_void_p, C.c_int32, C.c_int32,
                                               C.POINTER(C.c_void_p), C.c_float, C.c_void_p, C.c_void_p, C.c_size_t,
                                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_postnet_train_bwd": (C.c_int32, [C.c_void_p, C.POINTER(Weights), C.c_void_p, C.c_int32, C.c_int32,
                                               C.POINTER(C.c_void_p), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(Weights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_encoder_train_saved_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b200taco_encoder_train_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b200taco_encoder_train_fwd": (C.c_int32, [C.c_void_p, C.POINTER(Weights), C.c_void_p, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_encoder_train_bwd": (C.c_int32, [C.c_void_p, C.POINTER(Weights), C.c_void_p, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.POINTER(Weights), C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
    "b200taco_decoder_train_saved_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "b200taco_decoder_train_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "b200taco_decoder_train_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_float,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200taco_decoder_train_bwd": (C.c_int32, [C.c_void_p, C.POINTER(Weights), C.c_void_p, C.c_void_p,
                                               C.c_int32, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_float,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(Weights), C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


class B200TacoError(RuntimeError):
    """Raised for every non-zero status returned by the library."""


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise B200TacoError(
                    f"{LIB_PATH} not found: the CUDA library is not built. There is no "
                    "CPU/PyTorch fallback; run `python pytorch-tacotron2-implementation_b200/build.py`.")
            handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)          # AttributeError = missing export
                fn.restype, fn.argtypes = res, args
            got = handle.b200taco_abi_version()
            if got != ABI_VERSION:
                raise B200TacoError(f"ABI version mismatch: library {got}, binding {ABI_VERSION}")
            _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().b200taco_last_error()
        raise B200TacoError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")
