"""Hyper-parameters of the Tacotron 2 hot path.

Mirrors the model-facing constants of the reference's ``src/config.py:14-49`` (same
names, same values) so that code written against ``config.X`` keeps working.  Only the
constants the encoder/decoder/postnet path reads are kept; audio/text front-end and
training-schedule constants are out of scope.

The CUDA library is a compiled specialisation of exactly these sizes
(``csrc/dims.cuh``); :func:`as_hparams` is what the host passes to
``b200taco_create`` which rejects anything else.
"""

# ARPAbet phonemes with stress digits, alphabetical, then space and two punctuation
# marks: 72 symbols, identical in content and order to src/config.py:14-23.
_VOWELS = ("AA AE AH AO AW AY EH ER EY IH IY OW OY UH UW").split()
_CONSONANTS = ("B CH D DH F G HH JH K L M N NG P R S SH T TH V W Y Z ZH").split()
SYMBOLS = sorted([v + s for v in _VOWELS for s in "012"] + _CONSONANTS) + [" ", ".", ","]
assert len(SYMBOLS) == 72

# Encoder (src/config.py:26-31)
symbols_embedding_dim = 512
encoder_n_convolutions = 3
encoder_embedding_dim = 512
encoder_kernel_size = 5

# Decoder (src/config.py:34-40)
n_mels = 80
decoder_rnn_dim = 1024
prenet_dim = 256
max_decoder_steps = 1000
gate_threshold = 0.5
p_attention_dropout = 0.1
p_decoder_dropout = 0.1

# Attention (src/config.py:43-48)
attention_rnn_dim = 1024
attention_dim = 128
location_n_filters = 32
location_kernel_size = 31

# Guided-attention schedule of the loss (src/config.py:51-54), read by b200taco.loss
attention_initial_sigma_factor = 0.05
attention_sigma_warmup_steps = 4000
attention_min_sigma = 1.0
attention_max_sigma_cap = 20.0

# Post-net (src/config.py:57-59)
postnet_embedding_dim = 512
postnet_kernel_size = 5
postnet_n_convolutions = 5


def as_hparams() -> dict:
    """The hparams block handed to ``b200taco_create`` (include/b200taco.h)."""
    return dict(
        n_symbols=len(SYMBOLS),
        embed_dim=symbols_embedding_dim,
        enc_convs=encoder_n_convolutions,
        enc_kernel=encoder_kernel_size,
        n_mels=n_mels,
        rnn_dim=decoder_rnn_dim,
        prenet_dim=prenet_dim,
        attn_rnn_dim=attention_rnn_dim,
        attn_dim=attention_dim,
        loc_filters=location_n_filters,
        loc_kernel=location_kernel_size,
        postnet_dim=postnet_embedding_dim,
        postnet_kernel=postnet_kernel_size,
        postnet_convs=postnet_n_convolutions,
        max_decoder_steps=max_decoder_steps,
    )
