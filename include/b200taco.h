/*
 * b200taco.h -- C ABI of the B200-native Tacotron 2 hot path.
 *
 * The reference (Angad8285/pytorch-tacotron2-implementation) has no FFI: its boundary for
 * this path is the Python surface of `Tacotron2` in src/model.py.  This header is the
 * native boundary the drop-in `b200taco.model.Tacotron2` binds (ctypes, see
 * INTEGRATION.md); every entry point names the reference method it replaces.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch / C++ types.
 *  - every pointer named d_* is DEVICE memory owned by the caller (PyTorch's caching
 *    allocator); the library owns only the handle and its packed-weight cache.
 *  - `stream` is a cudaStream_t passed as void* (the caller's current stream).
 *  - all functions return 0 on success or a negative b200taco_status; the message is
 *    available from b200taco_last_error() (thread-local).  Nothing throws.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails
 *    with B200TACO_ERR_CUDA.
 *  - fp32 tensors are dense row-major with the shapes given in the comments.
 */
#ifndef B200TACO_H
#define B200TACO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200TACO_API __attribute__((visibility("default")))
#else
#define B200TACO_API
#endif

#define B200TACO_ABI_VERSION 1

typedef enum {
    B200TACO_OK = 0,
    B200TACO_ERR_INVALID = -1,     /* bad argument / shape */
    B200TACO_ERR_HPARAMS = -2,     /* hparams differ from the compiled specialisation */
    B200TACO_ERR_UNSUPPORTED = -3, /* size beyond a compiled limit (B, T) */
    B200TACO_ERR_CUDA = -4,        /* CUDA runtime error, message has the detail */
    B200TACO_ERR_STATE = -5,       /* weights not packed, workspace too small, ... */
    B200TACO_ERR_TIMEOUT = -6      /* a device-side wait gave up (kernel bug guard) */
} b200taco_status;

/* Mirrors the model constants of the reference's src/config.py:26-59. */
typedef struct {
    int32_t n_symbols, embed_dim, enc_convs, enc_kernel;
    int32_t n_mels, rnn_dim, prenet_dim, attn_rnn_dim, attn_dim;
    int32_t loc_filters, loc_kernel;
    int32_t postnet_dim, postnet_kernel, postnet_convs;
    int32_t max_decoder_steps;
} b200taco_hparams;

/*
 * Device pointers to the fp32 parameters/buffers, named after the reference state_dict
 * (SURVEY.md section 8b).  All are read-only for the library.
 */
typedef struct {
    /* Encoder, src/model.py:19-48 */
    const float* enc_embedding;            /* (72, 512) */
    const float* enc_conv_w[3];            /* (512, 512, 5) */
    const float* enc_conv_b[3];            /* (512) */
    const float* enc_bn_w[3];              /* BatchNorm1d weight/bias/running stats (512) */
    const float* enc_bn_b[3];
    const float* enc_bn_mean[3];
    const float* enc_bn_var[3];
    const float* enc_lstm_w_ih[2];         /* [0]=forward, [1]=reverse: (1024, 512) */
    const float* enc_lstm_w_hh[2];         /* (1024, 256) */
    const float* enc_lstm_b_ih[2];         /* (1024) */
    const float* enc_lstm_b_hh[2];
    /* Decoder, src/model.py:186-217 */
    const float* prenet_w[2];              /* (256, 80), (256, 256), no bias */
    const float* attn_query_w;             /* (128, 1024) */
    const float* attn_memory_w;            /* (128, 512) */
    const float* attn_loc_conv_w;          /* (32, 2, 31) */
    const float* attn_loc_dense_w;         /* (128, 32) */
    const float* attn_v_w;                 /* (1, 128) */
    const float* attn_v_b;                 /* (1) */
    const float* attn_energy_scale;        /* () */
    const float* attn_lstm_w_ih;           /* (4096, 768) */
    const float* attn_lstm_w_hh;           /* (4096, 1024) */
    const float* attn_lstm_b_ih;           /* (4096) */
    const float* attn_lstm_b_hh;
    const float* dec_lstm_w_ih;            /* (4096, 1536) */
    const float* dec_lstm_w_hh;            /* (4096, 1024) */
    const float* dec_lstm_b_ih;
    const float* dec_lstm_b_hh;
    const float* proj_w;                   /* (80, 1536) */
    const float* proj_b;                   /* (80) */
    const float* gate_w;                   /* (1, 1536) */
    const float* gate_b;                   /* (1) */
    /* PostNet, src/model.py:347-390 */
    const float* post_conv_w[5];           /* (512,80,5) (512,512,5)x3 (80,512,5) */
    const float* post_conv_b[5];
    const float* post_bn_w[5];
    const float* post_bn_b[5];
    const float* post_bn_mean[5];
    const float* post_bn_var[5];
} b200taco_weights;

typedef struct b200taco_handle b200taco_handle;

/* Stop rule of the free-running decoder. */
#define B200TACO_STOP_BATCH_WIDE 0    /* reference: Decoder.inference, src/model.py:327 */
#define B200TACO_STOP_PER_UTTERANCE 1 /* caller-side trim rule train.py:71-79, done on device */

B200TACO_API int32_t b200taco_abi_version(void);
B200TACO_API const char* b200taco_last_error(void);

/* Limits of the compiled specialisation (max batch, max encoder length). */
B200TACO_API int32_t b200taco_max_batch(void);
B200TACO_API int32_t b200taco_max_text_len(void);

/* Replaces: Tacotron2.__init__ reading src/config.py (src/model.py:415-422). */
B200TACO_API int32_t b200taco_create(const b200taco_hparams* hp, int32_t device, b200taco_handle** out);
B200TACO_API int32_t b200taco_destroy(b200taco_handle* h);

/*
 * Storage type of the four decoder LSTM matrices inside the packed cache (98 % of the bytes a
 * decoder step reads).  BF16 (default) halves the L2 traffic; FP32 keeps the reference's
 * values bit for bit.  Accumulation and all state are fp32 either way.  Takes effect at the
 * next b200taco_pack_weights.  Replaces: nothing in the reference (fp32 nn.LSTMCell).
 */
#define B200TACO_WEIGHTS_BF16 0
#define B200TACO_WEIGHTS_FP32 1
B200TACO_API int32_t b200taco_set_weight_dtype(b200taco_handle* h, int32_t dtype);

/*
 * Kernel selection for the LSTM gate contractions of the decoder.  AUTO (default): tcgen05 tensor-core
 * tiles (bf16 operands, hi/lo split activations, fp32 accumulators in TMEM) when the batch is large
 * at least 4 (measured crossover: the batch tile of the tensor-core kernel is 32 rows whatever B) and the weights are stored bf16; the vectorised fp32 SIMT
 * contraction otherwise.  SIMT forces the latter (A/B measurements, bit-for-bit fp32 activations).
 */
#define B200TACO_PATH_AUTO 0
#define B200TACO_PATH_SIMT 1
B200TACO_API int32_t b200taco_set_decoder_path(b200taco_handle* h, int32_t path);

/*
 * Kernel selection for the dense layers (encoder convolutions, BiLSTM input projection, memory layer,
 * teacher-forced prenet, postnet).  The tensor-core kernel is a tcgen05 implicit GEMM with 3xTF32 operand
 * splitting and fp32 accumulators in TMEM; the tensor core truncates when it accumulates, so its results sit
 * ~1e-6 relative from round-to-nearest fp32 (measured).
 *   AUTO (default): postnet and teacher-forced prenet on the tensor cores; the encoder and the memory layer,
 *                   whose outputs are re-read by every one of up to 1000 recurrent decoder steps, on the fp32
 *                   CUDA-core kernel.
 *   TC            : every dense layer on the tensor cores (encoder with 8 rotating accumulators).
 *   SIMT          : every dense layer on the CUDA cores.
 * Replaces: nn.Conv1d / nn.Linear dispatch (src/model.py:25-39, :112, :166-171, :347-407).
 */
#define B200TACO_PATH_TC 2
B200TACO_API int32_t b200taco_set_dense_path(b200taco_handle* h, int32_t path);

/*
 * Repack the fp32 parameters into the kernels' layouts (bf16 fragment order for the
 * two decoder LSTMs, hi/lo bf16 planes for the GEMM operands, folded BatchNorm).  Call
 * after load_state_dict / optimizer.step (the host tracks param._version).
 * Replaces: nothing in the reference (PyTorch modules read their own parameters).
 */
B200TACO_API int32_t b200taco_pack_weights(b200taco_handle* h, const b200taco_weights* w, void* stream);

/* Bytes of caller-provided scratch needed for a (B, T, S) problem; S = decoder steps. */
B200TACO_API size_t b200taco_workspace_bytes(const b200taco_handle* h, int32_t B, int32_t T, int32_t S);

/*
 * Encoder.forward, src/model.py:50-77 (eval-mode BatchNorm).
 * d_ids (B,T) int64 -> d_memory (B,T,512).
 */
B200TACO_API int32_t b200taco_encoder_fwd(b200taco_handle* h, const int64_t* d_ids, int32_t B, int32_t T,
                             float* d_memory, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * The same for a batch of utterances of DIFFERENT lengths padded to T: row b of d_memory, positions [0, d_lengths[b]), is what
 * Encoder.forward returns for that utterance ALONE at its exact length - the reference's inference encodes one unpadded
 * utterance at a time (inference.py:66, F8: no text lengths, no packing) - and is zero behind it.  The convolutions see zeros
 * behind an utterance's length at every layer, the reverse LSTM direction starts at its last position.
 * d_ids (B,T) int64 (anything behind the lengths), d_lengths (B) int32 in [1, T].
 */
B200TACO_API int32_t b200taco_encoder_fwd_ragged(b200taco_handle* h, const int64_t* d_ids, const int32_t* d_lengths, int32_t B,
                                                 int32_t T, float* d_memory, void* d_workspace, size_t workspace_bytes,
                                                 void* stream);

/*
 * Decoder.inference, src/model.py:299-333, as one persistent kernel: the discarded first
 * step, the frame loop and the stop rule all run on the device.
 *   d_memory (B,T,512); max_steps = min(max_decoder_steps, max_len_cap)
 *   d_mel (B,max_steps,80)  d_gate (B,max_steps)  d_align (max_steps,B,T)
 *   d_lengths (B) int32: per-utterance length under the chosen stop rule
 *   d_steps_run (1) int32: S, number of output frames (same for all rows)
 * Rows/steps beyond S are left untouched.
 */
B200TACO_API int32_t b200taco_decoder_infer(b200taco_handle* h, const float* d_memory, int32_t B, int32_t T,
                               int32_t max_steps, float gate_threshold, int32_t stop_mode,
                               float* d_mel, float* d_gate, float* d_align,
                               int32_t* d_lengths, int32_t* d_steps_run,
                               void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * Continuous batching of Decoder.inference (src/model.py:299-333) under the per-utterance stop rule
 * (train.py:71-79): N utterances of EQUAL text length T stream through `rows` decoder rows of ONE persistent
 * kernel.  A row whose utterance stopped (its own gate after >= 2 frames, or max_steps frames) takes the next
 * utterance from the queue at the following decode step with freshly initialised state
 * (Decoder._initialize_decoder_states, src/model.py:219-234, incl. the discarded first step, :310-316), so every
 * utterance is decoded exactly as b200taco_decoder_infer(stop_mode = PER_UTTERANCE) decodes it in a batch of
 * `rows`, without the idle row-steps behind early stops.  Tensor-core decoder kernel only: 4 <= rows <= 64,
 * bf16 LSTM weights, N a multiple of rows.
 *   d_memory (N,T,512)
 *   d_mel (N,max_steps,80)  d_gate (N,max_steps)  d_align (N/rows,max_steps,rows,T): utterance u is row u % rows of
 *   block u / rows.  Frames >= d_lengths[u] of utterance u are NOT written (zero-fill the buffers beforehand).
 *   d_lengths (N) int32; d_steps_run (1) int32: decode steps the launch took.
 *   d_lengths[u] is stored (system-scope release) the moment utterance u stops, after all of its frames: it may
 *   point to zero-initialised MAPPED PINNED HOST memory, which the host polls to post-process finished utterances
 *   (postnet, copies) on another stream while the launch is still decoding the rest.
 * Workspace: b200taco_stream_workspace_bytes(h, N, rows, T).
 */
B200TACO_API size_t b200taco_stream_workspace_bytes(const b200taco_handle* h, int32_t N, int32_t rows, int32_t T);
B200TACO_API int32_t b200taco_decoder_infer_stream(b200taco_handle* h, const float* d_memory, int32_t N, int32_t rows,
                                      int32_t T, int32_t max_steps, float gate_threshold,
                                      float* d_mel, float* d_gate, float* d_align,
                                      int32_t* d_lengths, int32_t* d_steps_run,
                                      void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * Ragged continuous batching: as b200taco_decoder_infer_stream, for utterances of DIFFERENT text lengths in one launch.
 * d_memory (N,T,512) holds every utterance's exact-length encoder output (the reference's Tacotron2.inference ignores
 * text_lengths, src/model.py:482-494, so an utterance must be encoded and decoded at its own length) in rows padded to the common
 * T with finite values (zeros); d_text_lengths (N) int64 are the true lengths.  Every utterance attends to its own positions only
 * (masked_fill of src/model.py:131-132 per utterance): its frames, gate logits and alignments [:, :length] are those of its solo
 * exact-length decode; alignment columns behind its length are 0.
 */
B200TACO_API int32_t b200taco_decoder_infer_stream_ragged(b200taco_handle* h, const float* d_memory, const int64_t* d_text_lengths,
                                                          int32_t N, int32_t rows, int32_t T, int32_t max_steps, float gate_threshold,
                                                          float* d_mel, float* d_gate, float* d_align, int32_t* d_lengths,
                                                          int32_t* d_steps_run, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * Decoder.forward (teacher forcing), src/model.py:277-297, eval-mode dropout.
 *   d_mel_targets (B,80,S) as the reference receives them; d_text_lengths (B) int64 or
 *   NULL (= full length): positions >= length are masked with -1e9 (src/model.py:132).
 *   d_mel (B,S,80)  d_gate (B,S)  d_align (S,B,T)
 */
B200TACO_API int32_t b200taco_decoder_teacher_forced(b200taco_handle* h, const float* d_memory,
                                        const float* d_mel_targets, const int64_t* d_text_lengths,
                                        int32_t B, int32_t T, int32_t S,
                                        float* d_mel, float* d_gate, float* d_align,
                                        void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * PostNet.forward + the residual add of Tacotron2.forward/inference
 * (src/model.py:392-407, :476-478, :498-500), eval mode.
 *   d_mel_coarse (B,S,80) -> d_mel_post (B,S,80) = coarse + postnet(coarse)
 */
B200TACO_API int32_t b200taco_postnet_fwd(b200taco_handle* h, const float* d_mel_coarse, int32_t B, int32_t S,
                             float* d_mel_post, void* d_workspace, size_t workspace_bytes, void* stream);

/* Measurement and bring-up aids (L2 probes, phase timers, stage traces, the tcgen05 unit-test GEMM) are declared in
 * b200taco_debug.h: they are exported by the same library but are NOT part of the stable product ABI. */

/*
 * Guided-attention target of the reference's loss (Tacotron2Loss.create_diagonal_attention_target, train.py:148-171), whose
 * Python loop over B x S dominates a training step: target[b][t][p] = N(p; min(int(t L_b / S), L_b - 1), sigma_b) over
 * p < L_b, normalised by (sum + 1e-8), zero for p >= L_b.
 *   d_text_lengths (B) int64, d_sigma (B) float (the annealed width the caller computed, train.py:151-156),
 *   d_target (B,S,Tmax) float.  No handle: the kernel reads no model state.
 */
B200TACO_API int32_t b200taco_guided_attention_target(const int64_t* d_text_lengths, const float* d_sigma, int32_t B, int32_t S,
                                         int32_t Tmax, float* d_target, void* stream);

/*
 * The whole Tacotron2Loss.forward of the reference (train.py:173-217) on the device, every scalar left in device memory
 * (the reference reads entropy and sigma back with .item() every step, train.py:151,157,208):
 *   masked L1 of mel_coarse and mel_post (B,S,80) against mel_target (B,80,S) over the frames s < mel_lengths[b]
 *   (train.py:176-192), BCE-with-logits of gate (B,S) against gate_target (B,S) (train.py:194), and - when d_align is not
 *   NULL - KL(guided target || clamp_min(attention, 1e-8)) / B / S clamped to 150 and the mean attention entropy
 *   (train.py:197-204) over align (S,B,T) with the guided target of b200taco_guided_attention_target recomputed on the fly,
 *   then the entropy-adaptive weight (train.py:205-210: w_start if entropy > entropy_target else max(w_min, w_start *
 *   entropy / entropy_target); without attention term the weight stays w_current) and the total (train.py:215).
 *   d_out8: [mel, gate, kl, entropy, weight, total, kl before the clamp, number of valid mel bins].
 *   d_scratch: b200taco_loss_scratch_bytes(B, S) bytes.  Three launches (two without attention term), fixed summation order.
 * b200taco_loss_bwd: gradients with respect to mel_post, mel_coarse, gate and (d_g_align not NULL) the alignments for the
 *   upstream gradient d_grad_out8 of the 8 results (entries 0, 1, 2, 5 are used; the weight is a constant as in the reference,
 *   where it passes through .item()).
 * b200taco_gate_targets: the stop-token targets of the training loop (train.py:505-508), out (B,S) = 1 from frame
 *   mel_lengths[b] - 1 on, else 0.
 */
B200TACO_API int64_t b200taco_loss_scratch_bytes(int32_t B, int32_t S);
B200TACO_API int32_t b200taco_loss_fwd(const float* d_mel_post, const float* d_mel_coarse, const float* d_mel_target,
                                       const float* d_gate, const float* d_gate_target, const int64_t* d_mel_lengths,
                                       const float* d_align, const int64_t* d_text_lengths, const float* d_sigma, int32_t B,
                                       int32_t S, int32_t T, float w_start, float w_min, float entropy_target, float w_current,
                                       float* d_scratch, float* d_out8, void* stream);
B200TACO_API int32_t b200taco_loss_bwd(const float* d_mel_post, const float* d_mel_coarse, const float* d_mel_target,
                                       const float* d_gate, const float* d_gate_target, const int64_t* d_mel_lengths,
                                       const float* d_align, const int64_t* d_text_lengths, const float* d_sigma, int32_t B,
                                       int32_t S, int32_t T, const float* d_out8, const float* d_grad_out8, float* d_g_post,
                                       float* d_g_coarse, float* d_g_gate, float* d_g_align, void* stream);
B200TACO_API int32_t b200taco_gate_targets(const int64_t* d_mel_lengths, int32_t B, int32_t S, float* d_out, void* stream);

/*
 * Vocoder hand-off.  The reference hands mel_outputs_postnet.transpose(1, 2), i.e. (B, 80, S), to the
 * vocoder (inference.py:71) and trims every utterance to its own stop frame on the host (train.py:71-79).
 * b200taco_postnet_fwd_vocoder = b200taco_postnet_fwd that ALSO writes the result channel major into d_mel_vocoder
 * (B, 80, S_out), S_out <= S, from the epilogue of the last postnet layer (d_mel_vocoder NULL: no second copy).
 * d_lengths (B) int32, may be NULL: utterance b has d_lengths[b] frames.  With lengths EVERY postnet layer writes zeros
 * behind an utterance's own length, so each utterance's frames [0, length) are exactly those of a solo run at its own length
 * (where the convolutions see zero padding behind the last frame) although the batch is as long as its longest member; the
 * frames of d_mel_post and d_mel_vocoder behind the length are zero.  (d_mel_coarse is expected zero there as well: the
 * residual is added as is.)  b200taco_vocoder_layout writes the layout for an existing (B, S, 80) tensor.
 */
B200TACO_API int32_t b200taco_postnet_fwd_vocoder(b200taco_handle* h, const float* d_mel_coarse, int32_t B, int32_t S,
                                                  const int32_t* d_lengths, int32_t S_out, float* d_mel_post, float* d_mel_vocoder,
                                                  void* d_workspace, size_t workspace_bytes, void* stream);
B200TACO_API int32_t b200taco_vocoder_layout(const float* d_mel, const int32_t* d_lengths, int32_t B, int32_t S, int32_t S_out,
                                             float* d_out, void* stream);

/*
 * Batch assembly on the device (TextMelCollate.__call__, src/data_utils.py:45-84): the B ragged samples of a
 * batch arrive as two flat arrays - text ids (sum of L_i int64) and mels (sample i stored (80, S_i) row major, sum of 80 S_i
 * floats) - with offsets (B + 1 entries; mel offsets count FRAMES).  Row i of text_padded (B, Lmax) / mel_padded (B, 80, Smax)
 * receives sample d_order[i] (the descending text-length order of src/data_utils.py:54-56), zero padded.
 */
B200TACO_API int32_t b200taco_collate(const int64_t* d_text_flat, const int64_t* d_text_offsets, const float* d_mel_flat,
                                      const int64_t* d_mel_offsets, const int32_t* d_order, int32_t B, int32_t Lmax, int32_t Smax,
                                      int64_t* d_text_padded, float* d_mel_padded, void* stream);

/* Number of kernels this library launched on this thread since the last reset. */
B200TACO_API int64_t b200taco_launch_count(int32_t reset);

/* ------------------------------------------------------------------------------------------------
 * Training (work in progress: first correct fp32 versions, stage by stage).
 *
 * Forward passes keep what the backward needs in a caller-owned "saved" buffer; the raw fp32 parameters
 * (b200taco_weights, as in b200taco_pack_weights) are read directly.  Backward passes write gradients in
 * the reference's parameter layouts into the buffers named by `grads`, a b200taco_weights-shaped struct
 * whose pointers are the destinations (only the fields of the stage are touched).  Dropout masks are
 * 0/1 float tensors drawn by the host RNG, laid out like the layer outputs; drop_scale = 1/(1-p).
 * BatchNorm uses batch statistics over all rows (padding included, as the reference does); the batch mean and
 * biased variance of layer i are left in the saved buffer for the host's running-statistics update.
 *
 * Postnet: PostNet.forward in train mode + residual, src/model.py:392-407, :476-477.
 * saved layout per layer: z [B*S][c_out], y [B*S][c_out], mean [c_out], var [c_out], rstd [c_out].
 */
B200TACO_API size_t b200taco_postnet_train_saved_bytes(int32_t B, int32_t S);
B200TACO_API size_t b200taco_postnet_train_scratch_bytes(int32_t B, int32_t S);
B200TACO_API int32_t b200taco_postnet_train_fwd(b200taco_handle* h, const b200taco_weights* w, const float* d_coarse,
                                                int32_t B, int32_t S, const float* const* d_masks, float drop_scale,
                                                float* d_post, float* d_saved, size_t saved_bytes, float* d_scratch,
                                                size_t scratch_bytes, void* stream);
B200TACO_API int32_t b200taco_postnet_train_bwd(b200taco_handle* h, const b200taco_weights* w, const float* d_coarse,
                                                int32_t B, int32_t S, const float* const* d_masks, float drop_scale,
                                                const float* d_saved, const float* d_dpost, float* d_dcoarse,
                                                const b200taco_weights* grads, float* d_scratch, size_t scratch_bytes,
                                                void* stream);


/*
 * Encoder in train mode: Encoder.forward with batch-statistics BatchNorm, src/model.py:50-77, and its backward
 * (BiLSTM backward through time, convolution stack, embedding).  saved layout: x0 [B*T][512] | 3 x (z, y, mean, var,
 * rstd) | xproj [B*T][2048] | per (b, t, direction): i, f, g, o, c, h (6 x 256).  Gradients of the encoder fields of
 * `grads` are overwritten.
 */
B200TACO_API size_t b200taco_encoder_train_saved_bytes(int32_t B, int32_t T);
B200TACO_API size_t b200taco_encoder_train_scratch_bytes(int32_t B, int32_t T);
B200TACO_API int32_t b200taco_encoder_train_fwd(b200taco_handle* h, const b200taco_weights* w, const int64_t* d_ids,
                                                int32_t B, int32_t T, float* d_memory, float* d_saved, size_t saved_bytes,
                                                float* d_scratch, size_t scratch_bytes, void* stream);
B200TACO_API int32_t b200taco_encoder_train_bwd(b200taco_handle* h, const b200taco_weights* w, const int64_t* d_ids,
                                                int32_t B, int32_t T, const float* d_saved, const float* d_dmemory,
                                                const b200taco_weights* grads, float* d_scratch, size_t scratch_bytes,
                                                void* stream);


/*
 * Teacher-forced decoder in train mode: Decoder.forward, src/model.py:277-297, with dropout on both prenet layers
 * (masks [S*B][256], rows ordered [step][b]) and on the carried attention / decoder LSTM hidden states (masks
 * [S][B][1024]); the persistent decoder kernel runs the forward and leaves per-step gates, cell / hidden states and
 * contexts in `d_saved`.  Needs b200taco_pack_weights with the current parameters first.  The backward walks the steps
 * in reverse (LSTM cells, attention incl. location conv and cumulative weights, prenet, memory layer) and overwrites
 * the decoder fields of `grads` and d_dmemory (B,T,512).  d_dalign (S,B,T) may be null.
 */
B200TACO_API size_t b200taco_decoder_train_saved_bytes(int32_t B, int32_t T, int32_t S);
B200TACO_API size_t b200taco_decoder_train_scratch_bytes(int32_t B, int32_t T, int32_t S);
B200TACO_API int32_t b200taco_decoder_train_fwd(b200taco_handle* h, const float* d_memory, const float* d_mel_targets,
                                                const int64_t* d_text_lengths, int32_t B, int32_t T, int32_t S,
                                                const float* d_mask_p1, const float* d_mask_p2, float prenet_scale,
                                                const float* d_mask_a, const float* d_mask_d, float lstm_scale,
                                                float* d_mel, float* d_gate, float* d_align, float* d_saved,
                                                size_t saved_bytes, void* d_workspace, size_t workspace_bytes, void* stream);
B200TACO_API int32_t b200taco_decoder_train_bwd(b200taco_handle* h, const b200taco_weights* w, const float* d_memory,
                                                const int64_t* d_text_lengths, int32_t B, int32_t T, int32_t S,
                                                const float* d_mask_p1, const float* d_mask_p2, float prenet_scale,
                                                const float* d_mask_a, const float* d_mask_d, float lstm_scale,
                                                const float* d_saved, const float* d_align, const float* d_dcoarse,
                                                const float* d_dgate, const float* d_dalign, float* d_dmemory,
                                                const b200taco_weights* grads, float* d_scratch, size_t scratch_bytes,
                                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TACO_H */
