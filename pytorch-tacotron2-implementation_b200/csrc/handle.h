// Internal definition of b200taco_handle: the packed-weight cache and device facts.
#pragma once

#include "common.cuh"

struct b200taco_handle {
    int device = 0;
    int num_sms = 0;
    int weight_dtype = 0;           // 0 = bf16 LSTM matrices (default), 1 = fp32
    int decoder_path = 0;           // 0 = auto (tcgen05 when batch >= 4 and bf16 weights), 1 = SIMT only
    int dense_path = 0;             // 0 = tcgen05 3xTF32 dense layers (dense_tc.cu), 1 = fp32 SIMT (dense.cu)
    bool packed = false;
    const unsigned long long* last_profile = nullptr;   // device: phase timers of the last decoder launch
    const unsigned long long* last_trace = nullptr;     // device: per-CTA event times of one step (resident kernel, TACO_RES_TRACE=step)

    // one device allocation holding every packed tensor below
    void* arena = nullptr;
    size_t arena_bytes = 0;

    // ---- decoder (decoder.cu) ----
    void* w_attn = nullptr;         // [128 cta][K_ATTN][32 rows]  bf16 or fp32, rows = unit*4+gate
    void* w_dec = nullptr;          // [128 cta][K_DEC ][32 rows]
    void* w_attn_tc = nullptr;      // bf16 mode: [128 cta][K/64 chunks][32 rows x 64 k UMMA K-major tile]
    void* w_dec_tc = nullptr;
    float* bias_attn = nullptr;     // [128][32]  b_ih + b_hh
    float* bias_dec = nullptr;
    float* wq_slices = nullptr;     // [128 cta][128 a][8 u]    query_layer.weight[a][8*cta+u]
    float* wproj_slices = nullptr;  // [128 cta][81 o][8 u]     rows 0..79 linear_projection, row 80 gate_layer (dh part)
    float* wproj_ctxT = nullptr;    // [512 j][96]              ctx part, transposed, zero padded
    float* proj_bias = nullptr;     // [96]
    float* prenet1T = nullptr;      // [80][256]
    float* prenet2T = nullptr;      // [256][256]
    float* loc_conv = nullptr;      // [32][2][31]
    float* loc_denseT = nullptr;    // [32 f][128 a]
    float* attn_v = nullptr;        // [128] v.weight, [128] v.bias, [129] energy_scale
    // ---- dense layers (dense.cu): weights as [tap][ci][co], folded BatchNorm scale/shift ----
    float* enc_emb = nullptr;       // [72][512]
    float* enc_conv_w[3] = {};      // [5][512][512]
    float* enc_conv_scale[3] = {};  // [512]
    float* enc_conv_shift[3] = {};
    float* enc_lstm_wih = nullptr;  // [512][2048]  (fwd i,f,g,o | bwd i,f,g,o)
    float* enc_lstm_bias = nullptr; // [2048]       b_ih + b_hh
    float* enc_lstm_whh = nullptr;  // [2 dir][8 rank][128 rows][256 k], rows = unit*4+gate
    float* attn_memory_wT = nullptr;// [512][128]
    float* post_conv_w[5] = {};     // [5][ci][co]
    float* post_conv_scale[5] = {};
    float* post_conv_shift[5] = {};
    // ---- the same dense weights pre-split into TF32 big/small planes, UMMA K-major tiles (dense_tc.cu) ----
    float* enc_conv_wtc[3] = {};
    float* enc_lstm_wih_tc = nullptr;
    float* attn_memory_wtc = nullptr;
    float* prenet1_tc = nullptr;
    float* prenet2_tc = nullptr;
    float* post_conv_wtc[5] = {};
};

namespace b200taco {

// pack.cu
int pack_weights(b200taco_handle* h, const b200taco_weights* w, cudaStream_t stream);
size_t packed_arena_bytes(int weight_dtype);

// dense.cu : Y[n][co] = act((sum_{tap,ci} X[n+tap-pad][ci] * W[tap][ci][co]) * scale[co] + shift[co]) (+ res[n][co])
enum Activation { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };
int conv1d_channels_last(const float* X, int ldx, const float* W, const float* scale, const float* shift,
                         const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len,
                         int c_in, int c_out, int taps, int act, cudaStream_t stream);
int embedding_gather(const int64_t* ids, const float* table, float* out, int n, cudaStream_t stream);
int transpose_bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t stream);

// dense_tc.cu : the same contract on the tensor cores (3xTF32); dense_layer picks the path of the handle
bool dense_tc_supported(int c_in, int c_out, int ldx, int ldy);
size_t dense_tc_weight_floats(int taps, int c_in, int c_out);
int pack_dense_tc(float* dst, const float* w_tap_ci_co, int taps, int c_in, int c_out, cudaStream_t stream);
constexpr int TRAIN_CONV_TC_MIN_ROWS = 1024;    // training convolutions with at least this many rows run on the tensor cores
// per-utterance options of a layer over B utterances of seg_len rows each.
//  y != nullptr: second copy of the result, channel major (B, c_out, ld) and zero from lengths[b] on: the vocoder hand-off
//  layout (inference.py:71, train.py:71-79).
//  zero_tail: the result rows of utterance b from lengths[b] on are written as ZERO - a convolution stack then sees exactly
//  what it sees when utterance b is run alone at its own length (zero padding behind the last frame), layer by layer.
struct ChannelMajorOut {
    float* y = nullptr;
    const int32_t* lengths = nullptr;
    int ld = 0;
    bool zero_tail = false;
};
int dense_layer(const b200taco_handle* h, const float* X, int ldx, const float* W, const float* Wtc, const float* scale,
                const float* shift, const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len, int c_in,
                int c_out, int taps, int act, bool precise, cudaStream_t stream, const ChannelMajorOut& cm = ChannelMajorOut());

// data.cu : batch assembly (src/data_utils.py:45-84) and the vocoder hand-off layout as stand-alone passes
int vocoder_layout(const float* mel, int ldm, const int32_t* lengths, int B, int S, int C, int S_out, float* out, cudaStream_t stream);
int zero_tail_rows(float* Y, int ldy, const int32_t* lengths, int B, int S, int C, cudaStream_t stream);

// bilstm.cu
int bilstm_recurrence(const b200taco_handle* h, const float* xproj, float* memory, int B, int T, cudaStream_t stream,
                      const int32_t* lengths = nullptr);

// decoder.cu
struct DecoderArgs {
    int B, T, max_steps, mode, stop_mode;
    float gate_threshold;
    const float* memory;          // (B,T,512)
    const float* pm;              // (B,T,128) processed memory
    const int64_t* text_lengths;  // (B) or null
    const void* p_all;            // teacher forcing: prenet outputs packed by decoder_pack_p_all, else null
    float* mel;                   // (B,max_steps,80)
    float* gate;                  // (B,max_steps)
    float* align;                 // (max_steps,B,T)
    int32_t* lengths;             // (B)
    int32_t* steps_run;           // (1)
    void* scratch;                // decoder_scratch_bytes()
    size_t scratch_bytes;
    // utterance queue (continuous batching; 0 = off): n_utts utterances of text length T stream through the B decoder
    // rows; memory / pm / mel / gate / lengths are then per utterance and align is (n_utts / B, max_steps, B, T)
    int n_utts = 0;
    // training (teacher-forced mode only; all null / 0 otherwise): per-step state for the backward pass, dropout masks
    float *sv_ga = nullptr, *sv_ca = nullptr, *sv_ah = nullptr, *sv_gd = nullptr, *sv_cd = nullptr, *sv_dh = nullptr,
          *sv_ctx = nullptr;
    const float *mask_a = nullptr, *mask_d = nullptr;
    float drop_scale = 1.f;
};
size_t decoder_scratch_bytes(int B, int T);
int decoder_batch_pad(int B);
bool decoder_uses_tensor_cores(const b200taco_handle* h, int B);
int decoder_pack_p_all(const b200taco_handle* h, const float* Y, void* P, int B, int S, cudaStream_t stream);
int decoder_run(const b200taco_handle* h, const DecoderArgs& a, cudaStream_t stream);

// decoder_res.cu: weight-resident, barrier-free kernel for one or two free-running utterances (bf16 weights)
bool decoder_res_supported(const b200taco_handle* h, const DecoderArgs& a);
size_t decoder_res_scratch_bytes(int B, int T);
int decoder_res_run(const b200taco_handle* h, const DecoderArgs& a, void* scratch, cudaStream_t stream);

// api.cu: layout of the decoder's training "saved" buffer (float offsets)
struct DecSaved { size_t frames, y1, y2, pm, ga, ca, ah, gd, cd, dh, ctx, total; };
DecSaved dec_saved(int B, int T, int S);

// measure.cu
int measure_l2_read(void* d_buf, size_t bytes, int iters, float* gbps, cudaStream_t stream);
int measure_sync_costs(long long* out8_host, cudaStream_t stream);
int measure_stream_probe(void* d_buf, size_t per_cta, int ctas, int iters, int mode, int chunk, int depth, float* gbps_per_sm,
                         cudaStream_t stream);

// train_decoder.cu
// What follows the attention backward in the reverse walk, fused into its tail when `Wq` is set: the query layer's data gradient
// d ah_t += dpq_t . Wq (src/model.py:123) and the pointwise backward of the attention-LSTM cell (both CTAs of the pair hold the
// summed dpq; CTA r takes the units [512 r, 512 r + 512) of its utterance, one per thread) - two kernels less in the chain.
struct AttnCellTail {
    const float* Wq;            // (128, 1024) query_layer.weight, null = not fused
    float* d_ah;                // [B][1024] carried gradient of the attention hidden state (consumed: zeroed)
    const float *gates, *c, *c_prev, *mask;
    float scale;
    float *dc, *dz;
    __nv_bfloat16* dzp;
    // second cell, optional (gates2 != null): the DECODER-LSTM cell backward of the PREVIOUS decode step t - 1.  Its input d dh_{t-1}
    // was completed by the data-gradient GEMM that precedes this kernel, so it needs nothing this kernel computes; it rides along
    // to take one more kernel out of the chain.  dh2 [B][1024] carried (consumed: zeroed), dh2_extra [B][ld2] the projection's share.
    float* dh2;
    const float* dh2_extra;
    int ld2;
    const float *gates2, *c2, *c_prev2, *mask2;
    float *dc2, *dz2;
    __nv_bfloat16* dzp2;
};

}  // namespace b200taco
