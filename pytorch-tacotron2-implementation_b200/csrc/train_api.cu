// Training-mode entry points of the C ABI (include/b200taco.h, "training" section): forward passes that keep what
// the backward needs in a caller-owned "saved" buffer, and the hand-written backward passes.  The raw fp32
// parameters are read directly (they change every optimiser step); gradients are written in the reference's
// parameter layouts so that the host can hand them to the optimiser unchanged.
#include "handle.h"

namespace b200taco {

// train_dense.cu
int train_add(float* dst, const float* a, const float* b, size_t n, cudaStream_t st);
int conv_w_to_packed(float* dst, const float* w, int c_out, int c_in, int taps, cudaStream_t st);
int conv_w_from_packed(float* w, const float* src, int c_out, int c_in, int taps, cudaStream_t st);
size_t convbn_saved_floats(int n_rows, int c_out);
int convbn_train_fwd(const float* X, int ldx, const float* Wp, const float* conv_b, const float* gamma, const float* beta,
                     const float* mask, float drop_scale, int act, float* saved, int n_rows, int seg_len, int c_in, int c_out,
                     int taps, float* wtc, cudaStream_t st);
int convbn_train_bwd(const float* dY, const float* X, int ldx, const float* Wp, const float* gamma, const float* beta,
                     const float* mask, float drop_scale, int act, const float* saved, float* dX, int lddx, float* dWp,
                     float* db, float* dgamma, float* dbeta, float* dz, float* wflip, int n_rows, int seg_len, int c_in,
                     int c_out, int taps, float* wtc, cudaStream_t st);

int train_colsum(float* dst, const float* src, int n_rows, int C, cudaStream_t st);
int dense_wgrad(const float* X, int ldx, const float* dY, int ldy, float* dW, int n_rows, int seg_len, int c_in, int c_out,
                int taps, cudaStream_t st);
// train_lstm.cu
constexpr int LSTM_SAVE_F = 6 * ENC_H;
int transpose2d(float* dst, const float* src, int R, int Cc, cudaStream_t st);
int bilstm_train_fwd(const float* whhT_f, const float* whhT_b, const float* xproj, float* out, float* saved, int B, int T,
                     cudaStream_t st);
int bilstm_train_bwd(const float* whh_f, const float* whh_b, const float* d_out, const float* saved, float* dz, float* hprev,
                     int B, int T, cudaStream_t st);
int embedding_bwd(const int64_t* ids, const float* dx, float* dE, int n, cudaStream_t st);

// train_decoder.cu
int lstm_cell_bwd(const float* gates, const float* c, const float* c_prev, float* dh_post, const float* dh_extra, int ld_extra, const float* mask, float scale,
                  float* dc, float* dz, void* dz_tiles, int B, cudaStream_t st);
int attn_cumsum(const float* align, float* cum, int S, int BT, cudaStream_t st);
int train_axpy(float* dst, int ld_dst, const float* a, int ld_a, int n, int cols, bool accumulate, cudaStream_t st);
int relu_drop_bwd(const float* dy, const float* y, const float* mask, float scale, float* dx, size_t n, cudaStream_t st);
size_t attention_bwd_scratch_floats(int B, int T);
int attention_bwd(const float* w, const float* prev, const float* cum, const float* pq, const float* pm, const float* memory,
                  float* dctx, const float* dw_ext, const float* dprev_next, float* gsum, const int64_t* lengths,
                  const float* Wloc, const float* Wld, const float* v, const float* vb, const float* escale, float* dmemory,
                  float* dpm, float* dpq, float* dprev_out, float* pacc, float* scratch, float* dctx_out,
                  const float* dctx_extra, int ld_extra, const AttnCellTail* tail, int B, int T, cudaStream_t st);
bool attention_bwd_defers_dmemory(int T);
bool attention_bwd_fuses_cell(int T);
int attn_bwd_phase_ns(unsigned long long* out8, int reset);
int skinny_gemm_multi(const float* X, int ldx, const float* W, int ldw, int B, int K, int nseg, float* const* outs, const int* ldos,
                      const int* n_ends, const bool* acc, cudaStream_t st);
int ctx_dmemory(const float* w, const float* dctx_all, float* dmemory, int S, int B, int T, cudaStream_t st);
constexpr int ATTN_PACC_F = ATT_DIM * LOC_F + LOC_F * 2 * LOC_K + ATT_DIM + 64;
size_t skinny_tc_packed_bytes(int K, int N);
bool skinny_tc_supported(int B, int K, int N, int nseg, const int* n_ends);
int pack_skinny_tc(void* dst, const float* W, int ldw, int K, int N, cudaStream_t st);
size_t skinny_tc_act_bytes(int K);
int skinny_gemm_pair_tc(const void* Xp0, const void* Wp0, int nseg0, float* const* outs0, const int* ldos0, const int* n_ends0, const bool* acc0,
                        const void* Xp1, const void* Wp1, int nseg1, float* const* outs1, const int* ldos1, const int* n_ends1, const bool* acc1,
                        int B, int K, cudaStream_t st);
int skinny_gemm_multi_tc(const void* Xp, const void* Wp, int B, int K, int nseg, float* const* outs, const int* ldos,
                         const int* n_ends, const bool* acc, cudaStream_t st);
int skinny_gemm(const float* X, int ldx, const float* W, int ldw, float* out, int ldo, int B, int K, int N, bool accumulate,
                cudaStream_t st);
int transpose_block(float* dst, int ldd, const float* src, int lds, int R, int C, cudaStream_t st);
int assemble_dout(const float* d_coarse, const float* d_gate, float* dout, int B, int S, cudaStream_t st);
int assemble_wpg(const float* wp, const float* wg, float* dst, cudaStream_t st);

namespace {

struct PostDims { int c_in[POST_CONVS], c_out[POST_CONVS]; };
PostDims post_dims() {
    PostDims d{};
    for (int i = 0; i < POST_CONVS; ++i) {
        d.c_in[i] = i == 0 ? N_MELS : POST_DIM;
        d.c_out[i] = i == POST_CONVS - 1 ? N_MELS : POST_DIM;
    }
    return d;
}
size_t post_wp_floats() { return (size_t)POST_KERNEL * POST_DIM * POST_DIM; }     // largest packed weight

int check_train_handle(const b200taco_handle* h) {
    TACO_REQUIRE(h != nullptr, B200TACO_ERR_INVALID, "null handle");
    TACO_CUDA(cudaSetDevice(h->device));
    return B200TACO_OK;
}

}  // namespace
}  // namespace b200taco

using namespace b200taco;

extern "C" {

// floats of the saved buffer: per layer z, y, mean, var, rstd
B200TACO_API size_t b200taco_postnet_train_saved_bytes(int32_t B, int32_t S) {
    const PostDims d = post_dims();
    size_t f = 0;
    for (int i = 0; i < POST_CONVS; ++i) f += convbn_saved_floats(B * S, d.c_out[i]);
    return f * sizeof(float);
}
B200TACO_API size_t b200taco_postnet_train_scratch_bytes(int32_t B, int32_t S) {
    // packed weights, flipped weights, packed dW, the tensor-core weight planes (2x); dz and two d-activation buffers
    return (5 * post_wp_floats() + (size_t)3 * B * S * POST_DIM) * sizeof(float);
}

B200TACO_API int32_t b200taco_postnet_train_fwd(b200taco_handle* h, const b200taco_weights* w, const float* d_coarse, int32_t B,
                                                int32_t S, const float* const* d_masks, float drop_scale, float* d_post,
                                                float* d_saved, size_t saved_bytes, float* d_scratch, size_t scratch_bytes,
                                                void* stream) {
    if (int e = check_train_handle(h)) return e;
    TACO_REQUIRE(w && d_coarse && d_post && d_saved && d_scratch && B >= 1 && S >= 1, B200TACO_ERR_INVALID,
                 "postnet_train_fwd: bad argument");
    TACO_REQUIRE(saved_bytes >= b200taco_postnet_train_saved_bytes(B, S) && scratch_bytes >= b200taco_postnet_train_scratch_bytes(B, S),
                 B200TACO_ERR_STATE, "postnet_train_fwd: buffers too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PostDims d = post_dims();
    const int n = B * S;
    const float* x = d_coarse;
    float* saved = d_saved;
    float* wp = d_scratch;
    float* wtc = d_scratch + 3 * post_wp_floats();
    for (int i = 0; i < POST_CONVS; ++i) {
        if (int e = conv_w_to_packed(wp, w->post_conv_w[i], d.c_out[i], d.c_in[i], POST_KERNEL, st)) return e;
        const int act = i == POST_CONVS - 1 ? ACT_NONE : ACT_TANH;
        if (int e = convbn_train_fwd(x, d.c_in[i], wp, w->post_conv_b[i], w->post_bn_w[i], w->post_bn_b[i],
                                     d_masks ? d_masks[i] : nullptr, drop_scale, act, saved, n, S, d.c_in[i], d.c_out[i],
                                     POST_KERNEL, wtc, st))
            return e;
        x = saved + (size_t)n * d.c_out[i];                     // y of this layer
        saved += convbn_saved_floats(n, d.c_out[i]);
    }
    return train_add(d_post, d_coarse, x, (size_t)n * N_MELS, st);          // residual, src/model.py:476-477
}

// grads: b200taco_weights-shaped struct whose post_* pointers receive the gradients (reference layouts)
B200TACO_API int32_t b200taco_postnet_train_bwd(b200taco_handle* h, const b200taco_weights* w, const float* d_coarse, int32_t B,
                                                int32_t S, const float* const* d_masks, float drop_scale, const float* d_saved,
                                                const float* d_dpost, float* d_dcoarse, const b200taco_weights* grads,
                                                float* d_scratch, size_t scratch_bytes, void* stream) {
    if (int e = check_train_handle(h)) return e;
    TACO_REQUIRE(w && d_coarse && d_saved && d_dpost && d_dcoarse && grads && d_scratch, B200TACO_ERR_INVALID,
                 "postnet_train_bwd: bad argument");
    TACO_REQUIRE(scratch_bytes >= b200taco_postnet_train_scratch_bytes(B, S), B200TACO_ERR_STATE, "postnet_train_bwd: scratch too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PostDims d = post_dims();
    const int n = B * S;
    float* wp = d_scratch;
    float* wflip = wp + post_wp_floats();
    float* dwp = wflip + post_wp_floats();
    float* wtc = dwp + post_wp_floats();
    float* dz = wtc + 2 * post_wp_floats();
    float* dbuf[2] = {dz + (size_t)n * POST_DIM, dz + (size_t)2 * n * POST_DIM};
    size_t off[POST_CONVS + 1];                                  // per-layer saved blocks
    off[0] = 0;
    for (int i = 0; i < POST_CONVS; ++i) off[i + 1] = off[i] + convbn_saved_floats(n, d.c_out[i]);
    const float* dy = d_dpost;                                   // d(residual) = d(post)
    for (int i = POST_CONVS - 1; i >= 0; --i) {
        const float* x = i == 0 ? d_coarse : d_saved + off[i - 1] + (size_t)n * d.c_out[i - 1];
        if (int e = conv_w_to_packed(wp, w->post_conv_w[i], d.c_out[i], d.c_in[i], POST_KERNEL, st)) return e;
        const int act = i == POST_CONVS - 1 ? ACT_NONE : ACT_TANH;
        float* dx = dbuf[i & 1];
        if (int e = convbn_train_bwd(dy, x, d.c_in[i], wp, w->post_bn_w[i], w->post_bn_b[i], d_masks ? d_masks[i] : nullptr,
                                     drop_scale, act, d_saved + off[i], dx, d.c_in[i], dwp,
                                     const_cast<float*>(grads->post_conv_b[i]), const_cast<float*>(grads->post_bn_w[i]),
                                     const_cast<float*>(grads->post_bn_b[i]), dz, wflip, n, S, d.c_in[i], d.c_out[i], POST_KERNEL,
                                     wtc, st))
            return e;
        if (int e = conv_w_from_packed(const_cast<float*>(grads->post_conv_w[i]), dwp, d.c_out[i], d.c_in[i], POST_KERNEL, st))
            return e;
        dy = dx;
    }
    return train_add(d_dcoarse, d_dpost, dy, (size_t)n * N_MELS, st);       // post = coarse + residual(coarse)
}

// ---------------------------------------------------------------------------------------------------
// Encoder in train mode: embedding -> 3 x relu(BN_batch(conv5)) -> BiLSTM, src/model.py:50-77.
// saved (floats): x0 [n][512] | 3 x convbn block (z, y, mean, var, rstd) | xproj [n][2048] | lstm [n][2][6*256]
namespace {
struct EncSaved { size_t x0, conv[ENC_CONVS], xproj, lstm, total; };
EncSaved enc_saved(int n) {
    EncSaved s{};
    size_t off = 0;
    s.x0 = off; off += (size_t)n * EMB;
    for (int i = 0; i < ENC_CONVS; ++i) { s.conv[i] = off; off += convbn_saved_floats(n, EMB); }
    s.xproj = off; off += (size_t)n * 8 * ENC_H;
    s.lstm = off; off += (size_t)n * 2 * LSTM_SAVE_F;
    s.total = off;
    return s;
}
struct EncScratch { size_t wp, wflip, dwp, dz, d0, d1, whhT, wihT, bias, dzl, hprev, dwl, total; };
EncScratch enc_scratch(int n) {
    EncScratch s{};
    size_t off = 0;
    const size_t W = (size_t)ENC_KERNEL * EMB * EMB;
    s.wp = off; off += W;
    s.wflip = off; off += W;
    s.dwp = off; off += W;
    s.dz = off; off += (size_t)n * EMB;
    s.d0 = off; off += (size_t)n * EMB;
    s.d1 = off; off += (size_t)n * EMB;
    s.whhT = off; off += (size_t)2 * ENC_H * 4 * ENC_H;
    s.wihT = off; off += (size_t)2 * EMB * 4 * ENC_H;
    s.bias = off; off += (size_t)8 * ENC_H;
    s.dzl = off; off += (size_t)n * 8 * ENC_H;
    s.hprev = off; off += (size_t)n * 2 * ENC_H;
    s.dwl = off; off += (size_t)EMB * 4 * ENC_H;
    s.total = off;
    return s;
}
}  // namespace

B200TACO_API size_t b200taco_encoder_train_saved_bytes(int32_t B, int32_t T) { return enc_saved(B * T).total * sizeof(float); }
B200TACO_API size_t b200taco_encoder_train_scratch_bytes(int32_t B, int32_t T) { return enc_scratch(B * T).total * sizeof(float); }

B200TACO_API int32_t b200taco_encoder_train_fwd(b200taco_handle* h, const b200taco_weights* w, const int64_t* d_ids, int32_t B,
                                                int32_t T, float* d_memory, float* d_saved, size_t saved_bytes, float* d_scratch,
                                                size_t scratch_bytes, void* stream) {
    if (int e = check_train_handle(h)) return e;
    TACO_REQUIRE(w && d_ids && d_memory && d_saved && d_scratch && B >= 1 && T >= 1, B200TACO_ERR_INVALID,
                 "encoder_train_fwd: bad argument");
    const int n = B * T;
    const EncSaved S = enc_saved(n);
    const EncScratch X = enc_scratch(n);
    TACO_REQUIRE(saved_bytes >= S.total * 4 && scratch_bytes >= X.total * 4, B200TACO_ERR_STATE, "encoder_train_fwd: buffers too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int e = embedding_gather(d_ids, w->enc_embedding, d_saved + S.x0, n, st)) return e;
    const float* x = d_saved + S.x0;
    // (no tensor-core weight planes: the encoder's ReLU + BatchNorm stack stays on the fp32 CUDA cores - with 3xTF32 at the wide
    // tile the forward error grows from 1.5e-7 to 3e-6 and flips enough ReLU units next to zero to move the encoder's gradients
    // outside the at-size test's bounds; the postnet (tanh, 5x the rows) is where the time was)
    for (int i = 0; i < ENC_CONVS; ++i) {
        if (int e = conv_w_to_packed(d_scratch + X.wp, w->enc_conv_w[i], EMB, EMB, ENC_KERNEL, st)) return e;
        if (int e = convbn_train_fwd(x, EMB, d_scratch + X.wp, w->enc_conv_b[i], w->enc_bn_w[i], w->enc_bn_b[i], nullptr, 1.f,
                                     ACT_RELU, d_saved + S.conv[i], n, T, EMB, EMB, ENC_KERNEL, nullptr, st))
            return e;
        x = d_saved + S.conv[i] + (size_t)n * EMB;
    }
    // input projections of both directions: xproj[:, dir*1024 ...] = x . W_ih[dir]^T + b_ih[dir] + b_hh[dir]
    for (int dir = 0; dir < 2; ++dir) {
        float* wihT = d_scratch + X.wihT + (size_t)dir * EMB * 4 * ENC_H;
        float* bias = d_scratch + X.bias + (size_t)dir * 4 * ENC_H;
        if (int e = transpose2d(wihT, w->enc_lstm_w_ih[dir], 4 * ENC_H, EMB, st)) return e;       // (1024,512) -> [512][1024]
        if (int e = train_add(bias, w->enc_lstm_b_ih[dir], w->enc_lstm_b_hh[dir], 4 * ENC_H, st)) return e;
        if (int e = conv1d_channels_last(x, EMB, wihT, nullptr, bias, nullptr, 0, d_saved + S.xproj + dir * 4 * ENC_H, 8 * ENC_H,
                                         n, n, EMB, 4 * ENC_H, 1, ACT_NONE, st))
            return e;
        if (int e = transpose2d(d_scratch + X.whhT + (size_t)dir * ENC_H * 4 * ENC_H, w->enc_lstm_w_hh[dir], 4 * ENC_H, ENC_H, st))
            return e;                                                                             // (1024,256) -> [256][1024]
    }
    return bilstm_train_fwd(d_scratch + X.whhT, d_scratch + X.whhT + (size_t)ENC_H * 4 * ENC_H, d_saved + S.xproj, d_memory,
                            d_saved + S.lstm, B, T, st);
}

B200TACO_API int32_t b200taco_encoder_train_bwd(b200taco_handle* h, const b200taco_weights* w, const int64_t* d_ids, int32_t B,
                                                int32_t T, const float* d_saved, const float* d_dmemory,
                                                const b200taco_weights* grads, float* d_scratch, size_t scratch_bytes,
                                                void* stream) {
    if (int e = check_train_handle(h)) return e;
    TACO_REQUIRE(w && d_ids && d_saved && d_dmemory && grads && d_scratch, B200TACO_ERR_INVALID, "encoder_train_bwd: bad argument");
    const int n = B * T;
    const EncSaved S = enc_saved(n);
    const EncScratch X = enc_scratch(n);
    TACO_REQUIRE(scratch_bytes >= X.total * 4, B200TACO_ERR_STATE, "encoder_train_bwd: scratch too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* dzl = d_scratch + X.dzl;
    float* hprev = d_scratch + X.hprev;
    float* dwl = d_scratch + X.dwl;
    // ---- BiLSTM backward through time
    if (int e = bilstm_train_bwd(w->enc_lstm_w_hh[0], w->enc_lstm_w_hh[1], d_dmemory, d_saved + S.lstm, dzl, hprev, B, T, st)) return e;
    const float* y3 = d_saved + S.conv[ENC_CONVS - 1] + (size_t)n * EMB;
    float* dx = d_scratch + X.d0;
    for (int dir = 0; dir < 2; ++dir) {
        const float* dz = dzl + dir * 4 * ENC_H;
        // dW_hh = dz^T . h_prev ; dW_ih = dz^T . x  (computed as [in][out], then transposed to the parameter layout)
        if (int e = dense_wgrad(hprev + dir * ENC_H, 2 * ENC_H, dz, 8 * ENC_H, dwl, n, n, ENC_H, 4 * ENC_H, 1, st)) return e;
        if (int e = transpose2d(const_cast<float*>(grads->enc_lstm_w_hh[dir]), dwl, ENC_H, 4 * ENC_H, st)) return e;
        if (int e = dense_wgrad(y3, EMB, dz, 8 * ENC_H, dwl, n, n, EMB, 4 * ENC_H, 1, st)) return e;
        if (int e = transpose2d(const_cast<float*>(grads->enc_lstm_w_ih[dir]), dwl, EMB, 4 * ENC_H, st)) return e;
        // dx += dz . W_ih  (raw parameter (1024, 512) is already [in = gate row][out = input channel])
        if (int e = conv1d_channels_last(dz, 8 * ENC_H, w->enc_lstm_w_ih[dir], nullptr, nullptr, dir ? dx : nullptr, EMB, dx, EMB,
                                         n, n, 4 * ENC_H, EMB, 1, ACT_NONE, st))
            return e;
    }
    // bias gradients: b_ih and b_hh enter as a sum
    if (int e = train_colsum(dwl, dzl, n, 8 * ENC_H, st)) return e;
    for (int dir = 0; dir < 2; ++dir) {
        TACO_CUDA(cudaMemcpyAsync(const_cast<float*>(grads->enc_lstm_b_ih[dir]), dwl + dir * 4 * ENC_H, 4 * ENC_H * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
        TACO_CUDA(cudaMemcpyAsync(const_cast<float*>(grads->enc_lstm_b_hh[dir]), dwl + dir * 4 * ENC_H, 4 * ENC_H * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
    }
    // ---- convolution stack
    const float* dy = dx;
    for (int i = ENC_CONVS - 1; i >= 0; --i) {
        const float* xin = i == 0 ? d_saved + S.x0 : d_saved + S.conv[i - 1] + (size_t)n * EMB;
        if (int e = conv_w_to_packed(d_scratch + X.wp, w->enc_conv_w[i], EMB, EMB, ENC_KERNEL, st)) return e;
        float* dxi = d_scratch + (((ENC_CONVS - 1 - i) & 1) ? X.d0 : X.d1);
        if (int e = convbn_train_bwd(dy, xin, EMB, d_scratch + X.wp, w->enc_bn_w[i], w->enc_bn_b[i], nullptr, 1.f, ACT_RELU,
                                     d_saved + S.conv[i], dxi, EMB, d_scratch + X.dwp, const_cast<float*>(grads->enc_conv_b[i]),
                                     const_cast<float*>(grads->enc_bn_w[i]), const_cast<float*>(grads->enc_bn_b[i]),
                                     d_scratch + X.dz, d_scratch + X.wflip, n, T, EMB, EMB, ENC_KERNEL, nullptr, st))
            return e;
        if (int e = conv_w_from_packed(const_cast<float*>(grads->enc_conv_w[i]), d_scratch + X.dwp, EMB, EMB, ENC_KERNEL, st)) return e;
        dy = dxi;
    }
    return embedding_bwd(d_ids, dy, const_cast<float*>(grads->enc_embedding), n, st);
}

// ---------------------------------------------------------------------------------------------------
// Decoder backward (teacher-forced, train mode).  The host walks the steps in reverse; per step: projection^T,
// decoder LSTM cell backward + input/recurrent data gradients, attention backward, query^T, attention LSTM cell
// backward + data gradients.  Weight gradients are deferred to a few large reductions over all steps.
namespace {
struct DecScratch {
    size_t dout, wpg, wqT, dzd, dza, dpq, dp, d_ah, d_dh, d_ctx, dca, dcd, tmp, tmp2, pq, P, gsum, cum, dpm, attn, dw, dvs, wcat_d, wcat_a, dctx_all,
        wpk_d, wpk_a, dproj, dzt, total;
};
DecScratch dec_scratch(int B, int T, int S) {
    DecScratch d{};
    const size_t n = (size_t)S * B;
    size_t off = 0;
    auto take = [&](size_t f) { size_t o = off; off += (f + 63) / 64 * 64; return o; };
    d.dout = take(n * PROJ_PAD);
    d.wpg = take((size_t)PROJ_PAD * (RNN + EMB));
    d.wqT = take((size_t)RNN * ATT_DIM);
    d.dzd = take(n * 4 * RNN);
    d.dza = take(n * 4 * RNN);
    d.dpq = take(n * ATT_DIM);
    d.dp = take(n * PRENET * 2);                  // dp_all, then reused for the prenet backward (two buffers)
    d.d_ah = take((size_t)B * RNN);
    d.d_dh = take((size_t)B * RNN);
    d.d_ctx = take((size_t)B * EMB);
    d.dca = take((size_t)B * RNN);
    d.dcd = take((size_t)B * RNN);
    d.tmp = take((size_t)B * (RNN + EMB));
    d.tmp2 = take((size_t)B * RNN);
    d.pq = take(n * ATT_DIM);                     // processed queries of all steps (recomputed once: ah . Wq^T)
    d.P = take((size_t)2 * B * T);
    d.gsum = take((size_t)B * T);
    d.cum = take(n * T);
    d.dpm = take((size_t)B * T * ATT_DIM);
    d.attn = take(attention_bwd_scratch_floats(B, T));
    d.dw = take((size_t)4 * RNN * (RNN + EMB));   // largest [in][out] weight gradient (1536 x 4096)
    d.dvs = take((size_t)2 * B * ATTN_PACC_F + ATTN_PACC_F);   // attention parameter-gradient accumulators (two per utterance: one per CTA of the pair) + their sum
    d.wcat_d = take((size_t)4 * RNN * (2 * RNN + EMB));        // [W_ih | W_hh] of the decoder LSTM, [4096][2560]
    d.wcat_a = take((size_t)4 * RNN * (PRENET + EMB + RNN));   // [W_ih | W_hh] of the attention LSTM, [4096][1792]
    d.dctx_all = take(n * EMB);                   // d ctx_t of every step: the memory gradient is one contraction after the walk
    d.wpk_d = take((size_t)4 * RNN * (2 * RNN + EMB) / 2);       // the same two matrices as bf16 UMMA tiles (train_skinny_tc.cu)
    d.wpk_a = take((size_t)4 * RNN * (PRENET + EMB + RNN) / 2);
    d.dproj = take(n * (RNN + EMB));              // gradient of the projection inputs [dh_t | ctx_t] of every step (one dense call)
    d.dzt = take(2 * skinny_tc_act_bytes(4 * RNN) / sizeof(float));   // this step's gate gradients of both cells as hi/lo bf16 UMMA tiles
    d.total = off;
    return d;
}
}  // namespace

B200TACO_API size_t b200taco_decoder_train_scratch_bytes(int32_t B, int32_t T, int32_t S) {
    return dec_scratch(B, T, S).total * sizeof(float);
}

B200TACO_API int32_t b200taco_decoder_train_bwd(b200taco_handle* h, const b200taco_weights* w, const float* d_memory,
                                                const int64_t* d_text_lengths, int32_t B, int32_t T, int32_t S,
                                                const float* d_mask_p1, const float* d_mask_p2, float prenet_scale,
                                                const float* d_mask_a, const float* d_mask_d, float lstm_scale,
                                                const float* d_saved, const float* d_align, const float* d_dcoarse,
                                                const float* d_dgate, const float* d_dalign, float* d_dmemory,
                                                const b200taco_weights* grads, float* d_scratch, size_t scratch_bytes,
                                                void* stream) {
    if (int e = check_train_handle(h)) return e;
    TACO_REQUIRE(w && d_memory && d_saved && d_align && d_dcoarse && d_dgate && d_dmemory && grads && d_scratch,
                 B200TACO_ERR_INVALID, "decoder_train_bwd: bad argument");
    const DecScratch X = dec_scratch(B, T, S);
    TACO_REQUIRE(scratch_bytes >= X.total * sizeof(float), B200TACO_ERR_STATE, "decoder_train_bwd: scratch too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const DecSaved sv = dec_saved(B, T, S);
    const int n = S * B, BT = B * T;
    float* sc = d_scratch;
    float *dout = sc + X.dout, *wpg = sc + X.wpg, *wqT = sc + X.wqT, *dzd = sc + X.dzd, *dza = sc + X.dza, *dpq = sc + X.dpq,
          *dp = sc + X.dp, *d_ah = sc + X.d_ah, *d_dh = sc + X.d_dh, *d_ctx = sc + X.d_ctx, *dca = sc + X.dca, *dcd = sc + X.dcd,
          *tmp = sc + X.tmp, *tmp2 = sc + X.tmp2, *pq = sc + X.pq, *P = sc + X.P, *gsum = sc + X.gsum, *cum = sc + X.cum,
          *dpm = sc + X.dpm, *attn = sc + X.attn, *dw = sc + X.dw, *dvs = sc + X.dvs, *dctx_all = sc + X.dctx_all, *wcat_d = sc + X.wcat_d,
          *wcat_a = sc + X.wcat_a, *dproj = sc + X.dproj;
    const float *frames = d_saved + sv.frames, *y1 = d_saved + sv.y1, *y2 = d_saved + sv.y2, *pm = d_saved + sv.pm,
                *ga = d_saved + sv.ga, *ca = d_saved + sv.ca, *ah = d_saved + sv.ah, *gd = d_saved + sv.gd, *cd = d_saved + sv.cd,
                *dh = d_saved + sv.dh, *ctx = d_saved + sv.ctx;
    auto G = [](const float* p) { return const_cast<float*>(p); };
    auto dense = [&](const float* Xp, int ldx, const float* W, float* Y, int ldy, int rows, int c_in, int c_out, const float* R) {
        return conv1d_channels_last(Xp, ldx, W, nullptr, nullptr, R, ldy, Y, ldy, rows, rows, c_in, c_out, 1, ACT_NONE, st);
    };
#define TT(expr) do { int _e = (expr); if (_e != B200TACO_OK) return _e; } while (0)

    // ---- set-up
    TT(assemble_dout(d_dcoarse, d_dgate, dout, B, S, st));
    TT(assemble_wpg(w->proj_w, w->gate_w, wpg, st));
    TT(transpose2d(wqT, w->attn_query_w, ATT_DIM, RNN, st));                    // (128,1024) -> [1024][128]
    TT(attn_cumsum(d_align, cum, S, BT, st));
    // [W_ih | W_hh] side by side: the three data gradients of a cell are one contraction over its gate gradients
    constexpr int ND = 2 * RNN + EMB, NA = PRENET + EMB + RNN;
    const size_t f4 = sizeof(float);
    TACO_CUDA(cudaMemcpy2DAsync(wcat_d, ND * f4, w->dec_lstm_w_ih, (RNN + EMB) * f4, (RNN + EMB) * f4, 4 * RNN, cudaMemcpyDeviceToDevice, st));
    TACO_CUDA(cudaMemcpy2DAsync(wcat_d + RNN + EMB, ND * f4, w->dec_lstm_w_hh, RNN * f4, RNN * f4, 4 * RNN, cudaMemcpyDeviceToDevice, st));
    TACO_CUDA(cudaMemcpy2DAsync(wcat_a, NA * f4, w->attn_lstm_w_ih, (PRENET + EMB) * f4, (PRENET + EMB) * f4, 4 * RNN, cudaMemcpyDeviceToDevice, st));
    TACO_CUDA(cudaMemcpy2DAsync(wcat_a + PRENET + EMB, NA * f4, w->attn_lstm_w_hh, RNN * f4, RNN * f4, 4 * RNN, cudaMemcpyDeviceToDevice, st));
    // mixed precision (the forward multiplied bf16-rounded LSTM matrices): the per-step data gradients run on the tensor cores with
    // the same bf16 values; an fp32-weight handle (train_lstm_weight_dtype = "fp32") keeps the fp32 CUDA-core contraction
    const int ends_d[3] = {RNN, RNN + EMB, ND}, ends_a[3] = {PRENET, PRENET + EMB, NA};
    const bool tc_skinny = h->weight_dtype == 0 && skinny_tc_supported(B, 4 * RNN, ND, 3, ends_d) && skinny_tc_supported(B, 4 * RNN, NA, 3, ends_a);
    void *wpk_d = sc + X.wpk_d, *wpk_a = sc + X.wpk_a;
    void* dzt_d = tc_skinny ? static_cast<void*>(sc + X.dzt) : nullptr;
    void* dzt_a = tc_skinny ? static_cast<void*>(reinterpret_cast<uint8_t*>(sc + X.dzt) + skinny_tc_act_bytes(4 * RNN)) : nullptr;
    if (tc_skinny) {
        TACO_CUDA(cudaMemsetAsync(sc + X.dzt, 0, 2 * skinny_tc_act_bytes(4 * RNN), st));     // rows >= B of the tiles stay zero
        TT(pack_skinny_tc(wpk_d, wcat_d, ND, 4 * RNN, ND, st));
        TT(pack_skinny_tc(wpk_a, wcat_a, NA, 4 * RNN, NA, st));
    }
    TACO_CUDA(cudaMemsetAsync(d_ah, 0, (X.tmp - X.d_ah) * sizeof(float), st));  // d_ah, d_dh, d_ctx, dca, dcd
    TACO_CUDA(cudaMemsetAsync(dp, 0, (size_t)n * PRENET * sizeof(float), st));   // the per-step GEMMs accumulate onto zeros
    // processed queries (src/model.py:123) of all steps in one dense call instead of one skinny GEMM per step
    TT(dense(ah, RNN, wqT, pq, ATT_DIM, n, RNN, ATT_DIM, nullptr));
    // projections (src/model.py:272-274): proj_in = [dh_t | ctx_t].  Their gradient does not depend on the walk (dout is the loss
    // gradient of every step), so it is ONE dense call here instead of a skinny GEMM per step; the consumers of the carried
    // gradients (decoder-LSTM cell backward: d dh_t, attention backward: d ctx_t) add row block t of it as they read them.
    TT(dense(dout, PROJ_PAD, wpg, dproj, RNN + EMB, n, PROJ_PAD, RNN + EMB, nullptr));
    // The consumers of the carried gradients (LSTM-cell backward: d ah / d dh; second attention backward: d ctx) zero them
    // after reading, so the data-gradient GEMM that follows accumulates the previous step's gradient without a memset.
    const bool v2 = attention_bwd_defers_dmemory(T);
    const bool fuse_cell = attention_bwd_fuses_cell(T);
    static const bool pair_gemm = [] { const char* e = getenv("TACO_SKINNY_PAIR"); return !(e && e[0] == '0'); }();
    TACO_CUDA(cudaMemsetAsync(gsum, 0, (size_t)BT * sizeof(float), st));
    TACO_CUDA(cudaMemsetAsync(dpm, 0, (size_t)BT * ATT_DIM * sizeof(float), st));
    TACO_CUDA(cudaMemsetAsync(d_dmemory, 0, (size_t)BT * EMB * sizeof(float), st));
    TACO_CUDA(cudaMemsetAsync(dvs, 0, ((size_t)2 * B * ATTN_PACC_F + ATTN_PACC_F) * sizeof(float), st));

    // ---- reverse walk over the decode steps
    // decoder LSTM of step t (input [ah_t | ctx_t], src/model.py:265-270): cell backward (the projection's share of d dh_t comes from
    // dproj) and the data gradients d ah_t +=, d ctx_t +=, d dh_{t-1} =.  With the fused tail the cell backward of step t - 1 rides
    // in the attention backward of step t (its input is complete once the data-gradient GEMM of step t is), and the GEMM of step
    // t - 1 follows the attention-LSTM GEMM of step t: three kernels per decode step.
    auto dec_cell = [&](int t) -> int {
        const size_t r = (size_t)t * B;
        return lstm_cell_bwd(gd + r * 4 * RNN, cd + r * RNN, t > 0 ? cd + (r - B) * RNN : nullptr, d_dh, dproj + r * (RNN + EMB), RNN + EMB,
                             d_mask_d ? d_mask_d + r * RNN : nullptr, lstm_scale, dcd, dzd + r * 4 * RNN, dzt_d, B, st);
    };
    auto dec_gemm = [&](int t) -> int {
        const size_t r = (size_t)t * B;
        float* outs[3] = {d_ah, d_ctx, d_dh};
        const int ldos[3] = {RNN, EMB, RNN}, ends[3] = {RNN, RNN + EMB, ND};
        const bool acc[3] = {true, true, true};
        if (tc_skinny) return skinny_gemm_multi_tc(dzt_d, wpk_d, B, 4 * RNN, 3, outs, ldos, ends, acc, st);
        return skinny_gemm_multi(dzd + r * 4 * RNN, 4 * RNN, wcat_d, ND, B, 4 * RNN, 3, outs, ldos, ends, acc, st);
    };
    if (fuse_cell) {
        TT(dec_cell(S - 1));
        TT(dec_gemm(S - 1));
    }
    for (int t = S - 1; t >= 0; --t) {
        const size_t r = (size_t)t * B;                       // first row of step t in the [S*B][...] buffers
        const float* dproj_t = dproj + r * (RNN + EMB);
        if (!fuse_cell) {
            TT(dec_cell(t));
            TT(dec_gemm(t));
        }
        // attention (src/model.py:117-148); with the cluster kernel its tail also does the query layer's data gradient and the
        // pointwise backward of the attention-LSTM cell (input [p_t | ctx_{t-1}], src/model.py:256-262)
        float* dza_t = dza + r * 4 * RNN;
        AttnCellTail tail{};
        tail.Wq = w->attn_query_w; tail.d_ah = d_ah; tail.gates = ga + r * 4 * RNN; tail.c = ca + r * RNN;
        tail.c_prev = t > 0 ? ca + (r - B) * RNN : nullptr; tail.mask = d_mask_a ? d_mask_a + r * RNN : nullptr; tail.scale = lstm_scale;
        tail.dc = dca; tail.dz = dza_t; tail.dzp = static_cast<__nv_bfloat16*>(dzt_a);
        if (t > 0) {                                          // decoder-LSTM cell backward of step t - 1
            const size_t r1 = r - B;
            tail.dh2 = d_dh; tail.dh2_extra = dproj + r1 * (RNN + EMB); tail.ld2 = RNN + EMB; tail.gates2 = gd + r1 * 4 * RNN;
            tail.c2 = cd + r1 * RNN; tail.c_prev2 = t > 1 ? cd + (r1 - B) * RNN : nullptr; tail.mask2 = d_mask_d ? d_mask_d + r1 * RNN : nullptr;
            tail.dc2 = dcd; tail.dz2 = dzd + r1 * 4 * RNN; tail.dzp2 = static_cast<__nv_bfloat16*>(dzt_d);
        }
        TT(attention_bwd(d_align + r * T, t > 0 ? d_align + (r - B) * T : nullptr, cum + r * T, pq + r * ATT_DIM, pm, d_memory, d_ctx,
                         d_dalign ? d_dalign + r * T : nullptr, t < S - 1 ? P + (size_t)((t + 1) & 1) * BT : nullptr, gsum,
                         d_text_lengths, w->attn_loc_conv_w, w->attn_loc_dense_w, w->attn_v_w, w->attn_v_b, w->attn_energy_scale,
                         d_dmemory, dpm, dpq + r * ATT_DIM, P + (size_t)(t & 1) * BT, dvs, attn, dctx_all + r * EMB, dproj_t + RNN, RNN + EMB,
                         fuse_cell ? &tail : nullptr, B, T, st));
        if (!fuse_cell) {
            TT(skinny_gemm(dpq + r * ATT_DIM, ATT_DIM, w->attn_query_w, RNN, d_ah, RNN, B, ATT_DIM, RNN, true, st));
            TT(lstm_cell_bwd(ga + r * 4 * RNN, ca + r * RNN, t > 0 ? ca + (r - B) * RNN : nullptr, d_ah, nullptr, 0, d_mask_a ? d_mask_a + r * RNN : nullptr,
                             lstm_scale, dca, dza_t, dzt_a, B, st));
        }
        {   // d p_t =, d ctx_{t-1} =, d ah_{t-1} =   (and, fused walk: the decoder LSTM's d ah_{t-1} +=, d ctx_{t-1} +=, d dh_{t-2} = of step
            // t - 1 in the same launch: the two GEMMs are independent once both cell backwards ran in the attention kernel)
            float* outs[3] = {dp + r * PRENET, d_ctx, d_ah};
            const int ldos[3] = {PRENET, EMB, RNN}, ends[3] = {PRENET, PRENET + EMB, NA};
            const bool acc[3] = {true, v2, true};
            float* outs_d[3] = {d_ah, d_ctx, d_dh};
            const int ldos_d[3] = {RNN, EMB, RNN}, ends_dd[3] = {RNN, RNN + EMB, ND};
            const bool acc_d[3] = {true, true, true};
            if (fuse_cell && t > 0 && tc_skinny && pair_gemm) {
                TT(skinny_gemm_pair_tc(dzt_a, wpk_a, 3, outs, ldos, ends, acc, dzt_d, wpk_d, 3, outs_d, ldos_d, ends_dd, acc_d, B, 4 * RNN, st));
            } else {
                if (tc_skinny) TT(skinny_gemm_multi_tc(dzt_a, wpk_a, B, 4 * RNN, 3, outs, ldos, ends, acc, st));
                else TT(skinny_gemm_multi(dza_t, 4 * RNN, wcat_a, NA, B, 4 * RNN, 3, outs, ldos, ends, acc, st));
                if (fuse_cell && t > 0) TT(dec_gemm(t - 1));
            }
        }
    }

    // ---- weight gradients: reductions over all steps, computed as [in][out] and transposed into the parameter layout
    auto wgrad_into = [&](const float* Xp, int ldx, int c_in, const float* dY, int ldy, int c_out, int rows, float* dst, int ldd) {
        if (int e = dense_wgrad(Xp, ldx, dY, ldy, dw, rows, rows, c_in, c_out, 1, st)) return e;
        return transpose_block(dst, ldd, dw, c_out, c_in, c_out, st);         // dst[out][in]
    };
    // decoder LSTM: W_ih (4096 x [ah 1024 | ctx 512]), W_hh (4096 x 1024, input dh_{t-1})
    TT(wgrad_into(ah, RNN, RNN, dzd, 4 * RNN, 4 * RNN, n, G(grads->dec_lstm_w_ih), RNN + EMB));
    TT(wgrad_into(ctx, EMB, EMB, dzd, 4 * RNN, 4 * RNN, n, G(grads->dec_lstm_w_ih) + RNN, RNN + EMB));
    TACO_CUDA(cudaMemsetAsync(G(grads->dec_lstm_w_hh), 0, (size_t)4 * RNN * RNN * sizeof(float), st));
    if (S > 1) TT(wgrad_into(dh, RNN, RNN, dzd + (size_t)B * 4 * RNN, 4 * RNN, 4 * RNN, n - B, G(grads->dec_lstm_w_hh), RNN));
    TT(train_colsum(G(grads->dec_lstm_b_ih), dzd, n, 4 * RNN, st));
    TACO_CUDA(cudaMemcpyAsync(G(grads->dec_lstm_b_hh), grads->dec_lstm_b_ih, 4 * RNN * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // attention LSTM: W_ih (4096 x [p 256 | ctx_{t-1} 512]), W_hh (input ah_{t-1})
    TT(wgrad_into(y2, PRENET, PRENET, dza, 4 * RNN, 4 * RNN, n, G(grads->attn_lstm_w_ih), PRENET + EMB));
    TACO_CUDA(cudaMemsetAsync(dw, 0, (size_t)EMB * 4 * RNN * sizeof(float), st));
    if (S > 1) TT(dense_wgrad(ctx, EMB, dza + (size_t)B * 4 * RNN, 4 * RNN, dw, n - B, n - B, EMB, 4 * RNN, 1, st));
    TT(transpose_block(G(grads->attn_lstm_w_ih) + PRENET, PRENET + EMB, dw, 4 * RNN, EMB, 4 * RNN, st));
    TACO_CUDA(cudaMemsetAsync(G(grads->attn_lstm_w_hh), 0, (size_t)4 * RNN * RNN * sizeof(float), st));
    if (S > 1) TT(wgrad_into(ah, RNN, RNN, dza + (size_t)B * 4 * RNN, 4 * RNN, 4 * RNN, n - B, G(grads->attn_lstm_w_hh), RNN));
    TT(train_colsum(G(grads->attn_lstm_b_ih), dza, n, 4 * RNN, st));
    TACO_CUDA(cudaMemcpyAsync(G(grads->attn_lstm_b_hh), grads->attn_lstm_b_ih, 4 * RNN * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // query layer (128 x 1024)
    TT(wgrad_into(ah, RNN, RNN, dpq, ATT_DIM, ATT_DIM, n, G(grads->attn_query_w), RNN));
    // projections: [dh_t | ctx_t] -> 80 mel rows + 1 gate row (dout has 96 padded columns)
    TT(dense_wgrad(dh, RNN, dout, PROJ_PAD, dw, n, n, RNN, PROJ_PAD, 1, st));
    TT(transpose_block(G(grads->proj_w), RNN + EMB, dw, PROJ_PAD, RNN, N_MELS, st));
    TT(transpose_block(G(grads->gate_w), RNN + EMB, dw + N_MELS, PROJ_PAD, RNN, 1, st));
    TT(dense_wgrad(ctx, EMB, dout, PROJ_PAD, dw, n, n, EMB, PROJ_PAD, 1, st));
    TT(transpose_block(G(grads->proj_w) + RNN, RNN + EMB, dw, PROJ_PAD, EMB, N_MELS, st));
    TT(transpose_block(G(grads->gate_w) + RNN, RNN + EMB, dw + N_MELS, PROJ_PAD, EMB, 1, st));
    TT(train_colsum(dw, dout, n, PROJ_PAD, st));
    TACO_CUDA(cudaMemcpyAsync(G(grads->proj_b), dw, N_MELS * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TACO_CUDA(cudaMemcpyAsync(G(grads->gate_b), dw + N_MELS, sizeof(float), cudaMemcpyDeviceToDevice, st));
    // attention parameters: sum of the per-utterance accumulators
    {
        float* tot = dvs + (size_t)2 * B * ATTN_PACC_F;
        TT(train_colsum(tot, dvs, 2 * B, ATTN_PACC_F, st));
        const size_t f = sizeof(float);
        TACO_CUDA(cudaMemcpyAsync(G(grads->attn_loc_dense_w), tot, ATT_DIM * LOC_F * f, cudaMemcpyDeviceToDevice, st));
        TACO_CUDA(cudaMemcpyAsync(G(grads->attn_loc_conv_w), tot + ATT_DIM * LOC_F, LOC_F * 2 * LOC_K * f, cudaMemcpyDeviceToDevice, st));
        TACO_CUDA(cudaMemcpyAsync(G(grads->attn_v_w), tot + ATT_DIM * LOC_F + LOC_F * 2 * LOC_K, ATT_DIM * f, cudaMemcpyDeviceToDevice, st));
        TACO_CUDA(cudaMemcpyAsync(G(grads->attn_v_b), tot + ATT_DIM * LOC_F + LOC_F * 2 * LOC_K + ATT_DIM, f, cudaMemcpyDeviceToDevice, st));
        TACO_CUDA(cudaMemcpyAsync(G(grads->attn_energy_scale), tot + ATT_DIM * LOC_F + LOC_F * 2 * LOC_K + ATT_DIM + 1, f,
                                  cudaMemcpyDeviceToDevice, st));
    }
    // context (src/model.py:147): dmemory[b][tau][:] = sum_t w_t[b][tau] * dctx_t[b][:], one [T x S] . [S x 512] contraction per
    // utterance when the attention backward deferred it
    if (attention_bwd_defers_dmemory(T)) TT(ctx_dmemory(d_align, dctx_all, d_dmemory, S, B, T, st));
    // memory layer (128 x 512): pm = memory . W^T
    TT(wgrad_into(d_memory, EMB, EMB, dpm, ATT_DIM, ATT_DIM, BT, G(grads->attn_memory_w), EMB));
    TT(dense(dpm, ATT_DIM, w->attn_memory_w, d_dmemory, EMB, BT, ATT_DIM, EMB, d_dmemory));               // dmemory += dpm . W
    // prenet (src/model.py:166-171): y2 = drop(relu(y1 W2^T)), y1 = drop(relu(x W1^T))
    float* dy2 = dp + (size_t)n * PRENET;
    TT(relu_drop_bwd(dp, y2, d_mask_p2, prenet_scale, dy2, (size_t)n * PRENET, st));
    TT(wgrad_into(y1, PRENET, PRENET, dy2, PRENET, PRENET, n, G(grads->prenet_w[1]), PRENET));
    TT(dense(dy2, PRENET, w->prenet_w[1], dp, PRENET, n, PRENET, PRENET, nullptr));                       // d y1 (post-dropout)
    TT(relu_drop_bwd(dp, y1, d_mask_p1, prenet_scale, dy2, (size_t)n * PRENET, st));
    TT(wgrad_into(frames, N_MELS, N_MELS, dy2, PRENET, PRENET, n, G(grads->prenet_w[0]), N_MELS));
#undef TT
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_attn_bwd_phase_ns(uint64_t* out8, int32_t reset) {
    TACO_CUDA(cudaDeviceSynchronize());
    return attn_bwd_phase_ns(reinterpret_cast<unsigned long long*>(out8), reset);
}

}  // extern "C"
