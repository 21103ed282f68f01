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
                     int taps, cudaStream_t st);
int convbn_train_bwd(const float* dY, const float* X, int ldx, const float* Wp, const float* gamma, const float* beta,
                     const float* mask, float drop_scale, int act, const float* saved, float* dX, int lddx, float* dWp,
                     float* db, float* dgamma, float* dbeta, float* dz, float* wflip, int n_rows, int seg_len, int c_in,
                     int c_out, int taps, cudaStream_t st);

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
    // packed weights, flipped weights, packed dW; dz and two d-activation buffers
    return (3 * post_wp_floats() + (size_t)3 * B * S * POST_DIM) * sizeof(float);
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
    for (int i = 0; i < POST_CONVS; ++i) {
        if (int e = conv_w_to_packed(wp, w->post_conv_w[i], d.c_out[i], d.c_in[i], POST_KERNEL, st)) return e;
        const int act = i == POST_CONVS - 1 ? ACT_NONE : ACT_TANH;
        if (int e = convbn_train_fwd(x, d.c_in[i], wp, w->post_conv_b[i], w->post_bn_w[i], w->post_bn_b[i],
                                     d_masks ? d_masks[i] : nullptr, drop_scale, act, saved, n, S, d.c_in[i], d.c_out[i],
                                     POST_KERNEL, st))
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
    float* dz = dwp + post_wp_floats();
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
                                     st))
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
    for (int i = 0; i < ENC_CONVS; ++i) {
        if (int e = conv_w_to_packed(d_scratch + X.wp, w->enc_conv_w[i], EMB, EMB, ENC_KERNEL, st)) return e;
        if (int e = convbn_train_fwd(x, EMB, d_scratch + X.wp, w->enc_conv_b[i], w->enc_bn_w[i], w->enc_bn_b[i], nullptr, 1.f,
                                     ACT_RELU, d_saved + S.conv[i], n, T, EMB, EMB, ENC_KERNEL, st))
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
                                     d_scratch + X.dz, d_scratch + X.wflip, n, T, EMB, EMB, ENC_KERNEL, st))
            return e;
        if (int e = conv_w_from_packed(const_cast<float*>(grads->enc_conv_w[i]), d_scratch + X.dwp, EMB, EMB, ENC_KERNEL, st)) return e;
        dy = dxi;
    }
    return embedding_bwd(d_ids, dy, const_cast<float*>(grads->enc_embedding), n, st);
}

}  // extern "C"
