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

}  // extern "C"
