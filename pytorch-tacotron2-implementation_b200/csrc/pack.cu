// Weight repacking: fp32 parameters in the reference's state_dict layout (include/b200taco.h,
// b200taco_weights) -> the layouts the kernels stream.  Runs once per load_state_dict /
// optimizer step, so these kernels are written for clarity, not speed.
#include "handle.h"

namespace b200taco {

namespace {

constexpr float BN_EPS = 1e-5f;   // nn.BatchNorm1d default (src/model.py:36, :357)

template <typename WT>
__device__ __forceinline__ WT to_wt(float v);
template <>
__device__ __forceinline__ float to_wt<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_wt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// dst[cta][k][r], r = unit*4 + gate, source row = gate*1024 + 8*cta + unit.
// k runs over three column segments: (src0, ld0, off0, n0), (src1, ...), (src2, ...).
template <typename WT>
__global__ void pack_lstm_kernel(WT* dst, int K, const float* s0, int ld0, int off0, int n0, const float* s1, int ld1,
                                 int off1, int n1, const float* s2, int ld2, int off2) {
    const size_t total = (size_t)GEMM_CTAS * K * ROWS_PER_CTA;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = i % ROWS_PER_CTA;
        const int k = (i / ROWS_PER_CTA) % K;
        const int cta = i / ((size_t)ROWS_PER_CTA * K);
        const int row = (r & 3) * RNN + cta * UNITS_PER_CTA + (r >> 2);
        float v;
        if (k < n0) v = s0[(size_t)row * ld0 + off0 + k];
        else if (k < n0 + n1) v = s1[(size_t)row * ld1 + off1 + (k - n0)];
        else v = s2[(size_t)row * ld2 + off2 + (k - n0 - n1)];
        dst[i] = to_wt<WT>(v);
    }
}

// tensor-core layout (split-K groups of 4 CTAs, decoder.cu gemm_tc): cta = 4*group + member owns, for the
// group's 128 gate rows (n = unit_local*4 + gate, unit = 32*group + unit_local), the K quarter of `member`:
// dst[cta][j = local chunk][UMMA K-major tile of 128 rows x 64 k], physical k = (member*K/4) + 64*j + kk
__global__ void pack_lstm_tc_kernel(__nv_bfloat16* dst, int K, int which /* 0 attention, 1 decoder LSTM */, const float* s0, int ld0, int off0, int n0, const float* s1,
                                    int ld1, int off1, int n1, const float* s2, int ld2, int off2) {
    const int KQ = K / 4;                                      // k per member
    const size_t per_cta = (size_t)KQ * 128;
    const size_t total = (size_t)GEMM_CTAS * per_cta;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kl = i % KQ;                                 // k inside the member's quarter
        const int n = (i / KQ) % 128;
        const int cta = i / per_cta;
        const int group = cta / 4, member = cta % 4;
        // local chunk kl/64 of this member -> chunk of the act vector -> position in this matrix' own K order
        const int ach = which ? tc_dec_act_chunk(member, kl / 64) - ACT_CTX / 64 : tc_attn_act_chunk(member, kl / 64);
        const int k = ach * 64 + kl % 64;
        const int row = (n & 3) * RNN + group * 32 + (n >> 2);
        float v;
        if (k < n0) v = s0[(size_t)row * ld0 + off0 + k];
        else if (k < n0 + n1) v = s1[(size_t)row * ld1 + off1 + (k - n0)];
        else v = s2[(size_t)row * ld2 + off2 + (k - n0 - n1)];
        const size_t o = (size_t)cta * per_cta + (size_t)(kl / 64) * (64 * 128) +
                         ((((n >> 3) * 8 + ((kl % 64) >> 3)) << 6) + ((n & 7) << 3) + (kl & 7));
        dst[o] = __float2bfloat16_rn(v);
    }
}

__global__ void pack_lstm_bias_kernel(float* dst, const float* b_ih, const float* b_hh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < GEMM_CTAS * ROWS_PER_CTA) {
        const int r = i % ROWS_PER_CTA, cta = i / ROWS_PER_CTA;
        const int row = (r & 3) * RNN + cta * UNITS_PER_CTA + (r >> 2);
        dst[i] = b_ih[row] + b_hh[row];
    }
}

// dst[cta][o][u] = src_o[o][col0 + 8*cta + u]; rows o < n_a come from a (ld), row n_a.. from b.
__global__ void pack_slices_kernel(float* dst, int n_out, const float* a, int n_a, const float* b, int ld) {
    const int total = GEMM_CTAS * n_out * UNITS_PER_CTA;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int u = i % UNITS_PER_CTA, o = (i / UNITS_PER_CTA) % n_out, cta = i / (UNITS_PER_CTA * n_out);
        const int col = cta * UNITS_PER_CTA + u;
        dst[i] = o < n_a ? a[(size_t)o * ld + col] : b[(size_t)(o - n_a) * ld + col];
    }
}

__global__ void pack_proj_ctx_kernel(float* wT, float* bias, const float* wp, const float* bp, const float* wg,
                                     const float* bg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < EMB * PROJ_PAD) {
        const int o = i % PROJ_PAD, j = i / PROJ_PAD;
        float v = 0.f;
        if (o < N_MELS) v = wp[(size_t)o * (RNN + EMB) + RNN + j];
        else if (o == N_MELS) v = wg[RNN + j];
        wT[i] = v;
    }
    if (i < PROJ_PAD) bias[i] = i < N_MELS ? bp[i] : (i == N_MELS ? bg[0] : 0.f);
}

// dst[c][r] = src[r][c]   (src is R x C row-major)
__global__ void transpose_kernel(float* dst, const float* src, int R, int C) {
    const int total = R * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i % R, c = i / R;
        dst[i] = src[(size_t)r * C + c];
    }
}

__global__ void copy_kernel(float* dst, const float* src, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void pack_attn_v_kernel(float* dst, const float* vw, const float* vb, const float* scale) {
    const int i = threadIdx.x;
    if (i < ATT_DIM) dst[i] = vw[i];
    if (i == ATT_DIM) dst[i] = vb[0];
    if (i == ATT_DIM + 1) dst[i] = scale[0];
}

// Conv1d weight (co, ci, k) -> [k][ci][co]; BatchNorm (eval) folded with the conv bias into
// y = conv * scale + shift  (src/model.py:25-39, :347-390).
__global__ void pack_conv_kernel(float* dst, const float* w, int c_out, int c_in, int taps) {
    const int total = c_out * c_in * taps;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % c_out, ci = (i / c_out) % c_in, k = i / (c_out * c_in);
        dst[i] = w[((size_t)co * c_in + ci) * taps + k];
    }
}
__global__ void pack_bn_kernel(float* scale, float* shift, const float* conv_b, const float* g, const float* b,
                               const float* mean, const float* var, int c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) {
        const float s = g[i] / sqrtf(var[i] + BN_EPS);
        scale[i] = s;
        shift[i] = b[i] + (conv_b[i] - mean[i]) * s;
    }
}

// encoder BiLSTM: input weights of both directions side by side [512][2048], summed biases
__global__ void pack_enc_wih_kernel(float* dst, float* bias, const float* w0, const float* w1, const float* bi0,
                                    const float* bh0, const float* bi1, const float* bh1) {
    const int total = EMB * 8 * ENC_H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int col = i % (8 * ENC_H), ci = i / (8 * ENC_H);
        const int d = col / (4 * ENC_H), r = col % (4 * ENC_H);
        dst[i] = (d ? w1 : w0)[(size_t)r * EMB + ci];
        if (ci == 0) bias[col] = d ? bi1[r] + bh1[r] : bi0[r] + bh0[r];
    }
}
// recurrent weights: [dir][rank][row = unit*4+gate][k], source row = gate*256 + 32*rank + unit
__global__ void pack_enc_whh_kernel(float* dst, const float* w0, const float* w1) {
    const int total = 2 * 4 * ENC_H * ENC_H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i % ENC_H, rl = (i / ENC_H) % 128, rank = (i / (ENC_H * 128)) % 8, d = i / (ENC_H * 128 * 8);
        const int row = (rl & 3) * ENC_H + rank * 32 + (rl >> 2);
        dst[i] = (d ? w1 : w0)[(size_t)row * ENC_H + k];
    }
}

struct Arena {
    uint8_t* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off = align_up(off + n * sizeof(T), 256);
        return p;
    }
};

// Single source of truth for the arena layout; with base == nullptr it only measures.
size_t carve(b200taco_handle* h, uint8_t* base, int weight_dtype) {
    Arena A{base};
    const size_t wb = weight_dtype == 1 ? 4 : 2;
    void* w_attn = A.take<uint8_t>((size_t)GEMM_CTAS * K_ATTN * ROWS_PER_CTA * wb);
    void* w_dec = A.take<uint8_t>((size_t)GEMM_CTAS * K_DEC * ROWS_PER_CTA * wb);
    void* w_attn_tc = A.take<uint8_t>(weight_dtype == 0 ? (size_t)GEMM_CTAS * K_ATTN * ROWS_PER_CTA * 2 : 0);
    void* w_dec_tc = A.take<uint8_t>(weight_dtype == 0 ? (size_t)GEMM_CTAS * K_DEC * ROWS_PER_CTA * 2 : 0);
    float* bias_attn = A.take<float>(GEMM_CTAS * ROWS_PER_CTA);
    float* bias_dec = A.take<float>(GEMM_CTAS * ROWS_PER_CTA);
    float* wq = A.take<float>(GEMM_CTAS * ATT_DIM * UNITS_PER_CTA);
    float* wproj = A.take<float>(GEMM_CTAS * PROJ_OUT * UNITS_PER_CTA);
    float* wctx = A.take<float>(EMB * PROJ_PAD);
    float* pbias = A.take<float>(PROJ_PAD);
    float* p1 = A.take<float>(N_MELS * PRENET);
    float* p2 = A.take<float>(PRENET * PRENET);
    float* lc = A.take<float>(LOC_F * 2 * LOC_K);
    float* ld = A.take<float>(LOC_F * ATT_DIM);
    float* av = A.take<float>(ATT_DIM + 4);
    float* emb = A.take<float>(N_SYMBOLS * EMB);
    float *ecw[3], *ecs[3], *ech[3];
    for (int i = 0; i < 3; ++i) {
        ecw[i] = A.take<float>((size_t)ENC_KERNEL * EMB * EMB);
        ecs[i] = A.take<float>(EMB);
        ech[i] = A.take<float>(EMB);
    }
    float* wih = A.take<float>((size_t)EMB * 8 * ENC_H);
    float* lb = A.take<float>(8 * ENC_H);
    float* whh = A.take<float>((size_t)2 * 4 * ENC_H * ENC_H);
    float* mw = A.take<float>(EMB * ATT_DIM);
    float *pcw[5], *pcs[5], *pch[5];
    for (int i = 0; i < 5; ++i) {
        const int ci = i == 0 ? N_MELS : POST_DIM, co = i == 4 ? N_MELS : POST_DIM;
        pcw[i] = A.take<float>((size_t)POST_KERNEL * ci * co);
        pcs[i] = A.take<float>(co);
        pch[i] = A.take<float>(co);
    }
    float *ecwt[3], *pcwt[5];
    for (int i = 0; i < 3; ++i) ecwt[i] = A.take<float>(dense_tc_weight_floats(ENC_KERNEL, EMB, EMB));
    float* wih_tc = A.take<float>(dense_tc_weight_floats(1, EMB, 8 * ENC_H));
    float* mw_tc = A.take<float>(dense_tc_weight_floats(1, EMB, ATT_DIM));
    float* p1_tc = A.take<float>(dense_tc_weight_floats(1, N_MELS, PRENET));
    float* p2_tc = A.take<float>(dense_tc_weight_floats(1, PRENET, PRENET));
    for (int i = 0; i < 5; ++i) {
        const int ci = i == 0 ? N_MELS : POST_DIM, co = i == 4 ? N_MELS : POST_DIM;
        pcwt[i] = A.take<float>(dense_tc_weight_floats(POST_KERNEL, ci, co));
    }
    if (h && base) {
        for (int i = 0; i < 3; ++i) h->enc_conv_wtc[i] = ecwt[i];
        for (int i = 0; i < 5; ++i) h->post_conv_wtc[i] = pcwt[i];
        h->enc_lstm_wih_tc = wih_tc; h->attn_memory_wtc = mw_tc; h->prenet1_tc = p1_tc; h->prenet2_tc = p2_tc;
        h->w_attn_tc = weight_dtype == 0 ? w_attn_tc : nullptr;
        h->w_dec_tc = weight_dtype == 0 ? w_dec_tc : nullptr;
        h->w_attn = w_attn; h->w_dec = w_dec; h->bias_attn = bias_attn; h->bias_dec = bias_dec;
        h->wq_slices = wq; h->wproj_slices = wproj; h->wproj_ctxT = wctx; h->proj_bias = pbias;
        h->prenet1T = p1; h->prenet2T = p2; h->loc_conv = lc; h->loc_denseT = ld; h->attn_v = av;
        h->enc_emb = emb;
        for (int i = 0; i < 3; ++i) { h->enc_conv_w[i] = ecw[i]; h->enc_conv_scale[i] = ecs[i]; h->enc_conv_shift[i] = ech[i]; }
        h->enc_lstm_wih = wih; h->enc_lstm_bias = lb; h->enc_lstm_whh = whh; h->attn_memory_wT = mw;
        for (int i = 0; i < 5; ++i) { h->post_conv_w[i] = pcw[i]; h->post_conv_scale[i] = pcs[i]; h->post_conv_shift[i] = pch[i]; }
    }
    return A.off;
}

inline int blocks_for(size_t n) { return (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256); }

}  // namespace

size_t packed_arena_bytes(int weight_dtype) { return carve(nullptr, nullptr, weight_dtype); }

int pack_weights(b200taco_handle* h, const b200taco_weights* w, cudaStream_t st) {
    const size_t need = packed_arena_bytes(h->weight_dtype);
    if (h->arena == nullptr || h->arena_bytes != need) {
        if (h->arena) TACO_CUDA(cudaFree(h->arena));
        h->arena = nullptr;
        TACO_CUDA(cudaMalloc(&h->arena, need));
        h->arena_bytes = need;
    }
    carve(h, static_cast<uint8_t*>(h->arena), h->weight_dtype);
    h->packed = false;

    // ---- decoder LSTMs: act row order is [prenet | ctx | ah | dh] (dims.cuh)
    // attention LSTM input = cat[prenet_out, context] (src/model.py:256), then h
    // decoder LSTM input   = cat[attention_hidden, context] (src/model.py:265): reordered to [ctx | ah | dh]
    const size_t na = (size_t)GEMM_CTAS * K_ATTN * ROWS_PER_CTA, nd = (size_t)GEMM_CTAS * K_DEC * ROWS_PER_CTA;
    if (h->weight_dtype == 1) {
        pack_lstm_kernel<float><<<blocks_for(na), 256, 0, st>>>(
            (float*)h->w_attn, K_ATTN, w->attn_lstm_w_ih, PRENET + EMB, 0, PRENET + EMB, w->attn_lstm_w_hh, RNN, 0, RNN,
            w->attn_lstm_w_hh, RNN, 0);
        TACO_LAUNCHED();
        pack_lstm_kernel<float><<<blocks_for(nd), 256, 0, st>>>(
            (float*)h->w_dec, K_DEC, w->dec_lstm_w_ih, RNN + EMB, RNN, EMB, w->dec_lstm_w_ih, RNN + EMB, 0, RNN,
            w->dec_lstm_w_hh, RNN, 0);
        TACO_LAUNCHED();
    } else {
        pack_lstm_kernel<__nv_bfloat16><<<blocks_for(na), 256, 0, st>>>(
            (__nv_bfloat16*)h->w_attn, K_ATTN, w->attn_lstm_w_ih, PRENET + EMB, 0, PRENET + EMB, w->attn_lstm_w_hh, RNN,
            0, RNN, w->attn_lstm_w_hh, RNN, 0);
        TACO_LAUNCHED();
        pack_lstm_kernel<__nv_bfloat16><<<blocks_for(nd), 256, 0, st>>>(
            (__nv_bfloat16*)h->w_dec, K_DEC, w->dec_lstm_w_ih, RNN + EMB, RNN, EMB, w->dec_lstm_w_ih, RNN + EMB, 0, RNN,
            w->dec_lstm_w_hh, RNN, 0);
        TACO_LAUNCHED();
        pack_lstm_tc_kernel<<<blocks_for(na), 256, 0, st>>>(
            (__nv_bfloat16*)h->w_attn_tc, K_ATTN, 0, w->attn_lstm_w_ih, PRENET + EMB, 0, PRENET + EMB, w->attn_lstm_w_hh, RNN,
            0, RNN, w->attn_lstm_w_hh, RNN, 0);
        TACO_LAUNCHED();
        pack_lstm_tc_kernel<<<blocks_for(nd), 256, 0, st>>>(
            (__nv_bfloat16*)h->w_dec_tc, K_DEC, 1, w->dec_lstm_w_ih, RNN + EMB, RNN, EMB, w->dec_lstm_w_ih, RNN + EMB, 0, RNN,
            w->dec_lstm_w_hh, RNN, 0);
        TACO_LAUNCHED();
    }
    pack_lstm_bias_kernel<<<16, 256, 0, st>>>(h->bias_attn, w->attn_lstm_b_ih, w->attn_lstm_b_hh);
    TACO_LAUNCHED();
    pack_lstm_bias_kernel<<<16, 256, 0, st>>>(h->bias_dec, w->dec_lstm_b_ih, w->dec_lstm_b_hh);
    TACO_LAUNCHED();
    pack_slices_kernel<<<128, 256, 0, st>>>(h->wq_slices, ATT_DIM, w->attn_query_w, ATT_DIM, w->attn_query_w, RNN);
    TACO_LAUNCHED();
    pack_slices_kernel<<<128, 256, 0, st>>>(h->wproj_slices, PROJ_OUT, w->proj_w, N_MELS, w->gate_w, RNN + EMB);
    TACO_LAUNCHED();
    pack_proj_ctx_kernel<<<blocks_for(EMB * PROJ_PAD), 256, 0, st>>>(h->wproj_ctxT, h->proj_bias, w->proj_w, w->proj_b,
                                                                     w->gate_w, w->gate_b);
    TACO_LAUNCHED();
    transpose_kernel<<<64, 256, 0, st>>>(h->prenet1T, w->prenet_w[0], PRENET, N_MELS);   // (256,80) -> [80][256]
    TACO_LAUNCHED();
    transpose_kernel<<<64, 256, 0, st>>>(h->prenet2T, w->prenet_w[1], PRENET, PRENET);
    TACO_LAUNCHED();
    copy_kernel<<<8, 256, 0, st>>>(h->loc_conv, w->attn_loc_conv_w, LOC_F * 2 * LOC_K);
    TACO_LAUNCHED();
    transpose_kernel<<<16, 256, 0, st>>>(h->loc_denseT, w->attn_loc_dense_w, ATT_DIM, LOC_F);  // (128,32) -> [32][128]
    TACO_LAUNCHED();
    pack_attn_v_kernel<<<1, 256, 0, st>>>(h->attn_v, w->attn_v_w, w->attn_v_b, w->attn_energy_scale);
    TACO_LAUNCHED();
    transpose_kernel<<<64, 256, 0, st>>>(h->attn_memory_wT, w->attn_memory_w, ATT_DIM, EMB);    // (128,512) -> [512][128]
    TACO_LAUNCHED();

    // ---- encoder
    copy_kernel<<<64, 256, 0, st>>>(h->enc_emb, w->enc_embedding, N_SYMBOLS * EMB);
    TACO_LAUNCHED();
    for (int i = 0; i < 3; ++i) {
        pack_conv_kernel<<<1024, 256, 0, st>>>(h->enc_conv_w[i], w->enc_conv_w[i], EMB, EMB, ENC_KERNEL);
        TACO_LAUNCHED();
        pack_bn_kernel<<<2, 256, 0, st>>>(h->enc_conv_scale[i], h->enc_conv_shift[i], w->enc_conv_b[i], w->enc_bn_w[i],
                                          w->enc_bn_b[i], w->enc_bn_mean[i], w->enc_bn_var[i], EMB);
        TACO_LAUNCHED();
    }
    pack_enc_wih_kernel<<<1024, 256, 0, st>>>(h->enc_lstm_wih, h->enc_lstm_bias, w->enc_lstm_w_ih[0], w->enc_lstm_w_ih[1],
                                              w->enc_lstm_b_ih[0], w->enc_lstm_b_hh[0], w->enc_lstm_b_ih[1],
                                              w->enc_lstm_b_hh[1]);
    TACO_LAUNCHED();
    pack_enc_whh_kernel<<<1024, 256, 0, st>>>(h->enc_lstm_whh, w->enc_lstm_w_hh[0], w->enc_lstm_w_hh[1]);
    TACO_LAUNCHED();

    // ---- postnet
    for (int i = 0; i < 5; ++i) {
        const int ci = i == 0 ? N_MELS : POST_DIM, co = i == 4 ? N_MELS : POST_DIM;
        pack_conv_kernel<<<1024, 256, 0, st>>>(h->post_conv_w[i], w->post_conv_w[i], co, ci, POST_KERNEL);
        TACO_LAUNCHED();
        pack_bn_kernel<<<2, 256, 0, st>>>(h->post_conv_scale[i], h->post_conv_shift[i], w->post_conv_b[i], w->post_bn_w[i],
                                          w->post_bn_b[i], w->post_bn_mean[i], w->post_bn_var[i], co);
        TACO_LAUNCHED();
    }
    // ---- tensor-core copies of the dense weights (TF32 big/small planes)
    for (int i = 0; i < 3; ++i)
        if (int e = pack_dense_tc(h->enc_conv_wtc[i], h->enc_conv_w[i], ENC_KERNEL, EMB, EMB, st)) return e;
    if (int e = pack_dense_tc(h->enc_lstm_wih_tc, h->enc_lstm_wih, 1, EMB, 8 * ENC_H, st)) return e;
    if (int e = pack_dense_tc(h->attn_memory_wtc, h->attn_memory_wT, 1, EMB, ATT_DIM, st)) return e;
    if (int e = pack_dense_tc(h->prenet1_tc, h->prenet1T, 1, N_MELS, PRENET, st)) return e;
    if (int e = pack_dense_tc(h->prenet2_tc, h->prenet2T, 1, PRENET, PRENET, st)) return e;
    for (int i = 0; i < 5; ++i) {
        const int ci = i == 0 ? N_MELS : POST_DIM, co = i == 4 ? N_MELS : POST_DIM;
        if (int e = pack_dense_tc(h->post_conv_wtc[i], h->post_conv_w[i], POST_KERNEL, ci, co, st)) return e;
    }
    h->packed = true;
    return B200TACO_OK;
}

}  // namespace b200taco
