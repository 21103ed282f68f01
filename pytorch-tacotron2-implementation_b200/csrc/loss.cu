// Guided-attention target of the reference's Tacotron2Loss (train.py:148-171): for every (utterance b, decoder step t) a
// Gaussian over the phoneme positions p < L_b centred on the diagonal position  e = min(int(t * L_b / S), L_b - 1), width
// sigma_b, normalised by (sum + 1e-8); zero for p >= L_b.  The reference builds it with a Python double loop over B x S
// (51,200 iterations at B=64, S=800: seconds per training step); here it is one CTA per (b, t).
#include "handle.h"

namespace b200taco {
namespace {

constexpr int GA_THREADS = 128;

__global__ void __launch_bounds__(GA_THREADS) guided_attention_target_kernel(const int64_t* __restrict__ text_lengths,
                                                                            const float* __restrict__ sigma,
                                                                            float* __restrict__ target, int S, int Tmax) {
    __shared__ float g_s[MAX_TEXT_LEN];
    __shared__ float red_s[GA_THREADS / 32];
    const int b = blockIdx.x / S, t = blockIdx.x % S, tid = threadIdx.x;
    long long L = text_lengths[b];
    L = L < 1 ? 1 : (L > Tmax ? Tmax : L);
    const long long e_ll = ((long long)t * L) / S;                       // int(t * Lb / num_steps), train.py:166
    const float e = (float)(e_ll < L - 1 ? e_ll : L - 1);
    const float sg = sigma[b];
    float part = 0.f;
    for (int p = tid; p < (int)L; p += GA_THREADS) {
        const float x = ((float)p - e) / sg;
        const float g = expf(-0.5f * (x * x));                           // train.py:167
        g_s[p] = g;
        part += g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red_s[tid >> 5] = part;
    __syncthreads();
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < GA_THREADS / 32; ++i) total += red_s[i];
    const float denom = total + 1e-8f;                                   // train.py:168
    float* row = target + ((size_t)b * S + t) * Tmax;
    for (int p = tid; p < Tmax; p += GA_THREADS) row[p] = p < (int)L ? g_s[p] / denom : 0.f;
}


// =====================================================================================================================
// Whole Tacotron2Loss.forward (train.py:173-217) without host round trips: masked L1 of both mel predictions, BCE of the
// gate logits, KL(guided target || attention) / T_dec, mean attention entropy, the entropy-adaptive weight of the KL term
// and the total, left as DEVICE scalars.  The guided target is recomputed on the fly (never materialised).
//   out[0] mel loss  out[1] gate loss  out[2] KL (clamped to 150)  out[3] entropy  out[4] attention weight
//   out[5] total     out[6] KL before the clamp                    out[7] number of valid mel bins
constexpr int LT = 256;                  // threads of the mel / gate kernels
constexpr int LS = 32;                   // frames per CTA

__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)blockDim.x / 32; ++i) t += red[i];
    return t;
}

// MODE 0: partial sums {|coarse - tgt|, |post - tgt|, bce} per CTA;  MODE 1: gradients
template <int MODE>
__global__ void __launch_bounds__(LT) loss_mel_gate_kernel(const float* __restrict__ post, const float* __restrict__ coarse,
                                                           const float* __restrict__ target, const float* __restrict__ gate,
                                                           const float* __restrict__ gate_target, const int64_t* __restrict__ mel_lengths,
                                                           int B, int S, float* __restrict__ partial, const float* __restrict__ outs,
                                                           const float* __restrict__ gout, float* __restrict__ d_post,
                                                           float* __restrict__ d_coarse, float* __restrict__ d_gate) {
    __shared__ float tgt_s[N_MELS][LS + 1];
    __shared__ float red[LT / 32];
    const int tid = threadIdx.x, b = blockIdx.y, s0 = blockIdx.x * LS;
    const int ns = min(LS, S - s0);
    long long len = mel_lengths[b];
    const int valid = (int)(len < 0 ? 0 : (len > S ? S : len));                   // frames s < valid count (train.py:143-147)
    for (int i = tid; i < N_MELS * LS; i += LT) {
        const int m = i / LS, sl = i % LS;
        tgt_s[m][sl] = sl < ns ? target[((size_t)b * N_MELS + m) * S + s0 + sl] : 0.f;      // (B,80,S): coalesced along s
    }
    __syncthreads();
    float c_mel = 0.f, c_gate = 0.f;
    if (MODE == 1) {
        c_mel = (gout[5] + gout[0]) / outs[7];
        c_gate = (gout[5] + gout[1]) / (float)((long long)B * S);
    }
    float l1c = 0.f, l1p = 0.f, bce = 0.f;
    for (int i = tid; i < ns * N_MELS; i += LT) {
        const int sl = i / N_MELS, m = i % N_MELS;
        const size_t o = ((size_t)b * S + s0 + sl) * N_MELS + m;
        const bool keep = s0 + sl < valid;
        const float t = tgt_s[m][sl], dc = coarse[o] - t, dp = post[o] - t;
        if (MODE == 0) {
            if (keep) { l1c += fabsf(dc); l1p += fabsf(dp); }
        } else {
            d_coarse[o] = keep ? c_mel * (dc > 0.f ? 1.f : (dc < 0.f ? -1.f : 0.f)) : 0.f;
            d_post[o] = keep ? c_mel * (dp > 0.f ? 1.f : (dp < 0.f ? -1.f : 0.f)) : 0.f;
        }
    }
    if (tid < ns) {
        const size_t o = (size_t)b * S + s0 + tid;
        const float x = gate[o], y = gate_target[o];
        if (MODE == 0) bce = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));       // BCEWithLogitsLoss, train.py:194
        else d_gate[o] = c_gate * (sigmoid_acc(x) - y);
    }
    if (MODE == 0) {
        l1c = block_sum(l1c, red, tid);
        l1p = block_sum(l1p, red, tid);
        bce = block_sum(bce, red, tid);
        if (tid == 0) {
            float* q = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3;
            q[0] = l1c; q[1] = l1p; q[2] = bce;
        }
    }
}

// one warp per (decoder step s, utterance b) row of the alignments (S,B,T).  MODE 0: {KL sum, entropy sum} per CTA;
// MODE 1: d loss / d alignment
constexpr int AW = 4;                    // warps (rows) per CTA
template <int MODE>
__global__ void __launch_bounds__(AW * 32) loss_attention_kernel(const float* __restrict__ align, const int64_t* __restrict__ text_lengths,
                                                                 const float* __restrict__ sigma, int B, int S, int T,
                                                                 float* __restrict__ partial, const float* __restrict__ outs,
                                                                 const float* __restrict__ gout, float* __restrict__ d_align) {
    __shared__ float red[2][AW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * AW + warp;
    float kl = 0.f, ent = 0.f;
    if (row < (long long)S * B) {
        const int s = (int)(row / B), b = (int)(row % B);
        long long L = text_lengths[b];
        L = L < 1 ? 1 : (L > T ? T : L);
        const long long e_ll = ((long long)s * L) / S;                               // int(t * Lb / num_steps), train.py:166
        const float e = (float)(e_ll < L - 1 ? e_ll : L - 1), sg = sigma[b];
        float part = 0.f;
        for (int p = lane; p < (int)L; p += 32) {
            const float x = ((float)p - e) / sg;
            part += expf(-0.5f * (x * x));
        }
        const float denom = warp_sum(part) + 1e-8f;                                  // train.py:168
        const float* a_row = align + (size_t)row * T;
        float coef = 0.f;
        if (MODE == 1) coef = outs[6] <= 150.0f ? -(gout[5] * outs[4] + gout[2]) / ((float)B * (float)S) : 0.f;     // clamp(max=150)
        for (int p = lane; p < T; p += 32) {
            const float a = a_row[p], a_safe = fmaxf(a, 1e-8f), lg = logf(a_safe);    // clamp_min(1e-8).log(), train.py:200-201
            float tg = 0.f;
            if (p < (int)L) {
                const float x = ((float)p - e) / sg;
                tg = expf(-0.5f * (x * x)) / denom;
            }
            if (MODE == 0) {
                if (tg > 0.f) kl += tg * (logf(tg) - lg);                            // F.kl_div pointwise: xlogy(t, t) - t * input
                ent -= a_safe * lg;
            } else {
                d_align[(size_t)row * T + p] = a >= 1e-8f ? coef * tg / a_safe : 0.f;
            }
        }
    }
    if (MODE == 0) {
        kl = warp_sum(kl);
        ent = warp_sum(ent);
        if (lane == 0) { red[0][warp] = kl; red[1][warp] = ent; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float k = 0.f, en = 0.f;
            for (int i = 0; i < AW; ++i) { k += red[0][i]; en += red[1][i]; }
            partial[(size_t)blockIdx.x * 2] = k;
            partial[(size_t)blockIdx.x * 2 + 1] = en;
        }
    }
}

// fixed-order (deterministic) reduction of the partials and the scalar logic of train.py:189-214
__global__ void __launch_bounds__(256) loss_finish_kernel(const float* __restrict__ pmel, int n_mel, const float* __restrict__ patt, int n_att,
                                                          const int64_t* __restrict__ mel_lengths, int B, int S, int T, int use_attention,
                                                          float w_start, float w_min, float ent_target, float w_current,
                                                          float* __restrict__ out) {
    __shared__ double red[5][256];
    const int tid = threadIdx.x;
    double a[5] = {0, 0, 0, 0, 0};
    for (int i = tid; i < n_mel; i += 256) { a[0] += pmel[3 * i]; a[1] += pmel[3 * i + 1]; a[2] += pmel[3 * i + 2]; }
    for (int i = tid; i < n_att; i += 256) { a[3] += patt[2 * i]; a[4] += patt[2 * i + 1]; }
    for (int k = 0; k < 5; ++k) red[k][tid] = a[k];
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (tid < st)
            for (int k = 0; k < 5; ++k) red[k][tid] += red[k][tid + st];
        __syncthreads();
    }
    if (tid == 0) {
        double valid = 0;
        for (int b = 0; b < B; ++b) {
            long long l = mel_lengths[b];
            valid += (double)(l < 0 ? 0 : (l > S ? S : l)) * N_MELS;
        }
        const float mel = (float)(red[0][0] / valid) + (float)(red[1][0] / valid);
        const float gate = (float)(red[2][0] / ((double)B * S));
        float kl_raw = 0.f, kl = 0.f, ent = 0.f, w = w_current;
        if (use_attention) {
            kl_raw = (float)(red[3][0] / B / S);                                  // batchmean / T_dec, train.py:202
            kl = fminf(kl_raw, 150.0f);
            ent = (float)(red[4][0] / ((double)B * S));
            w = ent <= ent_target ? fmaxf(w_min, w_start * fmaxf(ent / ent_target, 0.f)) : w_start;      // train.py:205-210
        }
        out[0] = mel; out[1] = gate; out[2] = kl; out[3] = ent; out[4] = w;
        out[5] = mel + gate + w * kl;                                              // train.py:215
        out[6] = kl_raw; out[7] = (float)valid;
    }
}

// gate targets of the training loop (train.py:506-508): 1 from frame mel_lengths[b] - 1 on
__global__ void gate_target_kernel(const int64_t* __restrict__ mel_lengths, int B, int S, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * S) return;
    const int b = (int)(i / S), s = (int)(i % S);
    long long first = mel_lengths[b] - 1;
    if (first < 0) first += S;                                                     // Python slice semantics of a negative start
    out[i] = s >= first ? 1.f : 0.f;
}

}  // namespace
}  // namespace b200taco

using namespace b200taco;

extern "C" {

B200TACO_API int32_t b200taco_guided_attention_target(const int64_t* d_text_lengths, const float* d_sigma, int32_t B, int32_t S,
                                                      int32_t Tmax, float* d_target, void* stream) {
    TACO_REQUIRE(d_text_lengths && d_sigma && d_target, B200TACO_ERR_INVALID, "guided_attention_target: null pointer");
    TACO_REQUIRE(B >= 1 && S >= 1 && Tmax >= 1 && Tmax <= MAX_TEXT_LEN, B200TACO_ERR_UNSUPPORTED,
                 "guided_attention_target: B %d, S %d, Tmax %d (max %d)", B, S, Tmax, MAX_TEXT_LEN);
    TACO_REQUIRE((long long)B * S < (1ll << 31), B200TACO_ERR_UNSUPPORTED, "guided_attention_target: B * S too large");
    guided_attention_target_kernel<<<B * S, GA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d_text_lengths, d_sigma, d_target,
                                                                                               S, Tmax);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // extern "C"

extern "C" {

B200TACO_API int64_t b200taco_loss_scratch_bytes(int32_t B, int32_t S) {
    const long long n_mel = (long long)((S + LS - 1) / LS) * B, n_att = ((long long)S * B + AW - 1) / AW;
    return (int64_t)((n_mel * 3 + n_att * 2) * sizeof(float));
}

B200TACO_API int32_t b200taco_loss_fwd(const float* d_mel_post, const float* d_mel_coarse, const float* d_mel_target,
                                       const float* d_gate, const float* d_gate_target, const int64_t* d_mel_lengths,
                                       const float* d_align, const int64_t* d_text_lengths, const float* d_sigma, int32_t B,
                                       int32_t S, int32_t T, float w_start, float w_min, float entropy_target, float w_current,
                                       float* d_scratch, float* d_out8, void* stream) {
    TACO_REQUIRE(d_mel_post && d_mel_coarse && d_mel_target && d_gate && d_gate_target && d_mel_lengths && d_scratch && d_out8,
                 B200TACO_ERR_INVALID, "loss_fwd: null pointer");
    TACO_REQUIRE(B >= 1 && B <= 65535 && S >= 1, B200TACO_ERR_UNSUPPORTED, "loss_fwd: B %d, S %d", B, S);
    const bool att = d_align != nullptr;
    TACO_REQUIRE(!att || (d_text_lengths && d_sigma && T >= 1), B200TACO_ERR_INVALID, "loss_fwd: attention term needs lengths, sigma, T");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int gx = (S + LS - 1) / LS, n_mel = gx * B;
    const long long n_att = ((long long)S * B + AW - 1) / AW;
    float* pmel = d_scratch;
    float* patt = d_scratch + (size_t)n_mel * 3;
    loss_mel_gate_kernel<0><<<dim3(gx, B), LT, 0, st>>>(d_mel_post, d_mel_coarse, d_mel_target, d_gate, d_gate_target, d_mel_lengths, B, S,
                                                        pmel, nullptr, nullptr, nullptr, nullptr, nullptr);
    TACO_LAUNCHED();
    if (att) {
        loss_attention_kernel<0><<<(unsigned)n_att, AW * 32, 0, st>>>(d_align, d_text_lengths, d_sigma, B, S, T, patt, nullptr, nullptr, nullptr);
        TACO_LAUNCHED();
    }
    loss_finish_kernel<<<1, 256, 0, st>>>(pmel, n_mel, patt, att ? (int)n_att : 0, d_mel_lengths, B, S, T, att ? 1 : 0, w_start, w_min,
                                          entropy_target, w_current, d_out8);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_loss_bwd(const float* d_mel_post, const float* d_mel_coarse, const float* d_mel_target,
                                       const float* d_gate, const float* d_gate_target, const int64_t* d_mel_lengths,
                                       const float* d_align, const int64_t* d_text_lengths, const float* d_sigma, int32_t B,
                                       int32_t S, int32_t T, const float* d_out8, const float* d_grad_out8, float* d_g_post,
                                       float* d_g_coarse, float* d_g_gate, float* d_g_align, void* stream) {
    TACO_REQUIRE(d_mel_post && d_mel_coarse && d_mel_target && d_gate && d_gate_target && d_mel_lengths && d_out8 && d_grad_out8 &&
                     d_g_post && d_g_coarse && d_g_gate,
                 B200TACO_ERR_INVALID, "loss_bwd: null pointer");
    TACO_REQUIRE(B >= 1 && B <= 65535 && S >= 1, B200TACO_ERR_UNSUPPORTED, "loss_bwd: B %d, S %d", B, S);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int gx = (S + LS - 1) / LS;
    loss_mel_gate_kernel<1><<<dim3(gx, B), LT, 0, st>>>(d_mel_post, d_mel_coarse, d_mel_target, d_gate, d_gate_target, d_mel_lengths, B, S,
                                                        nullptr, d_out8, d_grad_out8, d_g_post, d_g_coarse, d_g_gate);
    TACO_LAUNCHED();
    if (d_align && d_g_align) {
        TACO_REQUIRE(d_text_lengths && d_sigma && T >= 1, B200TACO_ERR_INVALID, "loss_bwd: attention term needs lengths, sigma, T");
        const long long n_att = ((long long)S * B + AW - 1) / AW;
        loss_attention_kernel<1><<<(unsigned)n_att, AW * 32, 0, st>>>(d_align, d_text_lengths, d_sigma, B, S, T, nullptr, d_out8, d_grad_out8,
                                                                      d_g_align);
        TACO_LAUNCHED();
    }
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_gate_targets(const int64_t* d_mel_lengths, int32_t B, int32_t S, float* d_out, void* stream) {
    TACO_REQUIRE(d_mel_lengths && d_out && B >= 1 && S >= 1, B200TACO_ERR_INVALID, "gate_targets: bad arguments");
    const long long n = (long long)B * S;
    gate_target_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_mel_lengths, B, S, d_out);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // extern "C"
