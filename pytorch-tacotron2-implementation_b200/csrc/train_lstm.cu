// Training-mode recurrent kernels (first correct fp32 versions).
//
// Encoder BiLSTM (reference nn.LSTM, 1 layer, bidirectional, hidden 256, zero initial state, full padded length,
// src/model.py:42-48, :75): forward that saves the gate activations and states of every position, and the
// backward through time.  One CTA per (direction, utterance); the recurrent matrix streams from L2 every step
// (transposed copy for the forward mat-vec, the raw parameter for the transposed mat-vec of the backward).
//   saved per (b, t, dir): i, f, g, o (post-activation, 4 x 256) | c (256) | h (256)      = LSTM_SAVE floats
#include "handle.h"

namespace b200taco {

constexpr int LSTM_SAVE = 6 * ENC_H;

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// whhT [256 k][1024 r] (r = gate*256 + unit, PyTorch order i,f,g,o); xproj [B*T][2048] = x W_ih^T + b_ih + b_hh (fwd | bwd)
__global__ void __launch_bounds__(1024) bilstm_train_fwd_kernel(const float* __restrict__ whhT_f, const float* __restrict__ whhT_b,
                                                                const float* __restrict__ xproj, float* __restrict__ out,
                                                                float* __restrict__ saved, int B, int T) {
    const int dir = blockIdx.x / B, b = blockIdx.x % B;
    const float* whhT = dir ? whhT_b : whhT_f;
    const int r = threadIdx.x;                       // gate row
    __shared__ float h_s[ENC_H];
    __shared__ float z_s[4 * ENC_H];
    float c = 0.f;                                   // cell state of unit r (threads r < 256)
    if (r < ENC_H) h_s[r] = 0.f;
    __syncthreads();
    for (int step = 0; step < T; ++step) {
        const int t = dir ? T - 1 - step : step;
        float acc = xproj[((size_t)b * T + t) * (8 * ENC_H) + dir * 4 * ENC_H + r];
#pragma unroll 8
        for (int k = 0; k < ENC_H; ++k) acc = fmaf(__ldg(whhT + (size_t)k * 4 * ENC_H + r), h_s[k], acc);
        const int gate = r / ENC_H;
        z_s[r] = gate == 2 ? tanhf(acc) : sigmoidf_(acc);
        __syncthreads();
        float* sv = saved + (((size_t)b * T + t) * 2 + dir) * LSTM_SAVE;
        sv[r] = z_s[r];
        if (r < ENC_H) {
            c = z_s[ENC_H + r] * c + z_s[r] * z_s[2 * ENC_H + r];
            const float h = z_s[3 * ENC_H + r] * tanhf(c);
            sv[4 * ENC_H + r] = c;
            sv[5 * ENC_H + r] = h;
            out[((size_t)b * T + t) * EMB + dir * ENC_H + r] = h;
            h_s[r] = h;
        }
        __syncthreads();
    }
}

// d_out [B][T][512] -> dz [B*T][2048] (gradient w.r.t. the gate pre-activations = w.r.t. xproj), hprev [B*T][2][256]
// (h_{t-1} of each direction, the left operand of the recurrent weight gradient).  whh = raw parameters (1024 x 256).
__global__ void __launch_bounds__(1024) bilstm_train_bwd_kernel(const float* __restrict__ whh_f, const float* __restrict__ whh_b,
                                                                const float* __restrict__ d_out, const float* __restrict__ saved,
                                                                float* __restrict__ dz_out, float* __restrict__ hprev, int B,
                                                                int T) {
    const int dir = blockIdx.x / B, b = blockIdx.x % B;
    const float* whh = dir ? whh_b : whh_f;
    const int r = threadIdx.x, k = r & (ENC_H - 1), part = r >> 8;
    __shared__ float dz_s[4 * ENC_H];
    __shared__ float part_s[4][ENC_H];
    __shared__ float dh_s[ENC_H];
    float dc = 0.f;                                   // dL/dc carried to the previous position (threads r < 256)
    if (r < ENC_H) dh_s[r] = 0.f;
    __syncthreads();
    for (int step = T - 1; step >= 0; --step) {      // reverse of the forward order of this direction
        const int t = dir ? T - 1 - step : step;
        const int tp = dir ? t + 1 : t - 1;          // position processed just before t in the forward pass
        const bool has_prev = step > 0;
        const float* sv = saved + (((size_t)b * T + t) * 2 + dir) * LSTM_SAVE;
        if (r < ENC_H) {
            const float i = sv[r], f = sv[ENC_H + r], g = sv[2 * ENC_H + r], o = sv[3 * ENC_H + r], c = sv[4 * ENC_H + r];
            const float cprev = has_prev ? saved[(((size_t)b * T + tp) * 2 + dir) * LSTM_SAVE + 4 * ENC_H + r] : 0.f;
            const float hp = has_prev ? saved[(((size_t)b * T + tp) * 2 + dir) * LSTM_SAVE + 5 * ENC_H + r] : 0.f;
            const float dh = d_out[((size_t)b * T + t) * EMB + dir * ENC_H + r] + dh_s[r];
            const float tc = tanhf(c);
            const float dct = dc + dh * o * (1.f - tc * tc);
            dz_s[r] = dct * g * i * (1.f - i);                       // input gate
            dz_s[ENC_H + r] = dct * cprev * f * (1.f - f);           // forget gate
            dz_s[2 * ENC_H + r] = dct * i * (1.f - g * g);           // cell candidate
            dz_s[3 * ENC_H + r] = dh * tc * o * (1.f - o);           // output gate
            dc = dct * f;
            hprev[(((size_t)b * T + t) * 2 + dir) * ENC_H + r] = hp;
        }
        __syncthreads();
        dz_out[((size_t)b * T + t) * (8 * ENC_H) + dir * 4 * ENC_H + r] = dz_s[r];
        // dh_{prev}[k] = sum_r whh[r][k] * dz[r]: four partial sums over 256 rows each
        float acc = 0.f;
        const float* wp = whh + (size_t)(part * ENC_H) * ENC_H + k;
#pragma unroll 8
        for (int j = 0; j < ENC_H; ++j) acc = fmaf(__ldg(wp + (size_t)j * ENC_H), dz_s[part * ENC_H + j], acc);
        part_s[part][k] = acc;
        __syncthreads();
        if (r < ENC_H) dh_s[r] = (part_s[0][r] + part_s[1][r]) + (part_s[2][r] + part_s[3][r]);
        __syncthreads();
    }
}

__global__ void transpose2d_kernel(float* __restrict__ dst, const float* __restrict__ src, int R, int Cc) {
    const size_t total = (size_t)R * Cc;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int rr = i % R, cc = i / R;
        dst[i] = src[(size_t)rr * Cc + cc];            // dst [Cc][R]
    }
}

__global__ void embedding_bwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dx, float* __restrict__ dE, int n) {
    const int row = blockIdx.x, c = threadIdx.x;
    if (row >= n) return;
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= N_SYMBOLS ? N_SYMBOLS - 1 : id);
    for (int j = c; j < EMB; j += blockDim.x) atomicAdd(dE + (size_t)id * EMB + j, dx[(size_t)row * EMB + j]);
}

}  // namespace

int transpose2d(float* dst, const float* src, int R, int Cc, cudaStream_t st) {
    transpose2d_kernel<<<256, 256, 0, st>>>(dst, src, R, Cc);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int bilstm_train_fwd(const float* whhT_f, const float* whhT_b, const float* xproj, float* out, float* saved, int B, int T,
                     cudaStream_t st) {
    bilstm_train_fwd_kernel<<<2 * B, 1024, 0, st>>>(whhT_f, whhT_b, xproj, out, saved, B, T);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int bilstm_train_bwd(const float* whh_f, const float* whh_b, const float* d_out, const float* saved, float* dz, float* hprev,
                     int B, int T, cudaStream_t st) {
    bilstm_train_bwd_kernel<<<2 * B, 1024, 0, st>>>(whh_f, whh_b, d_out, saved, dz, hprev, B, T);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int embedding_bwd(const int64_t* ids, const float* dx, float* dE, int n, cudaStream_t st) {
    TACO_CUDA(cudaMemsetAsync(dE, 0, (size_t)N_SYMBOLS * EMB * sizeof(float), st));
    embedding_bwd_kernel<<<n, 128, 0, st>>>(ids, dx, dE, n);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco
