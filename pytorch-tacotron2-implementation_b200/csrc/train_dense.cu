// Training-mode building blocks of the dense layers (reference: nn.Conv1d + nn.BatchNorm1d in train mode +
// tanh/ReLU + F.dropout, src/model.py:25-39, :66-67 encoder; :347-407 postnet) and their backward passes.
//
//   forward   z = conv(x) + b                                   (dense.cu kernel, channels-last)
//             mean, var = batch statistics of z over all rows (pad positions included, as the reference does)
//             t = act((z - mean) * rstd * gamma + beta) ;  y = t * mask * scale        (mask: 0/1 floats from the host RNG)
//   backward  g  = dy * mask * scale * act'(t)
//             dbeta = sum_n g ; dgamma = sum_n g * zhat ; dz = gamma * rstd * (g - dbeta/N - zhat * dgamma/N)
//             db = sum_n dz ;  dW[tap][ci][co] = sum_n x[n + tap - pad][ci] * dz[n][co] ;  dx = conv(dz, W flipped^T)
//
// First correct version: fp32 SIMT everywhere, one kernel per step of the formulas above.
#include <cstdlib>

#include "handle.h"

namespace b200taco {

namespace {

constexpr float BN_EPS = 1e-5f;

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_TANH) return tanhf(v);
    return v;
}
__device__ __forceinline__ float act_bwd(float t, int act) {          // derivative expressed with the output t
    if (act == ACT_RELU) return t > 0.f ? 1.f : 0.f;
    if (act == ACT_TANH) return 1.f - t * t;
    return 1.f;
}

// ---- per-channel sums over rows: out[k][c] = sum_n f_k(n, c); block = 32 channels x 8 row lanes, rows split over gridDim.y
// MODE 0: sum z                         -> out0
// MODE 1: sum (z - mean)^2              -> out0
// MODE 2: g, g*zhat (BN backward)       -> out0 (dbeta), out1 (dgamma)
// MODE 3: sum dz                        -> out0
template <int MODE>
__global__ void __launch_bounds__(256) colreduce_kernel(const float* __restrict__ a, const float* __restrict__ b2,
                                                        const float* __restrict__ z, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ mask,
                                                        float scale, int act, int n, int C, float* __restrict__ out0,
                                                        float* __restrict__ out1) {
    __shared__ float s0[8][33], s1[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int rows_per = (n + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    float acc0 = 0.f, acc1 = 0.f;
    if (c < C) {
        const float m = (MODE == 1 || MODE == 2) ? mean[c] : 0.f;
        const float rs = MODE == 2 ? rstd[c] : 0.f, ga = MODE == 2 ? gamma[c] : 0.f, be = MODE == 2 ? beta[c] : 0.f;
        for (int r = r0 + rl; r < r1; r += 8) {
            const size_t i = (size_t)r * C + c;
            if (MODE == 0 || MODE == 3) {
                acc0 += a[i];
            } else if (MODE == 1) {
                const float d = a[i] - m;
                acc0 += d * d;
            } else {
                const float zh = (z[i] - m) * rs;
                const float t = act_fwd(zh * ga + be, act);
                const float g = a[i] * (mask ? mask[i] * scale : 1.f) * act_bwd(t, act);
                acc0 += g;
                acc1 += g * zh;
            }
        }
    }
    s0[rl][threadIdx.x & 31] = acc0;
    s1[rl][threadIdx.x & 31] = acc1;
    __syncthreads();
    if (rl == 0 && c < C) {
        float t0 = 0.f, t1 = 0.f;
        for (int k = 0; k < 8; ++k) { t0 += s0[k][threadIdx.x]; t1 += s1[k][threadIdx.x]; }
        atomicAdd(out0 + c, t0);
        if (MODE == 2) atomicAdd(out1 + c, t1);
    }
    (void)b2;
}

__global__ void bn_finish_stats_kernel(float* mean_sum, float* var_sum, float* rstd, int n, int C, int pass) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (pass == 0) mean_sum[c] = mean_sum[c] / (float)n;                         // sum -> mean
    else {
        const float v = var_sum[c] / (float)n;                                   // biased variance (normalisation)
        var_sum[c] = v;
        rstd[c] = rsqrtf(v + BN_EPS);
    }
}

__global__ void bn_act_drop_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mask, float scale, int act, float* __restrict__ y, size_t total,
                                   int C) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % C;
        const float t = act_fwd((z[i] - mean[c]) * rstd[c] * gamma[c] + beta[c], act);
        y[i] = mask ? t * mask[i] * scale : t;
    }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mask, float scale, int act,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float inv_n,
                                    float* __restrict__ dz, size_t total, int C) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % C;
        const float zh = (z[i] - mean[c]) * rstd[c];
        const float t = act_fwd(zh * gamma[c] + beta[c], act);
        const float g = dy[i] * (mask ? mask[i] * scale : 1.f) * act_bwd(t, act);
        dz[i] = gamma[c] * rstd[c] * (g - dbeta[c] * inv_n - zh * dgamma[c] * inv_n);
    }
}

// dW[tap][ci][co] += sum_n X[n + tap - pad][ci] * dY[n][co]; 64 x 64 tile of (ci, co) per CTA, rows split over gridDim.z
__global__ void __launch_bounds__(256) dense_wgrad_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY,
                                                          int ldy, float* __restrict__ dW, int n_rows, int seg_len, int c_in,
                                                          int c_out, int taps, int rows_per) {
    __shared__ __align__(16) float Xs[16][64 + 4];
    __shared__ __align__(16) float Ys[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
    const int tap = blockIdx.z % taps, chunk = blockIdx.z / taps;
    const int pad = (taps - 1) / 2, sh = tap - pad;
    const int r0 = chunk * rows_per, r1 = min(n_rows, r0 + rows_per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 4, lc = (tid & 15) * 4;               // loader: row lr of 16, 4 channels at lc
    for (int rb = r0; rb < r1; rb += 16) {
        const int r = rb + lr;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r1) {
            const int t = r % seg_len + sh;
            if (t >= 0 && t < seg_len && ci0 + lc < c_in) xv = *reinterpret_cast<const float4*>(X + (size_t)(r + sh) * ldx + ci0 + lc);
            if (co0 + lc < c_out) yv = *reinterpret_cast<const float4*>(dY + (size_t)r * ldy + co0 + lc);
        }
        *reinterpret_cast<float4*>(&Xs[lr][lc]) = xv;
        *reinterpret_cast<float4*>(&Ys[lr][lc]) = yv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[k][4 * ty]);
            const float4 b = *reinterpret_cast<const float4*>(&Ys[k][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + 4 * ty + i;
        if (ci >= c_in) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + 4 * tx + j;
            if (co < c_out) atomicAdd(dW + ((size_t)tap * c_in + ci) * c_out + co, acc[i][j]);
        }
    }
}

// The same reduction with 128 x 128 tiles of (ci, co), an 8 x 8 micro-tile per thread and register double buffering of the
// row slabs (the 64 x 64 kernel is bound by its shared-memory reads and exposes every load); plain stores when one CTA owns
// the whole row range.
__global__ void __launch_bounds__(256) dense_wgrad128_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY,
                                                             int ldy, float* __restrict__ dW, int n_rows, int seg_len, int c_in,
                                                             int c_out, int taps, int rows_per, int atomic) {
    __shared__ __align__(16) float Xs[2][16][128 + 4];
    __shared__ __align__(16) float Ys[2][16][128];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ci0 = blockIdx.x * 128, co0 = blockIdx.y * 128;
    const int tap = blockIdx.z % taps, chunk = blockIdx.z / taps;
    const int pad = (taps - 1) / 2, sh = tap - pad;
    const int r0 = chunk * rows_per, r1 = min(n_rows, r0 + rows_per);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 5, lc = (tid & 31) * 4;                // loader: rows lr and lr + 8 of 16, 4 channels at lc
    float4 xv[2], yv[2];
    auto fetch = [&](int rb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = rb + lr + 8 * h;
            xv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            yv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r1) {
                const int t = r % seg_len + sh;
                if (t >= 0 && t < seg_len && ci0 + lc < c_in) xv[h] = __ldg(reinterpret_cast<const float4*>(X + (size_t)(r + sh) * ldx + ci0 + lc));
                if (co0 + lc < c_out) yv[h] = __ldg(reinterpret_cast<const float4*>(dY + (size_t)r * ldy + co0 + lc));
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            *reinterpret_cast<float4*>(&Xs[buf][lr + 8 * h][lc]) = xv[h];
            *reinterpret_cast<float4*>(&Ys[buf][lr + 8 * h][lc]) = yv[h];
        }
    };
    if (r0 < r1) {
        fetch(r0);
        stash(0);
    }
    __syncthreads();
    int buf = 0;
    for (int rb = r0; rb < r1; rb += 16, buf ^= 1) {
        if (rb + 16 < r1) fetch(rb + 16);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][k][4 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][k][64 + 4 * ty]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ys[buf][k][4 * tx]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ys[buf][k][64 + 4 * tx]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (rb + 16 < r1) stash(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ci = ci0 + 64 * (i >> 2) + 4 * ty + (i & 3);
        if (ci >= c_in) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = co0 + 64 * (j >> 2) + 4 * tx + (j & 3);
            if (co >= c_out) continue;
            float* dst = dW + ((size_t)tap * c_in + ci) * c_out + co;
            if (atomic) atomicAdd(dst, acc[i][j]);
            else *dst = acc[i][j];
        }
    }
}

// W[tap][ci][co] -> Wf[taps-1-tap][co][ci] (the data-gradient convolution uses the forward kernel with these weights)
__global__ void flip_transpose_kernel(float* __restrict__ dst, const float* __restrict__ src, int taps, int c_in, int c_out) {
    const size_t total = (size_t)taps * c_in * c_out;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ci = i % c_in, co = (i / c_in) % c_out, tp = i / ((size_t)c_in * c_out);
        dst[i] = src[((size_t)(taps - 1 - tp) * c_in + ci) * c_out + co];
    }
}

// reference layout (co, ci, k) <-> packed [k][ci][co]
__global__ void conv_w_to_packed_kernel(float* __restrict__ dst, const float* __restrict__ w, int c_out, int c_in, int taps) {
    const size_t total = (size_t)c_out * c_in * taps;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int co = i % c_out, ci = (i / c_out) % c_in, k = i / ((size_t)c_out * c_in);
        dst[i] = w[((size_t)co * c_in + ci) * taps + k];
    }
}
__global__ void conv_w_from_packed_kernel(float* __restrict__ w, const float* __restrict__ src, int c_out, int c_in, int taps) {
    const size_t total = (size_t)c_out * c_in * taps;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int co = i % c_out, ci = (i / c_out) % c_in, k = i / ((size_t)c_out * c_in);
        w[((size_t)co * c_in + ci) * taps + k] = src[i];
    }
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = a[i] + b[i];
}

inline int blocks_for(size_t n) { return (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256); }

}  // namespace

int train_add(float* dst, const float* a, const float* b, size_t n, cudaStream_t st) {
    add_kernel<<<blocks_for(n), 256, 0, st>>>(dst, a, b, n);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int conv_w_to_packed(float* dst, const float* w, int c_out, int c_in, int taps, cudaStream_t st) {
    conv_w_to_packed_kernel<<<blocks_for((size_t)c_out * c_in * taps), 256, 0, st>>>(dst, w, c_out, c_in, taps);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int conv_w_from_packed(float* w, const float* src, int c_out, int c_in, int taps, cudaStream_t st) {
    conv_w_from_packed_kernel<<<blocks_for((size_t)c_out * c_in * taps), 256, 0, st>>>(w, src, c_out, c_in, taps);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

// dst[c] = sum_n src[n][c] for a dense [n][C] matrix
int train_colsum(float* dst, const float* src, int n_rows, int C, cudaStream_t st) {
    TACO_CUDA(cudaMemsetAsync(dst, 0, (size_t)C * sizeof(float), st));
    const dim3 grid((C + 31) / 32, n_rows >= 4096 ? 32 : (n_rows >= 256 ? 8 : 1));
    colreduce_kernel<3><<<grid, 256, 0, st>>>(src, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 0, n_rows, C,
                                              dst, nullptr);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

bool dense_wgrad_tc_supported(int n_rows, int c_in, int c_out);
int dense_wgrad_tc(const float* X, int ldx, const float* dY, int ldy, float* dW, int n_rows, int seg_len, int c_in, int c_out, int taps,
                   int num_sms, cudaStream_t st);

// dW[tap][ci][co] = sum_n X[n + tap - pad][ci] * dY[n][co]  (overwrites dW)
int dense_wgrad(const float* X, int ldx, const float* dY, int ldy, float* dW, int n_rows, int seg_len, int c_in, int c_out,
                int taps, cudaStream_t st) {
    if (dense_wgrad_tc_supported(n_rows, c_in, c_out)) {     // tensor cores (3xTF32), train_wgrad_tc.cu
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return dense_wgrad_tc(X, ldx, dY, ldy, dW, n_rows, seg_len, c_in, c_out, taps, sms, st);
    }
    static const bool wide = [] { const char* e = getenv("TACO_WGRAD128"); return !(e && e[0] == '0'); }();
    if (wide && c_in >= 128 && c_out >= 128) {
        const int tiles2 = ((c_in + 127) / 128) * ((c_out + 127) / 128) * taps;
        int chunks2 = 1;
        while (tiles2 * chunks2 < 148 && n_rows / (chunks2 * 2) >= 256) chunks2 *= 2;
        const int rows_per2 = ((n_rows + chunks2 - 1) / chunks2 + 15) / 16 * 16;
        if (chunks2 > 1) TACO_CUDA(cudaMemsetAsync(dW, 0, (size_t)taps * c_in * c_out * sizeof(float), st));
        dense_wgrad128_kernel<<<dim3((c_in + 127) / 128, (c_out + 127) / 128, taps * chunks2), 256, 0, st>>>(
            X, ldx, dY, ldy, dW, n_rows, seg_len, c_in, c_out, taps, rows_per2, chunks2 > 1 ? 1 : 0);
        TACO_LAUNCHED();
        return B200TACO_OK;
    }
    TACO_CUDA(cudaMemsetAsync(dW, 0, (size_t)taps * c_in * c_out * sizeof(float), st));
    int chunks = 1;
    const int tiles = ((c_in + 63) / 64) * ((c_out + 63) / 64) * taps;
    while (tiles * chunks < 592 && n_rows / (chunks * 2) >= 256) chunks *= 2;
    const int rows_per = ((n_rows + chunks - 1) / chunks + 15) / 16 * 16;
    dense_wgrad_kernel<<<dim3((c_in + 63) / 64, (c_out + 63) / 64, taps * chunks), 256, 0, st>>>(
        X, ldx, dY, ldy, dW, n_rows, seg_len, c_in, c_out, taps, rows_per);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

size_t convbn_saved_floats(int n_rows, int c_out) { return (size_t)2 * n_rows * c_out + 3 * (size_t)c_out; }

// Forward of one conv + BatchNorm(train) + activation + dropout layer.  `saved` receives z [n][co], y [n][co], mean[co],
// var[co] (biased), rstd[co]; y is also the layer output.  Wp = weights as [tap][ci][co].
// The convolutions of the training passes run on the tensor cores (conv_tc_kernel, 3xTF32 at fp32 accuracy, dense_tc.cu)
// when the caller supplies `wtc` (room for the pre-split weight planes, dense_tc_weight_floats) and the problem is large
// enough to fill the machine; small problems (the golden-vector tests) keep the bit-reproducible CUDA-core kernel.
// TACO_TRAIN_CONV_TC=0 disables it (measurements).
int conv1d_channels_last_tc(const float* X, int ldx, const float* Wtc, const float* scale, const float* shift,
                            const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len, int c_in,
                            int c_out, int taps, int act, int num_sms, cudaStream_t stream, const ChannelMajorOut& cm);
static int train_conv(const float* X, int ldx, const float* Wp, float* wtc, const float* shift, float* Y, int ldy, int n_rows,
                      int seg_len, int c_in, int c_out, int taps, cudaStream_t st) {
    static const bool allow = [] { const char* e = getenv("TACO_TRAIN_CONV_TC"); return !(e && e[0] == '0'); }();
    static const int sms = [] { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
    if (allow && wtc != nullptr && n_rows >= TRAIN_CONV_TC_MIN_ROWS && dense_tc_supported(c_in, c_out, ldx, ldy)) {
        if (int e = pack_dense_tc(wtc, Wp, taps, c_in, c_out, st)) return e;
        return conv1d_channels_last_tc(X, ldx, wtc, nullptr, shift, nullptr, 0, Y, ldy, n_rows, seg_len, c_in, c_out, taps, ACT_NONE,
                                       sms, st, ChannelMajorOut{});
    }
    return conv1d_channels_last(X, ldx, Wp, nullptr, shift, nullptr, 0, Y, ldy, n_rows, seg_len, c_in, c_out, taps, ACT_NONE, st);
}

int convbn_train_fwd(const float* X, int ldx, const float* Wp, const float* conv_b, const float* gamma, const float* beta,
                     const float* mask, float drop_scale, int act, float* saved, int n_rows, int seg_len, int c_in, int c_out,
                     int taps, float* wtc, cudaStream_t st) {
    float* z = saved;
    float* y = z + (size_t)n_rows * c_out;
    float* mean = y + (size_t)n_rows * c_out;
    float* var = mean + c_out;
    float* rstd = var + c_out;
    if (int e = train_conv(X, ldx, Wp, wtc, conv_b, z, c_out, n_rows, seg_len, c_in, c_out, taps, st)) return e;
    TACO_CUDA(cudaMemsetAsync(mean, 0, 3 * (size_t)c_out * sizeof(float), st));
    const dim3 grid((c_out + 31) / 32, n_rows >= 4096 ? 32 : (n_rows >= 256 ? 8 : 1));
    colreduce_kernel<0><<<grid, 256, 0, st>>>(z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 0, n_rows,
                                              c_out, mean, nullptr);
    TACO_LAUNCHED();
    bn_finish_stats_kernel<<<(c_out + 127) / 128, 128, 0, st>>>(mean, var, rstd, n_rows, c_out, 0);
    TACO_LAUNCHED();
    colreduce_kernel<1><<<grid, 256, 0, st>>>(z, nullptr, nullptr, mean, nullptr, nullptr, nullptr, nullptr, 1.f, 0, n_rows,
                                              c_out, var, nullptr);
    TACO_LAUNCHED();
    bn_finish_stats_kernel<<<(c_out + 127) / 128, 128, 0, st>>>(mean, var, rstd, n_rows, c_out, 1);
    TACO_LAUNCHED();
    const size_t total = (size_t)n_rows * c_out;
    bn_act_drop_kernel<<<blocks_for(total), 256, 0, st>>>(z, mean, rstd, gamma, beta, mask, drop_scale, act, y, total, c_out);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

// Backward of the same layer.  dY [n][co] -> dX [n][ci] (nullable), dWp [tap][ci][co], db, dgamma, dbeta (all overwritten).
// scratch: dz = n_rows * c_out floats, wflip = taps * c_in * c_out floats (flipped weights, only if dX is wanted).
int convbn_train_bwd(const float* dY, const float* X, int ldx, const float* Wp, const float* gamma, const float* beta,
                     const float* mask, float drop_scale, int act, const float* saved, float* dX, int lddx, float* dWp,
                     float* db, float* dgamma, float* dbeta, float* dz, float* wflip, int n_rows, int seg_len, int c_in,
                     int c_out, int taps, float* wtc, cudaStream_t st) {
    const float* z = saved;
    const float* mean = z + (size_t)2 * n_rows * c_out;
    const float* rstd = mean + 2 * c_out;
    TACO_CUDA(cudaMemsetAsync(dgamma, 0, c_out * sizeof(float), st));
    TACO_CUDA(cudaMemsetAsync(dbeta, 0, c_out * sizeof(float), st));
    TACO_CUDA(cudaMemsetAsync(db, 0, c_out * sizeof(float), st));
    TACO_CUDA(cudaMemsetAsync(dWp, 0, (size_t)taps * c_in * c_out * sizeof(float), st));
    const dim3 grid((c_out + 31) / 32, n_rows >= 4096 ? 32 : (n_rows >= 256 ? 8 : 1));
    colreduce_kernel<2><<<grid, 256, 0, st>>>(dY, nullptr, z, mean, rstd, gamma, beta, mask, drop_scale, act, n_rows, c_out,
                                              dbeta, dgamma);
    TACO_LAUNCHED();
    const size_t total = (size_t)n_rows * c_out;
    bn_bwd_apply_kernel<<<blocks_for(total), 256, 0, st>>>(dY, z, mean, rstd, gamma, beta, mask, drop_scale, act, dgamma, dbeta,
                                                           1.f / (float)n_rows, dz, total, c_out);
    TACO_LAUNCHED();
    colreduce_kernel<3><<<grid, 256, 0, st>>>(dz, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 0, n_rows,
                                              c_out, db, nullptr);
    TACO_LAUNCHED();
    // weight gradient (tensor cores where the channel counts allow, train_wgrad_tc.cu)
    if (int e = dense_wgrad(X, ldx, dz, c_out, dWp, n_rows, seg_len, c_in, c_out, taps, st)) return e;
    if (dX != nullptr) {
        flip_transpose_kernel<<<blocks_for((size_t)taps * c_in * c_out), 256, 0, st>>>(wflip, Wp, taps, c_in, c_out);
        TACO_LAUNCHED();
        if (int e = train_conv(dz, c_out, wflip, wtc, nullptr, dX, lddx, n_rows, seg_len, c_out, c_in, taps, st)) return e;
    }
    return B200TACO_OK;
}

}  // namespace b200taco
