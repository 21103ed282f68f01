// Dense layers of the encoder / postnet / projections as one channels-last conv1d kernel:
//   Y[n][co] = act( (sum_{tap,ci} X[n + tap - pad][ci] * W[tap][ci][co]) * scale[co] + shift[co] ) (+ R[n][co])
// rows n = b * seg_len + t; taps never cross an utterance boundary (zero padding, as
// nn.Conv1d(padding=(k-1)/2) does: src/model.py:26-33, :349-357).  taps == 1 is nn.Linear.
// fp32 SIMT tiles (64 rows x 64 channels, 4x4 per thread): exact-precision first version of
// these layers; they are the tensor-core candidates named in DESIGN.md.
#include <cstdlib>

#include "handle.h"

namespace b200taco {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS = BM + 4;

template <int ACT>
__global__ void __launch_bounds__(256) conv_cl_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      const float* __restrict__ R, int ldres, float* __restrict__ Y,
                                                      int ldy, int n_rows, int seg_len, int c_in, int c_out, int taps) {
    __shared__ __align__(16) float Xs[2][BK][LDS];
    __shared__ __align__(16) float Ws[2][BK][BN];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * BM, co0 = blockIdx.y * BN;
    const int pad = (taps - 1) / 2;
    const int tx = tid & 15, ty = tid >> 4;
    // loader roles
    const int xm = tid >> 2, xq = tid & 3;          // X: row xm, channels 4*xq..
    const int wk = tid >> 4, wq = tid & 15;         // W: k row wk, channels 4*wq..
    const int xrow = row0 + xm;
    const int xt = xrow % seg_len;
    const int kchunks = c_in / BK;
    const int iters = taps * kchunks;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 xr, wr;
    auto fetch = [&](int it) {
        const int tap = it / kchunks, ci0 = (it % kchunks) * BK;
        const int tt = xt + tap - pad;
        xr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xrow < n_rows && tt >= 0 && tt < seg_len)
            xr = __ldg(reinterpret_cast<const float4*>(X + (size_t)(xrow + tap - pad) * ldx + ci0 + 4 * xq));
        wr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (co0 + 4 * wq < c_out)
            wr = __ldg(reinterpret_cast<const float4*>(W + ((size_t)tap * c_in + ci0 + wk) * c_out + co0 + 4 * wq));
    };
    auto stash = [&](int buf) {
        Xs[buf][4 * xq + 0][xm] = xr.x;
        Xs[buf][4 * xq + 1][xm] = xr.y;
        Xs[buf][4 * xq + 2][xm] = xr.z;
        Xs[buf][4 * xq + 3][xm] = xr.w;
        *reinterpret_cast<float4*>(&Ws[buf][wk][4 * wq]) = wr;
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) fetch(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[buf][kk][4 * ty]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][kk][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < iters) stash(buf ^ 1);
        __syncthreads();
    }
    const int co = co0 + 4 * tx;
    if (co >= c_out) return;
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (scale) { const float4 v = __ldg(reinterpret_cast<const float4*>(scale + co)); sc[0] = v.x; sc[1] = v.y; sc[2] = v.z; sc[3] = v.w; }
    if (shift) { const float4 v = __ldg(reinterpret_cast<const float4*>(shift + co)); sh[0] = v.x; sh[1] = v.y; sh[2] = v.z; sh[3] = v.w; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + 4 * ty + i;
        if (row >= n_rows) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = fmaf(acc[i][j], sc[j], sh[j]);
            if (ACT == ACT_RELU) v = fmaxf(v, 0.f);
            if (ACT == ACT_TANH) v = tanhf(v);
            o[j] = v;
        }
        if (R) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(R + (size_t)row * ldres + co));
            o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
        }
        *reinterpret_cast<float4*>(Y + (size_t)row * ldy + co) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// The same contract with 128 x 128 tiles and an 8 x 8 micro-tile per thread (two 4-wide halves 64 apart in each direction):
// four 128-bit shared-memory reads feed 64 FMAs instead of two feeding 16, which is what bounds the 64 x 64 kernel
// (32 TFLOP/s).  Every output element accumulates its (tap, channel) products in the same order as in conv_cl_kernel, so
// the two kernels are bit-identical; this one is used when its tiles fill the machine.
// BN2 = 64: 128 x 64 tiles (8 x 4 micro-tile) - twice the tiles for problems that would not fill the machine with 128 x 128.
constexpr int BM2 = 128, LDS2 = BM2 + 4;
template <int ACT, int BN2>
__global__ void __launch_bounds__(256) conv_cl128_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         const float* __restrict__ R, int ldres, float* __restrict__ Y,
                                                         int ldy, int n_rows, int seg_len, int c_in, int c_out, int taps) {
    __shared__ __align__(16) float Xs[2][BK][LDS2];
    __shared__ __align__(16) float Ws[2][BK][BN2];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * BM2, co0 = blockIdx.y * BN2;
    const int pad = (taps - 1) / 2;
    const int tx = tid & 15, ty = tid >> 4;
    // loader roles: X rows xm and xm + 64, channels 4*xq.. ; W k row wk, channels 4*wq.. and 64 + 4*wq..
    const int xm = tid >> 2, xq = tid & 3;
    const int wk = tid >> 4, wq = tid & 15;
    const int kchunks = c_in / BK;
    const int iters = taps * kchunks;
    int xrow[2], xt[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        xrow[h] = row0 + xm + 64 * h;
        xt[h] = xrow[h] % seg_len;
    }
    constexpr int NH = BN2 / 64, NJ = 4 * NH;             // column halves, columns per thread
    float acc[8][NJ];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    float4 xr[2], wr[NH];
    auto fetch = [&](int it) {
        const int tap = it / kchunks, ci0 = (it % kchunks) * BK;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tt = xt[h] + tap - pad;
            xr[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xrow[h] < n_rows && tt >= 0 && tt < seg_len)
                xr[h] = __ldg(reinterpret_cast<const float4*>(X + (size_t)(xrow[h] + tap - pad) * ldx + ci0 + 4 * xq));
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            wr[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (co0 + 64 * h + 4 * wq < c_out)
                wr[h] = __ldg(reinterpret_cast<const float4*>(W + ((size_t)tap * c_in + ci0 + wk) * c_out + co0 + 64 * h + 4 * wq));
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            Xs[buf][4 * xq + 0][xm + 64 * h] = xr[h].x;
            Xs[buf][4 * xq + 1][xm + 64 * h] = xr[h].y;
            Xs[buf][4 * xq + 2][xm + 64 * h] = xr[h].z;
            Xs[buf][4 * xq + 3][xm + 64 * h] = xr[h].w;
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) *reinterpret_cast<float4*>(&Ws[buf][wk][64 * h + 4 * wq]) = wr[h];
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) fetch(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][kk][4 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][kk][64 + 4 * ty]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[NJ];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 * h + 4 * tx]);
                bv[4 * h + 0] = b.x; bv[4 * h + 1] = b.y; bv[4 * h + 2] = b.z; bv[4 * h + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < iters) stash(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int hj = 0; hj < NH; ++hj) {
        const int co = co0 + 64 * hj + 4 * tx;
        if (co >= c_out) continue;
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (scale) { const float4 v = __ldg(reinterpret_cast<const float4*>(scale + co)); sc[0] = v.x; sc[1] = v.y; sc[2] = v.z; sc[3] = v.w; }
        if (shift) { const float4 v = __ldg(reinterpret_cast<const float4*>(shift + co)); sh[0] = v.x; sh[1] = v.y; sh[2] = v.z; sh[3] = v.w; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + 64 * (i >> 2) + 4 * ty + (i & 3);
            if (row >= n_rows) continue;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = fmaf(acc[i][4 * hj + j], sc[j], sh[j]);
                if (ACT == ACT_RELU) v = fmaxf(v, 0.f);
                if (ACT == ACT_TANH) v = tanhf(v);
                o[j] = v;
            }
            if (R) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(R + (size_t)row * ldres + co));
                o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
            }
            *reinterpret_cast<float4*>(Y + (size_t)row * ldy + co) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

__global__ void embedding_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table, float* __restrict__ out,
                                 int n) {
    const int row = blockIdx.x, c = threadIdx.x * 4;
    if (row >= n) return;
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= N_SYMBOLS ? N_SYMBOLS - 1 : id);
    *reinterpret_cast<float4*>(out + (size_t)row * EMB + c) =
        __ldg(reinterpret_cast<const float4*>(table + (size_t)id * EMB + c));
}

__global__ void transpose_bct_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int T) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? in[((size_t)b * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (c < C && t < T) out[((size_t)b * T + t) * C + c] = tile[threadIdx.x][i];
    }
}

}  // namespace

int conv1d_channels_last(const float* X, int ldx, const float* W, const float* scale, const float* shift,
                         const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len, int c_in,
                         int c_out, int taps, int act, cudaStream_t stream) {
    TACO_REQUIRE(c_in % BK == 0 && c_out % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, B200TACO_ERR_INVALID,
                 "conv1d: c_in %d must be a multiple of %d, c_out %d / ldx %d / ldy %d multiples of 4", c_in, BK, c_out,
                 ldx, ldy);
    if (n_rows <= 0) return B200TACO_OK;
    // wide tiles once they fill the machine (one CTA per SM); TACO_CONV128=0 keeps the 64 x 64 kernel (measurements)
    static const bool wide = [] { const char* e = getenv("TACO_CONV128"); return !(e && e[0] == '0'); }();
    static const int sms = [] { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
    // 128-row tiles when they fill the machine with >= 4 tiles per SM (with fewer, e.g. the 152 tiles of a 32-utterance
    // encoder layer at 128 x 128 on 148 SMs, the last tiles run alone and finer tiles win - measured): 128 x 128, else 128 x 64
    const int mt2 = (n_rows + BM2 - 1) / BM2;
    const int bn2 = !wide ? 0 : (c_out % 128 == 0 && mt2 * (c_out / 128) >= 4 * sms ? 128
                                 : (c_out % 64 == 0 && mt2 * (c_out / 64) >= 4 * sms ? 64 : 0));
    if (bn2 != 0) {
        dim3 grid2(mt2, c_out / bn2);
#define TACO_CL128(A, N)                                                                                                    \
    conv_cl128_kernel<A, N><<<grid2, 256, 0, stream>>>(X, ldx, W, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, \
                                                       c_in, c_out, taps)
        if (bn2 == 128) {
            if (act == ACT_NONE) TACO_CL128(ACT_NONE, 128);
            else if (act == ACT_RELU) TACO_CL128(ACT_RELU, 128);
            else TACO_CL128(ACT_TANH, 128);
        } else {
            if (act == ACT_NONE) TACO_CL128(ACT_NONE, 64);
            else if (act == ACT_RELU) TACO_CL128(ACT_RELU, 64);
            else TACO_CL128(ACT_TANH, 64);
        }
#undef TACO_CL128
        TACO_LAUNCHED();
        return B200TACO_OK;
    }
    dim3 grid((n_rows + BM - 1) / BM, (c_out + BN - 1) / BN);
    switch (act) {
        case ACT_NONE:
            conv_cl_kernel<ACT_NONE><<<grid, 256, 0, stream>>>(X, ldx, W, scale, shift, residual, ldres, Y, ldy, n_rows,
                                                               seg_len, c_in, c_out, taps);
            break;
        case ACT_RELU:
            conv_cl_kernel<ACT_RELU><<<grid, 256, 0, stream>>>(X, ldx, W, scale, shift, residual, ldres, Y, ldy, n_rows,
                                                               seg_len, c_in, c_out, taps);
            break;
        default:
            conv_cl_kernel<ACT_TANH><<<grid, 256, 0, stream>>>(X, ldx, W, scale, shift, residual, ldres, Y, ldy, n_rows,
                                                               seg_len, c_in, c_out, taps);
            break;
    }
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int embedding_gather(const int64_t* ids, const float* table, float* out, int n, cudaStream_t stream) {
    if (n <= 0) return B200TACO_OK;
    embedding_kernel<<<n, EMB / 4, 0, stream>>>(ids, table, out, n);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int transpose_bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t stream) {
    dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
    transpose_bct_kernel<<<grid, block, 0, stream>>>(in, out, B, C, T);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco
